/*
 * Minimal C host for the C ABI (include/cytvdn_b200.h): TV-denoise a small synthetic 4-D block with the
 * loop-level entry point, exactly what cyTVDN.denoise4D(data, mu, iterations=20, FISTA=True) does.
 *
 *   gcc examples/denoise_c_abi.c -Iinclude -Lcytvdn_b200 -lcytvdn_b200 -Wl,-rpath,$PWD/cytvdn_b200 -lm -o /tmp/denoise_c_abi
 *
 * Without a GPU it prints the ABI version and exits 0 (nothing to compute on: there is no CPU fallback).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "cytvdn_b200.h"

int main(void)
{
    int ndev = 0;
    printf("cytvdn ABI version %d\n", cytvdn_version());
    if (cytvdn_device_count(&ndev) != CYTVDN_OK || ndev < 1) {
        printf("no CUDA device: skipping the computation\n");
        return 0;
    }
    enum { N0 = 8, N1 = 8, N2 = 16, N3 = 32, ITERS = 20 };
    const size_t n = (size_t)N0 * N1 * N2 * N3;
    float *data = malloc(n * sizeof *data), *recon = malloc(n * sizeof *recon);
    unsigned s = 12345u;
    for (size_t x = 0; x < n; ++x) {                 /* a step edge plus noise, count-like values */
        s = s * 1664525u + 1013904223u;
        data[x] = (float)(((x / N3) % N2 < N2 / 2 ? 100 : 300) + (int)(s >> 24) % 40);
    }
    const double mu[4] = {1, 1, .5, .5};
    cytvdn_denoise_params p = {0};
    p.ndim = 4; p.dtype = CYTVDN_F32;
    p.shape[0] = N0; p.shape[1] = N1; p.shape[2] = N2; p.shape[3] = N3;
    p.iters_fista = ITERS; p.bc_mode = 2; p.device = -1;
    for (int k = 0; k < 4; ++k) {
        const float lam = (float)mu[k] / 32.0f;       /* cyTVDN.py:67-68 */
        p.clip[k] = 1.0f / lam;                       /* lambdaInv, cyTVDN.py:77 */
        p.lambda_mu[k] = lam / (float)mu[k];          /* cyTVDN.py:78 */
    }
    double bnorm[ITERS], delta[ITERS], ms[3];
    int32_t done[3];
    if (cytvdn_denoise(&p, data, recon, NULL, bnorm, delta, NULL, done, ms) != CYTVDN_OK) {
        fprintf(stderr, "cytvdn_denoise failed: %s\n", cytvdn_last_error());
        return 1;
    }
    double change = 0;
    for (size_t x = 0; x < n; ++x) change = fmax(change, fabs((double)recon[x] - data[x]));
    printf("%d FISTA iterations (schedule %d), loop %.3f ms, delta[0]=%.3e delta[last]=%.3e, max |recon-data| = %.2f\n",
           done[0], done[2] & 0xff, ms[1], delta[0], delta[ITERS - 1], change);

    /* The same run sharded over scan axis 0 from this one process (what cyTVDN/mpi.py does with one MPI rank per
       tile): every device -- two GPUs when the machine has them, else the same GPU twice -- takes a tile with one
       overlap plane, the halo planes are pushed by the copy engines under the interior sweep.  The result must equal
       the single-GPU one bit for bit. */
    {
        float *recon2 = malloc(n * sizeof *recon2);
        double bnorm2[ITERS], delta2[ITERS];
        const int devices[2] = {0, ndev > 1 ? 1 : 0};
        if (cytvdn_denoise_sharded(&p, 2, devices, data, recon2, bnorm2, delta2, done, ms) != CYTVDN_OK) {
            fprintf(stderr, "cytvdn_denoise_sharded failed: %s\n", cytvdn_last_error());
            return 1;
        }
        size_t diff = 0;
        for (size_t x = 0; x < n; ++x) diff += recon2[x] != recon[x];
        printf("sharded over devices {%d, %d}: %d iterations on %d shards, %zu voxels differ from the single-GPU result, "
               "delta[last]=%.3e\n", devices[0], devices[1], done[0], done[2] >> 8, diff, delta2[ITERS - 1]);
        free(recon2);
        if (diff) return 3;
    }
    free(data); free(recon);
    return delta[ITERS - 1] < delta[0] ? 0 : 2;
}
