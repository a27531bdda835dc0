/*
 * tv_oracle.c -- CPU restatement of the cyTVDN hot path.   TEST INFRASTRUCTURE ONLY.
 *
 * Nothing under cytvdn_b200/ may call, link or import this file; it exists so that tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg can check / time the CUDA path
 * against an independent statement of the reference arithmetic.
 *
 * Parity pin: this file is checked bit-for-bit (arrays AND the array-dtype scalars at one
 * thread) against the compiled, unmodified reference kernels in oracle/_ref (built by
 * oracle/build_ref.py) by tests/test_oracle_pin.py, and against the golden vectors in
 * tests/golden (produced by the reference's own Python driver, tests/golden/make_golden.py).
 *
 * Every function states the per-voxel arithmetic of SURVEY.md section 3.4 and cites the
 * reference lines it follows.  It is written from that specification (flat indices, one
 * generic 4-axis routine; 3-D arrays are passed as shape (N0,N1,N2,1)), not transcribed
 * from the .pyx loops.
 *
 * Build (see oracle/Makefile):  gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC
 *   -ffp-contract=off: the reference is built for baseline x86-64 (no FMA), every product
 *   and sum is rounded separately; the oracle must do the same on any host.
 *
 * Scalars: the reference accumulates Sigma|b|, Sigma|delta|, Sigma|old| in the ARRAY dtype
 * (anisotropic.pyx:38, utils.pyx:81-82), which is wrong by 0.1%..90% in fp32 at >=4M voxels
 * (SURVEY.md section 7.3-1).  Each routine therefore exists in two flavours:
 *   *_T   accumulators in the array dtype, same summation order as the reference at
 *         OMP_NUM_THREADS=1 (interior region first, then the boundary slab) -> bit-equal
 *         to the reference's returned scalars when run with one thread;
 *   *_D   accumulators in double: the "truth" the GPU path is compared with.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#ifdef _OPENMP
#include <omp.h>
#endif

int tvo_version(void) { return 1; }
int tvo_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void tvo_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

typedef int64_t i64;

/* min(max(a,-c),c) by comparisons (anisotropic.pyx:11-12; Cython expands max(a,b) to
 * (b > a) ? b : a and min(a,b) to (b < a) ? b : a, so a NaN input is returned unchanged). */
#define CLIPVAL(a, c, T) __extension__({ T _a = (a), _c = (c), _m = -_c;          \
                              T _t = (_m > _a) ? _m : _a; (_c < _t) ? _c : _t; })

/* ------------------------------------------------------------------------------------ */
/* half-step A, anisotropic, one axis   (anisotropic.pyx:17-84 / :89-164 / :169-317)     */
/*   g = a[x] - a[x - e_ax]            (x_ax >= 1)                        :46-54         */
/*   g = a[x] - a[x + delta e_ax]      (x_ax == 0; delta = N-1 | 1 | 0 for BC 0|1|2) :60-82 */
/*   v = clip(g + b[x]);  plain: b=v, norm+=|v|                                           */
/*   FISTA: bn = v + tk*(v - d[x]); b=bn; d=v; norm+=|bn|                  :127-132       */
/* ------------------------------------------------------------------------------------ */
#define DEF_ACC(NAME, T, ACC, FABS)                                                         \
double NAME(const T *a, T *b, T *d, const i64 *shape, int ax, T tk, T clip, int bc_mode)    \
{                                                                                           \
    const i64 n0 = shape[0], n1 = shape[1], n2 = shape[2], n3 = shape[3];                   \
    const i64 st[4] = { n1 * n2 * n3, n2 * n3, n3, 1 };                                     \
    const i64 nax = shape[ax], sax = st[ax];                                                \
    const i64 bdelta = (bc_mode == 0) ? (nax - 1) * sax : (bc_mode == 1) ? sax : 0;         \
    ACC norm = 0;                                                                           \
    const i64 rows = n0 * n1;                                                               \
    /* region 1: x_ax >= 1 (every voxel has its backward neighbour inside the array) */     \
    _Pragma("omp parallel for reduction(+:norm) schedule(static)")                          \
    for (i64 ij = 0; ij < rows; ++ij) {                                                     \
        const i64 i = ij / n1, j = ij % n1;                                                 \
        if ((ax == 0 && i == 0) || (ax == 1 && j == 0)) continue;                           \
        for (i64 k = (ax == 2); k < n2; ++k)                                                \
            for (i64 l = (ax == 3); l < n3; ++l) {                                          \
                const i64 x = ij * st[1] + k * n3 + l;                                      \
                T v = CLIPVAL(a[x] - a[x - sax] + b[x], clip, T);                           \
                if (d) { T bn = v + tk * (v - d[x]); b[x] = bn; d[x] = v; norm += FABS(bn); } \
                else   { b[x] = v; norm += FABS(v); }                                       \
            }                                                                               \
    }                                                                                       \
    /* region 2: the slab x_ax == 0 (serial in the reference, :75-82) */                    \
    for (i64 i = 0; i < (ax == 0 ? 1 : n0); ++i)                                            \
      for (i64 j = 0; j < (ax == 1 ? 1 : n1); ++j)                                          \
        for (i64 k = 0; k < (ax == 2 ? 1 : n2); ++k)                                        \
          for (i64 l = 0; l < (ax == 3 ? 1 : n3); ++l) {                                    \
                const i64 x = i * st[0] + j * st[1] + k * n3 + l;                           \
                T v = CLIPVAL(a[x] - a[x + bdelta] + b[x], clip, T);                        \
                if (d) { T bn = v + tk * (v - d[x]); b[x] = bn; d[x] = v; norm += FABS(bn); } \
                else   { b[x] = v; norm += FABS(v); }                                       \
          }                                                                                 \
    return (double)norm;                                                                    \
}

DEF_ACC(tvo_acc_f32_T, float,  float,  fabs)
DEF_ACC(tvo_acc_f32_D, float,  double, fabs)
DEF_ACC(tvo_acc_f64_T, double, double, fabs)
DEF_ACC(tvo_acc_f64_D, double, double, fabs)

/* ------------------------------------------------------------------------------------ */
/* half-step A, half-isotropic, axis pair (p,q)  (halfisotropic.pyx:17-97 / :102-188)    */
/*   dp = a[x] - a[x - e_p*(x_p>0)] + bp[x]   (backward difference is 0 at x_p == 0) :70-85 */
/*   m  = (T) hypot((double)dp,(double)dq)                                  :87           */
/*   if m > clip: dp = dp/(m/clip); dq = dq/(m/clip)                        :89-91        */
/*   plain: bp=dp, bq=dq, norm += |dp|+|dq|                                 :93-95        */
/*   FISTA: bpn = dp + tk*(dp - d_p[x]) ...; norm += |bpn|+|bqn|; d_p=dp    :176-186      */
/* Race-free semantics (the reference's shared stride scratch races with >1 thread,      */
/* SURVEY.md section 0-4); summation order = one pass over all voxels.                   */
/* ------------------------------------------------------------------------------------ */
#define DEF_ISO(NAME, T, ACC, FABS)                                                         \
double NAME(const T *a, T *b1, T *b2, T *d1, T *d2, const i64 *shape, int ax1, int ax2,     \
            T tk, T clip)                                                                   \
{                                                                                           \
    const i64 n0 = shape[0], n1 = shape[1], n2 = shape[2], n3 = shape[3];                   \
    const i64 st[4] = { n1 * n2 * n3, n2 * n3, n3, 1 };                                     \
    ACC norm = 0;                                                                           \
    const i64 rows = n0 * n1;                                                               \
    _Pragma("omp parallel for reduction(+:norm) schedule(static)")                          \
    for (i64 ij = 0; ij < rows; ++ij) {                                                     \
        i64 c[4]; c[0] = ij / n1; c[1] = ij % n1;                                           \
        for (c[2] = 0; c[2] < n2; ++c[2])                                                   \
            for (c[3] = 0; c[3] < n3; ++c[3]) {                                             \
                const i64 x = ij * st[1] + c[2] * n3 + c[3];                                \
                T dp = a[x] - a[x - (c[ax1] > 0 ? st[ax1] : 0)] + b1[x];                    \
                T dq = a[x] - a[x - (c[ax2] > 0 ? st[ax2] : 0)] + b2[x];                    \
                T m = (T)hypot((double)dp, (double)dq);                                     \
                if (m > clip) { dp = dp / (m / clip); dq = dq / (m / clip); }               \
                if (d1) {                                                                   \
                    T p = dp + tk * (dp - d1[x]);                                           \
                    T q = dq + tk * (dq - d2[x]);                                           \
                    b1[x] = p; b2[x] = q;                                                   \
                    norm += FABS(p) + FABS(q);                                              \
                    d1[x] = dp; d2[x] = dq;                                                 \
                } else {                                                                    \
                    norm += FABS(dp) + FABS(dq);                                            \
                    b1[x] = dp; b2[x] = dq;                                                 \
                }                                                                           \
            }                                                                               \
    }                                                                                       \
    return (double)norm;                                                                    \
}

DEF_ISO(tvo_iso_f32_T, float,  float,  fabs)
DEF_ISO(tvo_iso_f32_D, float,  double, fabs)
DEF_ISO(tvo_iso_f64_T, double, double, fabs)
DEF_ISO(tvo_iso_f64_D, double, double, fabs)

/* ------------------------------------------------------------------------------------ */
/* half-step B  (utils.pyx:54-125 4-D, :131-199 3-D; BC 0 and 2 share one path :85,:163) */
/*   old = u[x]                                                                           */
/*   u[x] = f[x] - (((w0*(b0[x]-b0[x+e0 mod N0]) + w1*(...)) + w2*(...)) [+ w3*(...)])    */
/*   delta += |u[x]-old| ; rnorm += |old| ;  returns the two sums (ratio taken by caller) */
/* nterms = 3 for a 3-D array passed as (N0,N1,N2,1), 4 for 4-D.                          */
/* BC_mode 1 is undefined behaviour in the reference (utils.pyx:117-120,192-197: the      */
/* forward index is written max(i+1, N-1), which reads out of bounds) and is not restated.*/
/* mirror != 0 is NOT reference behaviour: it is the evident intent of those lines, the   */
/* forward index CLAMPED to min(i+1, N-1) -- the spec of this repo's BC_mode 3.           */
/* ------------------------------------------------------------------------------------ */
#define DEF_DCU(NAME, T, ACC, FABS)                                                         \
void NAME(const T *f, T *u, const T *b0, const T *b1, const T *b2, const T *b3,             \
          const T *w, const i64 *shape, int nterms, int mirror, double *sums)               \
{                                                                                           \
    const i64 n0 = shape[0], n1 = shape[1], n2 = shape[2], n3 = shape[3];                   \
    const i64 st[4] = { n1 * n2 * n3, n2 * n3, n3, 1 };                                     \
    ACC delta = 0, rnorm = 0;                                                               \
    const i64 rows = n0 * n1;                                                               \
    _Pragma("omp parallel for reduction(+:delta,rnorm) schedule(static)")                   \
    for (i64 ij = 0; ij < rows; ++ij) {                                                     \
        const i64 i = ij / n1, j = ij % n1;                                                 \
        const i64 f0 = (i + 1 == n0) ? (mirror ? 0 : -(n0 - 1) * st[0]) : st[0];            \
        const i64 f1 = (j + 1 == n1) ? (mirror ? 0 : -(n1 - 1) * st[1]) : st[1];            \
        for (i64 k = 0; k < n2; ++k) {                                                      \
            const i64 f2 = (k + 1 == n2) ? (mirror ? 0 : -(n2 - 1) * st[2]) : st[2];        \
            for (i64 l = 0; l < n3; ++l) {                                                  \
                const i64 f3 = (l + 1 == n3) ? (mirror ? 0 : -(n3 - 1)) : 1;                \
                const i64 x = ij * st[1] + k * n3 + l;                                      \
                const T old = u[x];                                                         \
                T s = (w[0] * (b0[x] - b0[x + f0])) + (w[1] * (b1[x] - b1[x + f1]));        \
                s = s + (w[2] * (b2[x] - b2[x + f2]));                                      \
                if (nterms == 4) s = s + (w[3] * (b3[x] - b3[x + f3]));                     \
                const T nu = f[x] - s;                                                      \
                u[x] = nu;                                                                  \
                delta += FABS(nu - old);                                                    \
                rnorm += FABS(old);                                                         \
            }                                                                               \
        }                                                                                   \
    }                                                                                       \
    sums[0] = (double)delta; sums[1] = (double)rnorm;                                       \
}

DEF_DCU(tvo_dcu_f32_T, float,  float,  fabs)
DEF_DCU(tvo_dcu_f32_D, float,  double, fabs)
DEF_DCU(tvo_dcu_f64_T, double, double, fabs)
DEF_DCU(tvo_dcu_f64_D, double, double, fabs)

/* Sigma (a-b)^2  (utils.pyx:14-30, :35-49) */
#define DEF_SSE(NAME, T, ACC)                                                               \
double NAME(const T *a, const T *b, i64 n)                                                  \
{                                                                                           \
    ACC s = 0;                                                                              \
    _Pragma("omp parallel for reduction(+:s) schedule(static)")                             \
    for (i64 x = 0; x < n; ++x) { T t = a[x] - b[x]; s += (t * t); }                        \
    return (double)s;                                                                       \
}
DEF_SSE(tvo_sse_f32_T, float,  float)
DEF_SSE(tvo_sse_f32_D, float,  double)
DEF_SSE(tvo_sse_f64_T, double, double)
DEF_SSE(tvo_sse_f64_D, double, double)
