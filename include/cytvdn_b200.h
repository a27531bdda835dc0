/*
 * cytvdn_b200.h -- C ABI of libcytvdn_b200.so: the TV-denoising hot path of cyTVDN
 * (tv.denoise3D / tv.denoise4D) as hand-written CUDA for NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary.  The reference has no C ABI: its
 * native interface is three CPython extension modules taking typed memoryviews.  Each entry
 * point below names the reference function it replaces (file:line under /root/reference).
 * Plain pointers and sizes only; no torch / numpy types.  All functions return 0 on success
 * or a CYTVDN_E_* code; cytvdn_last_error() gives the message of the calling thread's last
 * failure.  Nothing here ever falls back to the CPU.
 *
 * Conventions
 *  - Arrays are C-contiguous, dtype CYTVDN_F32 or CYTVDN_F64, ndim 3 or 4, `shape` has ndim
 *    entries.  All arrays of one call share dtype and shape (the reference raises
 *    "Buffer dtype mismatch" otherwise).
 *  - Step-level functions take DEVICE pointers and are stream-ordered (asynchronous); their
 *    reductions are written as doubles to DEVICE memory (`*_dev`), never accumulated.
 *    Scalars that the reference casts to the array dtype at the call (clip, tk, lambda_mu)
 *    are passed as double and rounded to the array dtype inside, exactly like the reference.
 *  - bc_mode: 0 periodic, 1 mirror, 2 Jia-Zhao (anisotropic.pyx:20-24).  The reference's mirror is
 *    only defined for the accumulator update (backward neighbour of index 0 is index 1,
 *    anisotropic.pyx:69-70); its reconstruction update with bc_mode 1 reads out of bounds
 *    (utils.pyx:117-120 take the forward index as max(i+1, N-1)) and is rejected here.
 *    bc_mode 3 (CYTVDN_BC_MIRROR, NOT in the reference) is the well-defined reading of those lines:
 *    accumulator update as bc_mode 1, reconstruction update with the forward index CLAMPED,
 *    min(i+1, N-1), i.e. the term of an axis vanishes at its last index.  Anisotropic only, every extent >= 2;
 *    all schedules (two-pass, fused, PCIe pipeline, out of core, sharded).
 *  - Per-voxel arithmetic is bit-identical to the reference (separately rounded mul/add,
 *    IEEE division, comparison-based clip); only the reductions differ: they are
 *    accumulated in float64 (the reference's array-dtype sums are inaccurate).
 */
#ifndef CYTVDN_B200_H
#define CYTVDN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CYTVDN_VERSION 100          /* 0.1.0 */

#define CYTVDN_F32 0
#define CYTVDN_F64 1

#define CYTVDN_BC_PERIODIC 0
#define CYTVDN_BC_MIRROR_A 1        /* the reference's mirror: accumulator update only */
#define CYTVDN_BC_JIA_ZHAO 2
#define CYTVDN_BC_MIRROR   3        /* clamped-index mirror, defined for both half-steps (extension) */

#define CYTVDN_OK            0
#define CYTVDN_E_INVALID     1      /* bad argument (message says which) */
#define CYTVDN_E_CUDA        2      /* CUDA runtime error */
#define CYTVDN_E_NOMEM       3      /* device allocation failed */
#define CYTVDN_E_UNSUPPORTED 4      /* e.g. bc_mode 1 in the reconstruction update */

int         cytvdn_version(void);
const char *cytvdn_last_error(void);
/* number of CUDA devices; 0 (and CYTVDN_OK) on a machine without a GPU */
int         cytvdn_device_count(int *count);

/*
 * Optional extras for the step functions (NULL = whole array, reference semantics).
 * They exist for the scan-axis sharding that replaces cyTVDN/mpi.py:155-210,314-438:
 *  box_*   : half-open range of axis-0 / axis-1 indices swept by this launch (lets the
 *            caller launch halo planes first and the interior on another stream);
 *            box_hi[k] <= 0 means "up to the extent".
 *  own_*   : range whose voxels enter the reductions (overlap planes excluded);
 *            own_hi[k] <= 0 means "up to the extent".
 *  zero_wrap_mask (reconstruction update only): bit k set -> the forward neighbour of the
 *            last index on axis k is taken as 0 instead of reading index 0 (tile on the
 *            global upper edge that holds a received plane at index 0).
 *  flags bit 1 : (fused iteration only) store recon_out only for voxels inside the owned range; overlap planes of
 *            recon are then written by the halo exchange alone (cytvdn_shard_*).
 *  flags bits 4, 5 : (fused iteration only) half-isotropic update of the pair (0,1) / (2,3), i.e. isotropic_R /
 *            isotropic_Q of cyTVDN.py:159-180; 4-D arrays, rows 16-byte aligned, Jia-Zhao on the pair's axes.
 *  flags bit 0 : hand the tiles of the sweep out dynamically (one global counter) instead of a static
 *            stride per CTA.  Use it for a sweep that overlaps with other GPU work (the NCCL halo exchange):
 *            the CTAs that are resident then share all tiles.  ~2 % slower when the kernel runs alone.
 *  flags bits 8..11 : axis k (bit 8+k) uses the Jia-Zhao boundary in the accumulator update whatever
 *            bc_mode says.  For periodic runs that are split on axis k: the wrap there is done by the halo
 *            exchange, the block's own index 0 on that axis is an overlap plane.
 *  row_pitch : all arrays of the call store their rows (fast axis) `row_pitch` elements apart instead of
 *            densely; the pad voxels are ignored (never read as neighbours of real voxels, excluded from the
 *            reductions; they are overwritten with unspecified values).  cytvdn_denoise pads rows to the
 *            vector width internally so that odd row lengths run on the 16-byte path.
 *  l2_budget_bytes : working-set budget that sizes the axis-1 strips of the sweep
 *            (0 = library default).
 */
typedef struct cytvdn_step_opts {
    int64_t box_lo[2];
    int64_t box_hi[2];
    int64_t own_lo[2];
    int64_t own_hi[2];
    int32_t zero_wrap_mask;
    int32_t flags;              /* bit 0 dynamic tiles, 1 owned-only recon stores, 4/5 iso pairs, 8..11 per-axis Jia-Zhao */
    int64_t l2_budget_bytes;
    int64_t row_pitch;          /* elements between consecutive rows of the fast axis; 0 = dense (= extent) */
    /* cytvdn_fused_iteration only: axis-0 halo read from neighbouring GPUs through peer pointers
       (cytvdn_ipc_open); NULL = no neighbour on that side.  peer_lo_recon points at the START OF THE LAST
       PLANE of the lower neighbour's recon_in; the peer_hi_* point at the upper neighbour's recon_in / b_in[0] /
       d_in[0] (their first plane is used).  Same inner shape and row pitch as the local arrays. */
    const void *peer_lo_recon;
    const void *peer_hi_recon;
    const void *peer_hi_b0;
    const void *peer_hi_d0;
    /* cytvdn_fused_iteration (anisotropic) and cytvdn_datacube_update: device array of the arrays' shape and row
       pitch; when set, sum (sse_reference - recon_out)^2 over the owned voxels is written behind the other sums
       (sums_dev[3] / sums_dev[2]) -- sum_square_error_* (utils.pyx:14-49) of cyTVDN.py:186-187 folded into the pass. */
    const void *sse_reference;
} cytvdn_step_opts;

/*
 * Half-step A, one axis.  Replaces
 *   accumulator_update_4D        anisotropic.pyx:17-84     (ndim 4, d == NULL)
 *   accumulator_update_4D_FISTA  anisotropic.pyx:89-164    (ndim 4, d != NULL)
 *   accumulator_update_3D        anisotropic.pyx:169-237   (ndim 3, d == NULL)
 *   accumulator_update_3D_FISTA  anisotropic.pyx:243-317   (ndim 3, d != NULL)
 * b (and d) are updated in place; *norm_dev = sum |b_new| (their return value).
 */
int cytvdn_accumulator_update(int ndim, const int64_t *shape, int dtype,
                              const void *a, void *b, void *d, double tk, int ax, double clip,
                              int bc_mode, double *norm_dev, const cytvdn_step_opts *opts,
                              void *stream);

/*
 * Half-step A for an axis pair with joint 2-norm shrink (4-D, Jia-Zhao boundary only).  Replaces
 *   iso_accumulator_update_4D        halfisotropic.pyx:17-97     (d1 == d2 == NULL)
 *   iso_accumulator_update_4D_FISTA  halfisotropic.pyx:102-188
 * with race-free semantics (the reference shares scratch between OpenMP threads).
 */
int cytvdn_iso_accumulator_update(const int64_t *shape, int dtype, const void *a,
                                  void *b1, void *b2, void *d1, void *d2, double tk,
                                  int ax1, int ax2, double clip, double *norm_dev,
                                  const cytvdn_step_opts *opts, void *stream);

/*
 * Half-step A for ALL axes in one pass over `a` (what denoise3D/4D use).  Equivalent to the
 * ndim calls of cyTVDN.py:159-180 (4-D; :378-386 3-D): b[k]/d[k]/clip[k] belong to axis k;
 * iso_R / iso_Q select the half-isotropic update for the pairs (0,1) / (2,3) with
 * clip[0] / clip[2] as in cyTVDN.py:160-162,171-173.  d == NULL -> unaccelerated.
 * *norm_dev = sum over all axes of sum |b_new| (= b_norm[i] of the reference).
 */
int cytvdn_accumulator_update_all(int ndim, const int64_t *shape, int dtype, const void *a,
                                  void *const *b, void *const *d, double tk,
                                  const double *clip, int iso_R, int iso_Q, int bc_mode,
                                  double *norm_dev, const cytvdn_step_opts *opts, void *stream);

/*
 * Half-step B.  Replaces datacube_update_4D utils.pyx:54-125 and datacube_update_3D
 * utils.pyx:131-199 (bc_mode 0 and 2 share one path there, :85/:163; bc_mode 3: see the top of this file).
 * recon_out[x] = orig[x] - sum_k lambda_mu[k] * (b[k][x] - b[k][x + e_k mod N_k]),
 * summed left to right.  recon_in may equal recon_out (the reference updates in place).
 * sums_dev[0] = sum |recon_out - recon_in|, sums_dev[1] = sum |recon_in|; the reference
 * returns their ratio.
 */
int cytvdn_datacube_update(int ndim, const int64_t *shape, int dtype, const void *orig,
                           const void *recon_in, void *recon_out, const void *const *b,
                           const double *lambda_mu, int bc_mode, double *sums_dev,
                           const cytvdn_step_opts *opts, void *stream);

/*
 * One WHOLE iteration in a single pass: half-step A for all axes fused with half-step B
 * (= cytvdn_accumulator_update_all followed by cytvdn_datacube_update, bit-identical results, i.e. the
 * loop body cyTVDN.py:159-184 / :378-390).  Each array crosses HBM once: 76 B/voxel instead of 96
 * (4-D FISTA fp32).  OUT OF PLACE: the new state goes to recon_out / b_out / d_out, which must not
 * alias the inputs (forward neighbours are recomputed from the old state).  Anisotropic, or -- opts->flags bits
 * 4 / 5 -- half-isotropic pairs (halfisotropic.pyx:63-95, :146-186); bc_mode 0, 2 or (anisotropic, rows 16-byte
 * aligned) 3.  d_in == d_out == NULL -> unaccelerated.
 * sums_dev[0] = sum |b_new| over all axes, [1] = sum |recon_out - recon_in|, [2] = sum |recon_in|.
 */
int cytvdn_fused_iteration(int ndim, const int64_t *shape, int dtype, const void *orig,
                           const void *recon_in, void *recon_out, const void *const *b_in,
                           void *const *b_out, const void *const *d_in, void *const *d_out,
                           double tk, const double *clip, const double *lambda_mu, int bc_mode,
                           double *sums_dev, const cytvdn_step_opts *opts, void *stream);

/* sum (a-b)^2 over n elements.  Replaces sum_square_error_4D utils.pyx:14-30, _3D :35-49. */
int cytvdn_sum_square_error(int64_t n, int dtype, const void *a, const void *b,
                            double *sse_dev, void *stream);

/*
 * The whole iteration loop, device resident.  Replaces the host loops of
 * denoise4D cyTVDN.py:127-247 and denoise3D cyTVDN.py:350-435.
 */
typedef struct cytvdn_denoise_params {
    int32_t ndim;               /* 3 or 4 */
    int32_t dtype;              /* CYTVDN_F32 / CYTVDN_F64 */
    int64_t shape[4];
    int32_t iters_fista;        /* FISTA iterations first ... */
    int32_t iters_plain;        /* ... then unaccelerated ones (cyTVDN.py:98-108) */
    int32_t isotropic_R;        /* 4-D only */
    int32_t isotropic_Q;        /* 4-D only */
    int32_t bc_mode;            /* 0, 2 or 3 (3: anisotropic only) */
    int32_t use_stopping;       /* stop a phase when delta < stopping_relative_change */
    double  stopping_relative_change;
    double  clip[4];            /* lambdaInv = 1/lam  (cyTVDN.py:77) */
    double  lambda_mu[4];       /* lam/mu             (cyTVDN.py:78) */
    int32_t device;             /* device to run on when `data` is a host pointer; -1 = current */
    int32_t schedule;           /* 0 auto, 1 two passes per iteration (in place), 2 fused single pass
                                   (out of place, second set of b/d arrays; not for bc_mode 3), 3 out of core
                                   (host arrays, tiles streamed over PCIe; see cytvdn_denoise) */
    void   *stream;             /* NULL = default stream */
} cytvdn_denoise_params;

/*
 * data, recon, reference_data: host or device pointers (detected); data is never modified;
 * recon receives the result.  bnorm / delta: HOST arrays of iters_fista + iters_plain doubles
 * (entries of iterations that did not run stay 0, like the reference's trailing zeros);
 * mse: HOST array of that length + 1, or NULL when reference_data is NULL.
 * iters_done (3 ints): [0], [1] FISTA / unaccelerated iterations actually executed, [2] low byte: the
 * schedule that ran (1 two-pass, 2 fused); bits 8..: number of boxes of the PCIe pipeline, 0 = not pipelined.
 * timing_ms (may be NULL): [0] allocation + host->device, [1] iteration loop (CUDA events),
 * [2] device->host + free.  In a pipelined run the copies overlap the loop and are part of [1].
 * The call is synchronous.
 *
 * PCIe pipeline (host `data` and/or host `recon`, fused schedule, bc_mode 2, no stopping test, no
 * reference_data, arrays >= 256 MB): the array is cut into 16 boxes of axis-0 planes; the first and last 16
 * iterations run box by box in wavefront order, so a box starts iterating when it (and the next one) has
 * arrived and is copied back when ITS last iteration is done.  The reconstruction is bit-identical to the
 * unpipelined run for finite data; bnorm/delta are summed box by box in double (fixed order).  The wrap term of
 * the last plane of axis 0 -- identically 0 under the Jia-Zhao boundary -- is taken as 0 instead of being
 * recomputed from plane 0, so a NaN/Inf in plane 0 does not reach the last plane as it would in the reference.
 * Environment CYTVDN_PIPELINE=0 disables it, =N (N >= 2) forces N boxes whatever the size.
 *
 * CYTVDN_TRACE=1 prints host-clock milestones of the call on stderr.
 *
 * Out of core (schedule 3; chosen by schedule 0 when `data` and `recon` are host arrays and not even the in-place
 * schedule fits in the free device memory): temporal blocking with overlapped tiles of axis-0 planes.  A pass
 * advances all voxels by K iterations: a tile (core + K halo planes per side) is copied in, iterated K times on a
 * box shrinking by one plane per iteration and side, its core copied back; between passes recon lives in the
 * caller's `recon`, b and d in pinned host arrays owned by the call (2 ndim arrays of the input's size with
 * FISTA).  Two tile slots, copy-in of tile t+1 under the iterations of tile t, the planes two tiles share handed
 * over on the device (carry buffer) so that every plane crosses the bus once per pass; the tiles are iterated with the
 * in-place two-pass kernels (fewest arrays per slot = most iterations per pass; the schedule is PCIe bound).
 * Needs bc_mode 2, no stopping test, no reference_data (half-isotropic is fine); reconstruction bit-identical to
 * the in-core schedules.  iters_done[2] = 3 | (tiles << 8); timing_ms from the host clock.  CYTVDN_STREAM_BUDGET_MB=N
 * forces this schedule with N MB of device memory (testing / benchmarking).
 */
int cytvdn_denoise(const cytvdn_denoise_params *params, const void *data, void *recon,
                   const void *reference_data, double *bnorm, double *delta, double *mse,
                   int32_t *iters_done, double *timing_ms);

/* Device working set of cytvdn_denoise in bytes (GPU analogue of check_memory, cyTVDN.py:438): every array the
   call allocates for the schedule `params` asks for (0 = auto counts the fused schedule when it applies), padding
   and reduction scratch included -- the number to hand to cytvdn_workspace_reserve. */
int cytvdn_denoise_workspace_bytes(const cytvdn_denoise_params *params, int data_on_device,
                                   int recon_on_device, int64_t *bytes);

/*
 * Optional reservation of the device working set (no counterpart in the reference, whose state lives in NumPy
 * arrays the driver allocates per call, cyTVDN.py:127-145).  cytvdn_denoise allocates its state (up to 19 arrays,
 * ~86 GB for BASELINE config 3) at entry and frees it at exit; on B200 that pair costs 60 ms to 0.7 s of host time.
 * A caller that denoises repeatedly reserves once: the library then holds ONE block of `bytes` on the current
 * device and every later call on that device whose state fits carves it up instead of calling cudaMalloc (calls
 * that do not fit allocate as before).  reserve(bytes) grows an existing block if needed (never shrinks);
 * release() frees it together with the per-stream reduction scratch.  Both fail with CYTVDN_E_INVALID while a call
 * is using the block.  info: size of the block and how much of it is handed out right now.
 */
int cytvdn_workspace_reserve(int64_t bytes);
int cytvdn_workspace_release(void);
int cytvdn_workspace_info(int64_t *reserved_bytes, int64_t *in_use_bytes);

/* Host-clock milestones (ms since entry) of the calling thread's last cytvdn_denoise and their labels (static
   strings): where the wall time outside the CUDA events went.  *count = number recorded (<= 12). */
int cytvdn_last_trace(double *ms, const char **what, int capacity, int *count);

/*
 * Host-side plans of the two PCIe schedules of cytvdn_denoise (pure functions, no GPU needed; the call runs exactly
 * what they return -- exported so that the schedules can be checked without a device).
 *
 * cytvdn_pipeline_schedule: launch order of the PCIe pipeline for `n_iter` iterations over `nbox` boxes of axis-0
 * planes: pairs (box, iteration), box -1 = one sweep of the whole array.  *count receives the number of launches;
 * box / iter (capacity entries each) may be NULL to query the count.  Invariant: when (c, m) runs, (c-1, m-1),
 * (c, m-1) and (c+1, m-1) have run, and (c-1, m+1), (c, m+1), (c+1, m+1) have not (iteration m reads state set
 * m%2 of the three boxes and writes set (m+1)%2 of box c).
 *
 * cytvdn_stream_plan: tile geometry of the out-of-core schedule for a device budget (two slots of P planes + a
 * carry buffer of 2K planes of every array): out8 = { planes per tile slot P, iterations per pass K, core planes per tile (P - 2K, or the whole axis when it fits), tiles, passes,
 * arrays per slot, bytes of one axis-0 plane of an internal array, bytes of page-locked host state the call
 * allocates }.  Tile t of a pass of Kp iterations holds planes [max(0, t*core - Kp), min(N0, (t+1)*core + Kp));
 * iteration k of the pass sweeps that range shrunk by k+1 planes on every side that is not an end of the array.
 */
int cytvdn_pipeline_schedule(int nbox, int n_iter, int32_t *box, int32_t *iter, int64_t capacity,
                             int64_t *count);
int cytvdn_stream_plan(const cytvdn_denoise_params *params, int64_t budget_bytes, int64_t *out8);
/* the same for cytvdn_denoise_sharded_streamed: budget per device (two slots + carry + the K-plane edge snapshot =
   2.75 P planes of every array), tiles a multiple of `ndev` of equal size; device r takes tiles
   [r * ceil(tiles / ndev), (r + 1) * ceil(tiles / ndev)) of every pass.  Per pass: every device copies in its first tile
   and the K planes above its last tile, BARRIER, then streams its tiles, BARRIER. */
int cytvdn_stream_plan_sharded(const cytvdn_denoise_params *params, int64_t budget_bytes_per_device, int ndev, int64_t *out8);

/*
 * ---------------------------------------------------------------------------------------------------------------
 * Sharded loop: scan-axis shards over several GPUs.  Replaces the hot path of the reference's MPI driver,
 * cyTVDN/mpi.py:130-210 (partition: tiles of ceil(N/w) planes with one overlap plane per neighbour, :161-196),
 * :265-294 (exchange buffers) and :314-438 (iteration with a plane exchange after each half-step).
 *
 * 1-D split of scan axis 0 (contiguous halo planes, two neighbours; the reference's 2-D (wx, wy) grid stays
 * available in the Python layer over NCCL), fused single-pass schedule, 4-D anisotropic, BC_mode 2 (mpi.py:84) or --
 * `periodic` -- 0 with the wrap done by the exchange.  One cytvdn_shard per GPU.  Per iteration a shard sweeps its
 * halo planes first, then its COPY ENGINES push the new first / last owned recon plane into the neighbours' overlap
 * planes through peer pointers (same process: cudaDeviceEnablePeerAccess; one process per GPU: CUDA IPC) while the
 * interior is swept; a 4-byte copy behind each plane raises a flag the neighbour's next iteration waits for.  No
 * NCCL, no SM is spent on the exchange, no host synchronisation inside the loop.  Deviations from mpi.py that make
 * the sharded result equal the single-GPU one bit for bit: the first / last OWNED planes
 * travel (mpi.py:325,344,408,414 send the overlap planes), a shard on the global upper edge takes its wrap term
 * as 0, FISTA is supported (mpi.py:310-311 is not), and bnorm / delta exist: sums over OWNED voxels.
 *
 * Use (every rank, or a loop over the ranks in one process):
 *   create -> export -> [exchange the 128-byte handles] -> connect(0, lower's handle), connect(1, upper's handle)
 *   -> load(block) -> iterate(nF, nU) [-> iterate ...] -> sums / store -> [all ranks synchronised] -> disconnect ->
 *   [barrier] -> destroy.
 * All ranks must enqueue the same iterations.  A shard is driven by one host thread at a time (different shards may be
 * driven from different threads).  A shard may be re-loaded and re-run any number of times; all ranks
 * must have finished (cytvdn_shard_synchronize + a barrier of the caller's) before any of them is destroyed.
 * ---------------------------------------------------------------------------------------------------------------
 */
typedef struct cytvdn_shard cytvdn_shard;
typedef struct cytvdn_shard_params {
    int32_t dtype;              /* CYTVDN_F32 / CYTVDN_F64 */
    int32_t world, rank;        /* tiles along scan axis 0 (mpi.py:130-150 with wy = 1) / this tile (mpi.py:156) */
    int32_t periodic;           /* 0: Jia-Zhao (BC_mode 2), 1: periodic (BC_mode 0, first and last tile are neighbours),
                                   2: the clamped mirror (BC_mode 3) at the global edges */
    int32_t fista;              /* allocate the FISTA auxiliaries d */
    int32_t max_iters;          /* iterations per load the shard keeps sums for */
    int32_t device;             /* CUDA device, -1 = current */
    int32_t reserved;
    int64_t gshape[4];          /* GLOBAL array shape */
    double  clip[4];            /* lambdaInv (mpi.py:249) */
    double  lambda_mu[4];       /* lam / mu  (mpi.py:250) */
} cytvdn_shard_params;

int cytvdn_shard_create(const cytvdn_shard_params *params, cytvdn_shard **shard);
/* Tear-down, one process per GPU: an arena that a neighbour still has mapped (CUDA IPC) is not given back to the device
   by cudaFree until that neighbour unmaps it.  So: all ranks synchronise, barrier, every rank DISCONNECTS (unmaps its
   neighbours' arenas), barrier, every rank DESTROYS.  destroy alone also disconnects, which is enough in one process. */
int cytvdn_shard_disconnect(cytvdn_shard *shard);
int cytvdn_shard_destroy(cytvdn_shard *shard);
/* out12 = { stored planes, first owned local plane, one past the last owned local plane, owned global range lo, hi,
   global index of local plane 0 (-1 / wraps on a periodic axis), has lower neighbour, has upper neighbour, arena
   bytes, row pitch in elements, kernels launched so far, iterations since the last load } */
int cytvdn_shard_info(const cytvdn_shard *shard, int64_t out12[12]);
/* 128-byte handle of this shard's device arena for its neighbours (CUDA IPC handle + layout; usable in the same
   process too, where it resolves to the pointer itself) */
int cytvdn_shard_export(const cytvdn_shard *shard, unsigned char handle[128]);
/* side 0: the lower neighbour (rank-1), 1: the upper neighbour (rank+1); a no-op where there is no neighbour */
int cytvdn_shard_connect(cytvdn_shard *shard, int side, const unsigned char handle[128]);
/* block: the stored planes (owned + overlap: global planes [out12[5], out12[5] + out12[0])) with DENSE rows, host or
   device pointer; NULL = the caller has written the shard's own `orig` array (cytvdn_shard_array) in place.
   Resets accumulators, auxiliaries and the FISTA schedule (recon = datacube.copy(), mpi.py:247).  Asynchronous. */
int cytvdn_shard_load(cytvdn_shard *shard, const void *block);
/* enqueue n_fista FISTA iterations, then n_plain unaccelerated ones (returns at once) */
int cytvdn_shard_iterate(cytvdn_shard *shard, int n_fista, int n_plain);
/* load + iterate + store in one call, HOST blocks in and out, with the PCIe copies overlapped with the iterations: the
   wavefront pipeline of cytvdn_denoise with boxes cut along scan axis 1 (the axis that is not sharded); per box and
   iteration the box's rows of the halo planes travel.  block: stored planes (as for load), owned_out: owned planes,
   both dense and preferably page-locked.  Asynchronous (cytvdn_shard_synchronize / _sums wait for it).  Falls back to
   load / iterate / store for periodic runs, padded rows or fewer than 8 rows of axis 1.  Bit-identical results. */
int cytvdn_shard_run_host(cytvdn_shard *shard, const void *block, void *owned_out, int n_fista, int n_plain);
/* wait for everything this shard has enqueued; reports a halo wait that timed out (dead neighbour) */
int cytvdn_shard_synchronize(cytvdn_shard *shard);
/* out[i*3 + {0,1,2}] = this shard's sum|b|, sum|recon' - recon|, sum|recon| over OWNED voxels of iteration i < n
   (synchronises) */
int cytvdn_shard_sums(cytvdn_shard *shard, double *out, int n);
/* the owned planes of the current reconstruction, dense rows, to a host or device pointer (synchronous) */
int cytvdn_shard_store(cytvdn_shard *shard, void *owned_block);
/* device pointer of an internal array (rows `row pitch` apart): which 0 orig, 1 recon, 2 b[axis], 3 d[axis];
   set 0 = the state the next iteration reads, 1 = the other state set */
int cytvdn_shard_array(const cytvdn_shard *shard, int which, int set, int axis, void **ptr);
int cytvdn_shard_streams(const cytvdn_shard *shard, void **compute_stream, void **copy_stream);
/* per-phase CUDA events for the first 256 iterations after a load; timeline: out[i*6 + k] in ms, k = 0 start of
   iteration i since iteration 0, 1 wait for the neighbours' planes, 2 halo planes, 3 interior sweep, 4 halo done ->
   first push issued (copy-stream latency), 5 pushes (copy engines; overlaps 3) */
int cytvdn_shard_profile(cytvdn_shard *shard, int on);
int cytvdn_shard_timeline(cytvdn_shard *shard, double *out, int n);

/*
 * The whole sharded run from ONE process: `ndev` devices (devices[r], or 0..ndev-1 when NULL) each take a tile of
 * scan axis 0 of the HOST (or managed) arrays `data` -> `recon`; same parameters and outputs as cytvdn_denoise
 * (4-D, anisotropic, BC_mode 2 or 0, schedule 0/2 in core, 3 = out of core: see cytvdn_denoise_sharded_streamed).
 * Bit-identical to cytvdn_denoise on one GPU.  iters_done[2] = schedule | ndev << 8.
 */
int cytvdn_denoise_sharded(const cytvdn_denoise_params *params, int ndev, const int *devices, const void *data,
                           void *recon, double *bnorm, double *delta, int32_t *iters_done, double *timing_ms);
/*
 * Sharded AND out of core: arrays whose state does not fit the GPUs' memory together (BASELINE config 5 on 2 or 4
 * GPUs; README.md:104-120 of the reference: "data larger than memory").  The out-of-core schedule of
 * cytvdn_denoise (temporal blocking: tiles of axis-0 planes + K halo planes are iterated K times per pass, state
 * between passes in page-locked host arrays) with the tiles of every pass dealt to `ndev` devices, one host thread
 * per device.  The K halo planes a device's first / last tile needs from the neighbouring device's range are read
 * from the host arrays before any device writes its results of the pass back (a barrier per pass instead of a halo
 * exchange per iteration).  Host arrays in and out, BC_mode 2, fixed iteration counts.  Bit-identical to the
 * in-core schedules.  budget_bytes_per_device: 0 = what is free on each device.
 */
int cytvdn_denoise_sharded_streamed(const cytvdn_denoise_params *params, int ndev, const int *devices, const void *data,
                                    void *recon, double *bnorm, double *delta, int32_t *iters_done, double *timing_ms);

/*
 * Deterministic synthetic 4D-STEM-like counts for benchmarks and sharded parity runs
 * (SURVEY.md section 8d): out[x] = max(0, rint(c + sqrt(c) * z(x))) with
 * c = counts * scan_mod[i,j] * templ[k,l] + 0.02 * counts and z an approximately normal
 * deviate derived from a 64-bit hash of (seed, GLOBAL linear index), so any sharding of the
 * global array yields the same values.  `out` is the local block starting at axis-0 index
 * `offset0` of a global array with shape `gshape`; `lshape0` planes are written.
 * scan_mod: device float[gshape0*gshape1]; templ: device float[gshape2*gshape3].
 */
int cytvdn_synth_counts(const int64_t *gshape, int64_t offset0, int64_t lshape0, int dtype,
                        const float *scan_mod_dev, const float *templ_dev, double counts,
                        uint64_t seed, void *out_dev, void *stream);

/* Small helpers so that a ctypes host needs no other CUDA binding. */
int cytvdn_malloc(void **ptr, int64_t bytes);
int cytvdn_free(void *ptr);
/* pinned host memory: huge-page mapping faulted in from all cores + cudaHostRegister (about 7x faster than
   cudaMallocHost for multi-GB buffers), cudaMallocHost as fallback or with CYTVDN_HOST_ALLOC=cuda */
int cytvdn_host_alloc(void **ptr, int64_t bytes);
int cytvdn_host_free(void *ptr);
int cytvdn_memcpy(void *dst, const void *src, int64_t bytes, void *stream);   /* any direction */
int cytvdn_memset(void *dst, int value, int64_t bytes, void *stream);
int cytvdn_stream_synchronize(void *stream);
int cytvdn_set_device(int device);
int cytvdn_get_device(int *device);
int cytvdn_mem_info(int64_t *free_bytes, int64_t *total_bytes);
/*
 * Peer access to another process's device allocation on the same node (one process per GPU): the owner
 * exports a 64-byte handle of an allocation made with cytvdn_malloc, the peer opens it and gets a pointer that
 * kernels on ITS device can dereference over NVLink.  Used by the sharded path to read halo planes straight
 * from the neighbour's HBM inside the fused sweep (no exchange step).
 */
int cytvdn_ipc_get_handle(void *ptr, unsigned char handle[64]);
int cytvdn_ipc_open(const unsigned char handle[64], void **peer_ptr);
int cytvdn_ipc_close(void *peer_ptr);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
int64_t cytvdn_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* CYTVDN_B200_H */
