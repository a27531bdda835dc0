// cytvdn_shard.cu -- the sharded iteration loop behind the C ABI (cytvdn_shard_*, cytvdn_denoise_sharded).
//
// Replaces the reference's MPI driver for the hot path: cyTVDN/mpi.py:130-210 (scan-axis partition with one overlap
// plane per neighbour), :265-294 (buffers) and :314-438 (iteration + per-half-step plane exchange).  One `Shard`
// object per GPU; the objects of one run live in one process (cytvdn_denoise_sharded: one host thread drives all
// devices) or in one process per GPU (torchrun: the arenas are mapped into the neighbours with CUDA IPC).
//
// Layout (1-D split of scan axis 0, the default on NVSwitch -- contiguous halo planes, two neighbours): every shard
// stores its owned planes plus one overlap plane towards each existing neighbour, like mpi.py:165-196, in ONE
// device allocation: header (flags, sums) | orig | state set 0 (recon, b x4, d x4) | state set 1.
//
// Schedule per iteration (the fused single-pass kernel, out of place between the two state sets):
//   compute stream : wait until the neighbours' planes of the previous iteration have landed (spin on two flags in
//                    this shard's header), sweep the halo planes first (first / last owned plane and the upper
//                    overlap plane), record an event, sweep the interior;
//   copy stream    : after that event, the COPY ENGINES push the new first owned recon plane into the lower
//                    neighbour's upper overlap plane and the new last owned plane into the upper neighbour's lower
//                    overlap plane (cudaMemcpyAsync through peer pointers over NVLink), each followed by a 4-byte
//                    copy that raises the neighbour's flag to the iteration number.
// The exchange therefore runs under the interior sweep and costs no SM (round 1 used ncclSend/ncclRecv inline on
// the compute stream: its SM-resident copy kernels and the rank skew they created cost 1.1 - 1.9 ms of a 28 ms
// iteration).  Only the reconstruction travels (one plane each way per iteration, the same volume as mpi.py's
// acc-right / recon-left pair): the forward neighbour b'[last owned + 1] is recomputed locally from the overlap
// plane's own state, and the corrected plane indices of SURVEY.md section 5.8 apply (sender's first / last OWNED
// plane, not mpi.py's overlap planes).  The fused kernel is told not to store recon outside the owned range
// (cytvdn_step_opts.flags bit 1), so overlap planes of recon are written by the neighbours' pushes only and no
// "ready to receive" handshake is needed: a push of iteration i can only be issued after this shard's halo planes of
// iteration i-1 were swept (the pusher waited for OUR flag of i-1, which we raise after that sweep), and those
// sweeps are the last readers of the plane being overwritten.
//
// Sums (sum|b|, sum|recon' - recon|, sum|recon|) are taken over owned voxels only per launch, added in
// a fixed order; the caller adds the shards' triples (3 doubles per iteration -- on the host in the single-process
// driver, with one all-reduce at the end under torchrun).
#include "../../include/cytvdn_b200.h"
#include "internal.hh"

#include <unistd.h>

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

using cytvdn_internal::fail;

namespace {

constexpr int kMaxBox = 16;               // boxes of scan axis 1 in the host-pipelined run
constexpr int kSlots = 3 * kMaxBox;       // launches per iteration that write sums (3 per box), 4 doubles each
constexpr size_t kAlign = 256;
constexpr int kProfileMax = 256;          // iterations with per-phase events (cytvdn_shard_timeline)
// header words: "planes of iteration < value have landed" counters, raised by the neighbours
//   [0] whole planes from the lower neighbour, [1] ... from the upper neighbour, [2] error,
//   [4 + c] box c from the lower neighbour, [4 + kMaxBox + c] box c from the upper neighbour
constexpr uint32_t kFlagLo = 0, kFlagHi = 1, kFlagErr = 2, kBoxLo = 4, kBoxHi = 4 + kMaxBox, kFlagWords = 4 + 2 * kMaxBox;

inline size_t up(size_t x) { return (x + kAlign - 1) & ~(kAlign - 1); }

// Spin until, for every box c in [c_lo, c_hi], the neighbour's planes of the iteration before `want` have landed:
// the whole-plane counter or the box's own counter has reached `want` (each raised by the neighbour's copy stream
// behind the data).  One warp, one lane per box; gives up after ~20 s of SM clock and records the failure in *err so
// that a dead neighbour cannot hang the GPU.
__global__ void wait_flags_kernel(const volatile uint32_t *whole, const volatile uint32_t *box, int c_lo, int c_hi, uint32_t want,
                                  uint32_t *err)
{
    const int c = c_lo + (int)threadIdx.x;
    if (c <= c_hi && *(volatile uint32_t *)err == 0) {        // (after one timeout every later wait gives up at once)
        const long long t0 = clock64();
        for (;;) {
            uint32_t v, w;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(whole) : "memory");
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(w) : "l"(box + c) : "memory");
            if ((int32_t)(v - want) >= 0 || (int32_t)(w - want) >= 0) break;
            if (clock64() - t0 > 40000000000ll) { atomicExch(err, want ? want : 1u); break; }
            __nanosleep(200);
        }
    }
    __threadfence_system();
}

// raise a counter in the NEIGHBOUR's header (peer pointer); runs on the copy stream behind the plane it announces
__global__ void set_flag_kernel(uint32_t *flag, uint32_t value)
{
    __threadfence_system();
    *(volatile uint32_t *)flag = value;
    __threadfence_system();
}

struct Peer {
    char *base = nullptr;         // the neighbour's arena as seen from this device
    bool ipc = false;             // opened with cudaIpcOpenMemHandle (must be closed)
    int rank = -1;
    int64_t n_local = 0;          // its number of stored planes
    int64_t pitch = 0;            // bytes between its arrays
    int64_t hdr = 0;              // bytes of its header
};

}  // namespace

struct cytvdn_shard {
    // ---- geometry ----
    int dtype = 0, world = 1, rank = 0, device = 0;
    size_t elem = 4;
    int64_t g[4] = {0, 0, 0, 0};
    bool periodic = false, mirror = false, has_lo = false, has_hi = false;
    int64_t valid_lo = 0, valid_hi = 0;       // owned global planes [lo, hi)
    int64_t read_lo = 0;                      // global index of local plane 0 (may be -1 -> wraps when periodic)
    int64_t n_local = 0, own_lo = 0, own_hi = 0;
    int64_t n3 = 0, n3p = 0;
    size_t plane_b = 0;                       // bytes of one stored plane (padded rows)
    bool fista = true;
    int max_iters = 0;
    double clip[4], w[4];
    // ---- memory ----
    char *arena = nullptr;
    size_t arena_bytes = 0, hdr = 0, pitch = 0;
    int n_arrays = 0;
    uint32_t *flags = nullptr;
    double *sums = nullptr;                   // [max_iters][kSlots][4]
    Peer lo, hi;
    // ---- execution ----
    cudaStream_t comp = nullptr, copy = nullptr, up = nullptr, down = nullptr;     // up / down: host copies of run_host
    cudaEvent_t ev_halo = nullptr;
    cudaEvent_t ev_pushed[kMaxBox + 1][2] = {};      // [box (kMaxBox: whole planes)][iteration parity]
    int64_t pushed_it[kMaxBox + 1][2];               // the iteration each was last recorded for (-1: never)
    cudaEvent_t ev_up[kMaxBox] = {}, ev_box_done[kMaxBox] = {}, ev_down = nullptr;
    int64_t it_global_base = 0;               // iterations enqueued since creation up to the last load (flags count
    int64_t it_run = 0;                       // base + it_run); it_run: iterations since the last load
    int64_t it_global() const { return it_global_base + it_run; }
    double tk = 1.0;
    bool profile = false;
    std::vector<cudaEvent_t> pev;             // kProfileMax x 6 events
    int64_t launches = 0;

    char *array(int idx) const { return arena + hdr + (size_t)idx * pitch; }
    // array indices: 0 orig | 1 + s*per: recon of set s, then b0..3, then d0..3
    int per() const { return 1 + 4 * (fista ? 2 : 1); }
    int idx_recon(int s) const { return 1 + s * per(); }
    int idx_b(int s, int k) const { return 1 + s * per() + 1 + k; }
    int idx_d(int s, int k) const { return 1 + s * per() + 5 + k; }
};

namespace {

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

struct ExportBlob {                // what cytvdn_shard_export hands to the neighbours (<= 128 bytes)
    unsigned char ipc[64];
    int64_t n_local, pitch, hdr;
    int32_t device, pid, rank, pad;
    uint64_t base;                 // arena address in the exporting process (valid for same-process peers)
};
static_assert(sizeof(ExportBlob) <= 128, "export blob must fit the 128-byte handle");

void plan_1d(cytvdn_shard *s)
{
    // mpi.py:161-170: n = ceil(N / w) planes per tile, the last tile takes what is left
    const int64_t n = (s->g[0] + s->world - 1) / s->world;
    s->valid_lo = s->rank * n;
    s->valid_hi = std::min<int64_t>((s->rank + 1) * n, s->g[0]);
    s->has_lo = s->rank > 0;
    s->has_hi = s->rank < s->world - 1;
    if (s->periodic && s->world > 1) s->has_lo = s->has_hi = true;
    s->read_lo = s->valid_lo - (s->has_lo ? 1 : 0);
    s->n_local = (s->valid_hi - s->valid_lo) + (s->has_lo ? 1 : 0) + (s->has_hi ? 1 : 0);
    s->own_lo = s->has_lo ? 1 : 0;
    s->own_hi = s->n_local - (s->has_hi ? 1 : 0);
}

// One step of the schedule: iteration `m_run` (counted from the last load) on the rows [j0, j1) of scan axis 1 --
// the whole array (box < 0) or box `box` of a host-pipelined run.  Halo planes first, then the pushes of their rows on
// the copy stream, then the interior planes.
int enqueue_step(cytvdn_shard *s, int64_t m_run, bool fista_it, double tkr, int box, int nbox, int64_t j0, int64_t j1, bool zero_wrap1)
{
    const int64_t it = s->it_global_base + m_run;             // iteration number since creation (flags count these)
    const int in = (int)(it & 1), out = in ^ 1;
    const bool first = m_run == 0;
    const bool whole = box < 0;
    const int eb = whole ? kMaxBox : box;                     // index into the push events
    const bool prof = s->profile && whole && m_run < kProfileMax;
    cudaEvent_t *pe = prof ? &s->pev[(size_t)m_run * 6] : nullptr;
    if (prof) CYTVDN_CUDA_TRY(cudaEventRecord(pe[0], s->comp));
    // ---- wait for the planes of the previous iteration (they went into the set this iteration reads); a box also
    //      reads one row of each neighbouring box on the overlap planes.  Also at the first iteration after a re-load:
    //      a neighbour that has raised our counter to `it` has swept its halo planes of iteration it-1, the last
    //      readers of what our first push will overwrite. ----
    if (it > 0 && (s->has_lo || s->has_hi)) {
        const int c_lo = whole ? 0 : std::max(0, box - 1), c_hi = whole ? nbox - 1 : std::min(nbox - 1, box + 1);
        if (s->has_lo)
            wait_flags_kernel<<<1, 32, 0, s->comp>>>(s->flags + kFlagLo, s->flags + kBoxLo, c_lo, c_hi, (uint32_t)it, s->flags + kFlagErr);
        if (s->has_hi)
            wait_flags_kernel<<<1, 32, 0, s->comp>>>(s->flags + kFlagHi, s->flags + kBoxHi, c_lo, c_hi, (uint32_t)it, s->flags + kFlagErr);
        CYTVDN_CUDA_TRY(cudaGetLastError());
    }
    // this step overwrites planes of the set that the copy engines read two iterations ago
    if (s->has_lo || s->has_hi)
        for (int q = 0; q <= kMaxBox; ++q)
            if (s->pushed_it[q][it & 1] == it - 2 && (whole || q == kMaxBox || q == box))
                CYTVDN_CUDA_TRY(cudaStreamWaitEvent(s->comp, s->ev_pushed[q][it & 1], 0));
    if (prof) CYTVDN_CUDA_TRY(cudaEventRecord(pe[1], s->comp));

    int64_t shape[4] = {s->n_local, s->g[1], s->g[2], s->g[3]};
    const void *bin[4], *din[4];
    void *bout[4], *dout[4];
    for (int k = 0; k < 4; ++k) {
        bin[k] = s->array(s->idx_b(in, k)); bout[k] = s->array(s->idx_b(out, k));
        din[k] = s->fista ? s->array(s->idx_d(in, k)) : nullptr;
        dout[k] = s->fista ? s->array(s->idx_d(out, k)) : nullptr;
    }
    const void *uin = first ? s->array(0) : s->array(s->idx_recon(in));       // recon = datacube.copy(), cyTVDN.py:145
    void *uout = s->array(s->idx_recon(out));
    double *sums = s->sums + (size_t)m_run * kSlots * 4 + (size_t)(whole ? 0 : box) * 3 * 4;

    cytvdn_step_opts o;
    memset(&o, 0, sizeof o);
    o.row_pitch = s->n3p;
    o.own_lo[0] = s->own_lo; o.own_hi[0] = s->own_hi;
    o.box_lo[1] = j0; o.box_hi[1] = j1;
    o.flags = 2;                                              // recon is stored for owned voxels only
    if (s->periodic && s->world > 1) o.flags |= 1 << 8;       // the wrap of axis 0 is the exchange's job
    if (!s->periodic && !s->mirror && s->has_lo && !s->has_hi) o.zero_wrap_mask = 1;   // global upper edge
    if (zero_wrap1) o.zero_wrap_mask |= 2;                    // last box of a pipelined run: row 0 is iterations ahead
    auto sweep = [&](int64_t lo, int64_t hi, int slot) -> int {
        if (hi <= lo) return CYTVDN_OK;
        cytvdn_step_opts ob = o;
        ob.box_lo[0] = lo; ob.box_hi[0] = hi;
        ++s->launches;
        return cytvdn_fused_iteration(4, shape, s->dtype, s->array(0), uin, uout, bin, bout, fista_it ? din : nullptr,
                                      fista_it ? dout : nullptr, tkr, s->clip, s->w, s->periodic ? 0 : (s->mirror ? 3 : 2), sums + slot * 4,
                                      &ob, s->comp);
    };
    if (whole || box == 0) CYTVDN_CUDA_TRY(cudaMemsetAsync(s->sums + (size_t)m_run * kSlots * 4, 0, sizeof(double) * kSlots * 4, s->comp));
    // ---- halo planes first: the planes that travel and the upper overlap plane; the lower overlap plane is never
    //      swept (nothing owned depends on its accumulators, its recon is received) ----
    const int64_t n = s->n_local;
    int64_t lo_end = s->own_lo;                               // interior = [lo_end, hi_begin)
    int64_t hi_begin = n;
    if (s->has_lo) lo_end = std::min<int64_t>(s->own_lo + 1, n);
    if (s->has_hi) hi_begin = std::max<int64_t>(lo_end, n - 2);
    if (int rc = sweep(s->own_lo, lo_end, 0)) return rc;      // first owned plane (goes to the lower neighbour)
    if (int rc = sweep(hi_begin, n, 1)) return rc;            // last owned plane (goes up) + upper overlap plane
    if (prof) CYTVDN_CUDA_TRY(cudaEventRecord(pe[2], s->comp));
    if (s->has_lo || s->has_hi) {
        CYTVDN_CUDA_TRY(cudaEventRecord(s->ev_halo, s->comp));
        CYTVDN_CUDA_TRY(cudaStreamWaitEvent(s->copy, s->ev_halo, 0));
        if (prof) CYTVDN_CUDA_TRY(cudaEventRecord(pe[4], s->copy));
        const char *mine = s->array(s->idx_recon(out));
        const size_t row_b = (size_t)s->g[2] * s->n3p * s->elem;                 // one row of scan axis 1 inside a plane
        const size_t off = (size_t)j0 * row_b, bytes = (size_t)(j1 - j0) * row_b;
        if (s->has_lo) {                                      // my first owned plane -> lower neighbour's LAST plane
            const Peer &p = s->lo;
            char *dst = p.base + p.hdr + (size_t)s->idx_recon(out) * p.pitch + (size_t)(p.n_local - 1) * s->plane_b + off;
            CYTVDN_CUDA_TRY(cudaMemcpyAsync(dst, mine + (size_t)s->own_lo * s->plane_b + off, bytes, cudaMemcpyDeviceToDevice, s->copy));
            set_flag_kernel<<<1, 1, 0, s->copy>>>((uint32_t *)p.base + (whole ? kFlagHi : kBoxHi + box), (uint32_t)(it + 1));
        }
        if (s->has_hi) {                                      // my last owned plane -> upper neighbour's plane 0
            const Peer &p = s->hi;
            char *dst = p.base + p.hdr + (size_t)s->idx_recon(out) * p.pitch + off;
            CYTVDN_CUDA_TRY(cudaMemcpyAsync(dst, mine + (size_t)(s->own_hi - 1) * s->plane_b + off, bytes, cudaMemcpyDeviceToDevice, s->copy));
            set_flag_kernel<<<1, 1, 0, s->copy>>>((uint32_t *)p.base + (whole ? kFlagLo : kBoxLo + box), (uint32_t)(it + 1));
        }
        CYTVDN_CUDA_TRY(cudaGetLastError());
        CYTVDN_CUDA_TRY(cudaEventRecord(s->ev_pushed[eb][it & 1], s->copy));
        s->pushed_it[eb][it & 1] = it;
        if (prof) CYTVDN_CUDA_TRY(cudaEventRecord(pe[5], s->copy));
    }
    if (int rc = sweep(lo_end, hi_begin, 2)) return rc;       // interior, under the exchange
    if (prof) CYTVDN_CUDA_TRY(cudaEventRecord(pe[3], s->comp));
    return CYTVDN_OK;
}

// the FISTA ratio of the next iteration (cyTVDN.py:154-156), host float64
double next_tk_ratio(cytvdn_shard *s)
{
    const double tk_new = (1.0 + std::sqrt(1.0 + 4.0 * s->tk * s->tk)) / 2.0;
    const double r = (s->tk - 1.0) / tk_new;
    s->tk = tk_new;
    return r;
}

int enqueue_iteration(cytvdn_shard *s, bool fista_it)
{
    if (s->it_run >= s->max_iters) return fail(CYTVDN_E_INVALID, "shard was created for %d iterations per load", s->max_iters);
    const double tkr = fista_it ? next_tk_ratio(s) : 0.0;
    if (int rc = enqueue_step(s, s->it_run, fista_it, tkr, -1, 1, 0, s->g[1], false)) return rc;
    ++s->it_run;
    return CYTVDN_OK;
}

}  // namespace

extern "C" {

int cytvdn_shard_create(const cytvdn_shard_params *p, cytvdn_shard **out)
{
    if (!p || !out) return fail(CYTVDN_E_INVALID, "NULL argument");
    *out = nullptr;
    if (p->dtype != CYTVDN_F32 && p->dtype != CYTVDN_F64) return fail(CYTVDN_E_INVALID, "bad dtype");
    for (int k = 0; k < 4; ++k)
        if (p->gshape[k] < 1 || p->gshape[k] > 0x7fffffff) return fail(CYTVDN_E_INVALID, "gshape[%d] out of range", k);
    if (p->world < 1 || p->rank < 0 || p->rank >= p->world) return fail(CYTVDN_E_INVALID, "bad world / rank");
    if (p->max_iters < 1 || p->max_iters > 60000) return fail(CYTVDN_E_INVALID, "max_iters must be in 1..60000");
    const int64_t n = (p->gshape[0] + p->world - 1) / p->world;
    if ((int64_t)p->rank * n >= p->gshape[0])
        return fail(CYTVDN_E_INVALID, "axis 0: extent %lld cannot be split over %d tiles of %lld planes (tile %d would be empty)",
                    (long long)p->gshape[0], p->world, (long long)n, p->rank);
    cytvdn_shard *s = new cytvdn_shard();
    s->dtype = p->dtype; s->elem = p->dtype == CYTVDN_F32 ? 4 : 8;
    s->world = p->world; s->rank = p->rank; s->periodic = p->periodic == 1; s->mirror = p->periodic == 2;
    for (int k = 0; k < 4; ++k) { s->g[k] = p->gshape[k]; s->clip[k] = p->clip[k]; s->w[k] = p->lambda_mu[k]; }
    s->fista = p->fista != 0; s->max_iters = p->max_iters;
    plan_1d(s);
    const int64_t vw = 16 / (int64_t)s->elem;
    s->n3 = s->g[3]; s->n3p = (s->n3 + vw - 1) / vw * vw;
    s->plane_b = (size_t)s->g[1] * s->g[2] * s->n3p * s->elem;
    int dev = p->device;
    if (dev < 0) CYTVDN_CUDA_TRY(cudaGetDevice(&dev));
    s->device = dev;
    DeviceGuard guard(dev);
    s->n_arrays = 1 + 2 * s->per();
    s->pitch = up((size_t)s->n_local * s->plane_b);
    const size_t flags_b = up(sizeof(uint32_t) * kFlagWords), sums_b = up(sizeof(double) * kSlots * 4 * s->max_iters);
    s->hdr = flags_b + sums_b;
    s->arena_bytes = s->hdr + (size_t)s->n_arrays * s->pitch;
    cudaError_t e = cudaMalloc((void **)&s->arena, s->arena_bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        const size_t need = s->arena_bytes;
        delete s;
        return fail(CYTVDN_E_NOMEM, "cudaMalloc of the shard arena (%zu bytes) failed: %s", need, cudaGetErrorString(e));
    }
    s->flags = (uint32_t *)s->arena;
    s->sums = (double *)(s->arena + flags_b);
    for (auto &q : s->pushed_it) q[0] = q[1] = -100;
    auto bail = [&](cudaError_t err, const char *what) {
        cudaGetLastError();
        cudaFree(s->arena);
        delete s;
        return fail(CYTVDN_E_CUDA, "%s failed: %s", what, cudaGetErrorString(err));
    };
    if ((e = cudaStreamCreateWithFlags(&s->comp, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    if ((e = cudaStreamCreateWithFlags(&s->copy, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    if ((e = cudaStreamCreateWithFlags(&s->up, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    if ((e = cudaStreamCreateWithFlags(&s->down, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    if ((e = cudaEventCreateWithFlags(&s->ev_halo, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = cudaEventCreateWithFlags(&s->ev_down, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
    for (auto &q : s->ev_pushed)
        for (auto &ev : q)
            if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
    for (int c = 0; c < kMaxBox; ++c) {
        if ((e = cudaEventCreateWithFlags(&s->ev_up[c], cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
        if ((e = cudaEventCreateWithFlags(&s->ev_box_done[c], cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
    }
    if ((e = cudaMemsetAsync(s->arena, 0, s->hdr, s->comp)) != cudaSuccess) return bail(e, "cudaMemsetAsync");
    // plane 0 of the accumulators is never swept on a shard with a lower neighbour: keep it defined in both sets
    if ((e = cudaStreamSynchronize(s->comp)) != cudaSuccess) return bail(e, "cudaStreamSynchronize");
    // the sweeps' reduction scratch now, not at the first launch: an allocation in the middle of a run may synchronise
    // the device while another shard's wait kernel spins on it (several shards of one process on one GPU)
    if (cytvdn_internal::warm_workspace(s->comp) != CYTVDN_OK) { cudaFree(s->arena); delete s; return CYTVDN_E_CUDA; }
    *out = s;
    return CYTVDN_OK;
}

int cytvdn_shard_disconnect(cytvdn_shard *s)
{
    if (!s) return CYTVDN_OK;
    DeviceGuard guard(s->device);
    for (cudaStream_t st : {s->comp, s->copy, s->up, s->down}) if (st) cudaStreamSynchronize(st);
    const bool same = s->lo.base && s->lo.base == s->hi.base;                     // one mapping, two roles
    if (s->lo.base && s->lo.ipc) cudaIpcCloseMemHandle(s->lo.base);
    if (s->hi.base && s->hi.ipc && !same) cudaIpcCloseMemHandle(s->hi.base);
    s->lo.base = s->hi.base = nullptr;
    cudaGetLastError();
    return CYTVDN_OK;
}

int cytvdn_shard_destroy(cytvdn_shard *s)
{
    if (!s) return CYTVDN_OK;
    cytvdn_shard_disconnect(s);
    DeviceGuard guard(s->device);
    for (auto e : s->pev) cudaEventDestroy(e);
    if (s->ev_halo) cudaEventDestroy(s->ev_halo);
    if (s->ev_down) cudaEventDestroy(s->ev_down);
    for (auto &q : s->ev_pushed) for (auto ev : q) if (ev) cudaEventDestroy(ev);
    for (auto ev : s->ev_up) if (ev) cudaEventDestroy(ev);
    for (auto ev : s->ev_box_done) if (ev) cudaEventDestroy(ev);
    if (s->comp) cudaStreamDestroy(s->comp);
    if (s->copy) cudaStreamDestroy(s->copy);
    if (s->up) cudaStreamDestroy(s->up);
    if (s->down) cudaStreamDestroy(s->down);
    if (s->arena) cudaFree(s->arena);
    cudaGetLastError();
    delete s;
    return CYTVDN_OK;
}

int cytvdn_shard_info(const cytvdn_shard *s, int64_t out[12])
{
    if (!s || !out) return fail(CYTVDN_E_INVALID, "NULL argument");
    out[0] = s->n_local; out[1] = s->own_lo; out[2] = s->own_hi; out[3] = s->valid_lo; out[4] = s->valid_hi;
    out[5] = s->read_lo; out[6] = s->has_lo; out[7] = s->has_hi; out[8] = (int64_t)s->arena_bytes; out[9] = s->n3p;
    out[10] = s->launches; out[11] = s->it_run;
    return CYTVDN_OK;
}

int cytvdn_shard_export(const cytvdn_shard *s, unsigned char handle[128])
{
    if (!s || !handle) return fail(CYTVDN_E_INVALID, "NULL argument");
    DeviceGuard guard(s->device);
    ExportBlob b;
    memset(&b, 0, sizeof b);
    cudaIpcMemHandle_t h;
    CYTVDN_CUDA_TRY(cudaIpcGetMemHandle(&h, s->arena));
    memcpy(b.ipc, &h, 64);
    b.n_local = s->n_local; b.pitch = (int64_t)s->pitch; b.hdr = (int64_t)s->hdr; b.device = s->device;
    b.pid = (int32_t)getpid();
    b.rank = s->rank;
    b.base = (uint64_t)(uintptr_t)s->arena;
    memset(handle, 0, 128);
    memcpy(handle, &b, sizeof b);
    return CYTVDN_OK;
}

int cytvdn_shard_connect(cytvdn_shard *s, int side, const unsigned char handle[128])
{
    if (!s || !handle || (side != 0 && side != 1)) return fail(CYTVDN_E_INVALID, "bad argument");
    if ((side == 0 && !s->has_lo) || (side == 1 && !s->has_hi)) return CYTVDN_OK;     // no neighbour there
    DeviceGuard guard(s->device);
    ExportBlob b;
    memcpy(&b, handle, sizeof b);
    Peer &p = side == 0 ? s->lo : s->hi;
    Peer &other = side == 0 ? s->hi : s->lo;
    p.n_local = b.n_local; p.pitch = b.pitch; p.hdr = b.hdr; p.rank = b.rank;
    if (b.pid == (int32_t)getpid()) {                          // same process: the pointer itself, peer access enabled
        p.base = (char *)(uintptr_t)b.base;
        p.ipc = false;
        if (b.device != s->device) {
            int can = 0;
            CYTVDN_CUDA_TRY(cudaDeviceCanAccessPeer(&can, s->device, b.device));
            if (!can) return fail(CYTVDN_E_UNSUPPORTED, "device %d cannot access device %d as a peer", s->device, b.device);
            cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                return fail(CYTVDN_E_CUDA, "cudaDeviceEnablePeerAccess(%d) failed: %s", b.device, cudaGetErrorString(e));
            cudaGetLastError();
        }
        return CYTVDN_OK;
    }
    if (other.base && other.ipc && other.rank == b.rank) {     // two shards on a periodic axis: both neighbours are one
        p.base = other.base; p.ipc = true;
        return CYTVDN_OK;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, b.ipc, 64);
    void *q = nullptr;
    CYTVDN_CUDA_TRY(cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess));
    p.base = (char *)q;
    p.ipc = true;
    return CYTVDN_OK;
}

int cytvdn_shard_array(const cytvdn_shard *s, int which, int set, int axis, void **ptr)
{
    if (!s || !ptr) return fail(CYTVDN_E_INVALID, "NULL argument");
    // `set` 0: the state the NEXT iteration reads (current), 1: the other one
    const int cur = (int)((s->it_global() & 1) ^ (set ? 1 : 0));
    if (which == 0) *ptr = s->array(0);
    else if (which == 1) *ptr = (s->it_run == 0 && !set) ? s->array(0) : s->array(s->idx_recon(cur));
    else if (which == 2 && axis >= 0 && axis < 4) *ptr = s->array(s->idx_b(cur, axis));
    else if (which == 3 && axis >= 0 && axis < 4 && s->fista) *ptr = s->array(s->idx_d(cur, axis));
    else return fail(CYTVDN_E_INVALID, "which must be 0 (orig), 1 (recon), 2 (b), 3 (d); axis 0..3");
    return CYTVDN_OK;
}

namespace {
int reset_state(cytvdn_shard *s);
}

int cytvdn_shard_load(cytvdn_shard *s, const void *block)
{
    if (!s) return fail(CYTVDN_E_INVALID, "NULL argument");
    if (s->has_lo && !s->lo.base) return fail(CYTVDN_E_INVALID, "lower neighbour not connected");
    if (s->has_hi && !s->hi.base) return fail(CYTVDN_E_INVALID, "upper neighbour not connected");
    DeviceGuard guard(s->device);
    // the previous run's iterations (and the pushes they issued) are complete on this shard
    for (cudaStream_t st : {s->comp, s->copy, s->up, s->down}) CYTVDN_CUDA_TRY(cudaStreamSynchronize(st));
    const size_t rows = (size_t)s->n_local * s->g[1] * s->g[2];
    if (block) {                                              // dense rows -> padded rows
        char *dst = s->array(0);
        if (s->n3p == s->n3) CYTVDN_CUDA_TRY(cudaMemcpyAsync(dst, block, rows * s->n3 * s->elem, cudaMemcpyDefault, s->comp));
        else {
            CYTVDN_CUDA_TRY(cudaMemsetAsync(dst, 0, rows * s->n3p * s->elem, s->comp));
            CYTVDN_CUDA_TRY(cudaMemcpy2DAsync(dst, (size_t)s->n3p * s->elem, block, (size_t)s->n3 * s->elem, (size_t)s->n3 * s->elem,
                                              rows, cudaMemcpyDefault, s->comp));
        }
    }
    return reset_state(s);
}

namespace {
int reset_state(cytvdn_shard *s)
{
    // b = d = 0 in the set the first iteration reads; plane 0 of the other set too (never swept with a lower
    // neighbour, so it must not hold garbage -- ADVICE round 1)
    const int in = (int)(s->it_global() & 1);
    for (int k = 0; k < 4; ++k) {
        CYTVDN_CUDA_TRY(cudaMemsetAsync(s->array(s->idx_b(in, k)), 0, (size_t)s->n_local * s->plane_b, s->comp));
        CYTVDN_CUDA_TRY(cudaMemsetAsync(s->array(s->idx_b(in ^ 1, k)), 0, s->plane_b, s->comp));
        if (s->fista) {
            CYTVDN_CUDA_TRY(cudaMemsetAsync(s->array(s->idx_d(in, k)), 0, (size_t)s->n_local * s->plane_b, s->comp));
            CYTVDN_CUDA_TRY(cudaMemsetAsync(s->array(s->idx_d(in ^ 1, k)), 0, s->plane_b, s->comp));
        }
    }
    s->it_global_base += s->it_run;           // the counters the neighbours raise keep counting across loads
    s->it_run = 0;
    s->tk = 1.0;
    return CYTVDN_OK;
}
}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// Host-pipelined run (one process per GPU, or any single shard): load from a HOST block, iterate, store the owned planes
// to a HOST block, with the PCIe copies overlapped with the iterations -- the single-GPU wavefront pipeline of
// cytvdn_denoise carried into the sharded loop.  Boxes are cut along scan axis 1, the axis that is NOT sharded, so the
// wavefront over (box, iteration) runs in lockstep on all ranks and does not couple them; per box and iteration the
// box's rows of the halo planes travel, announced by the box's own counter.  Box c starts iterating when boxes c and
// c+1 have arrived and is copied back when ITS last iteration is done.  Needs a non-periodic boundary (box 0 would
// depend on the last box), dense rows (a 2-D copy per box) and at least 4 rows of axis 1 per box.
// ---------------------------------------------------------------------------------------------------------------
int cytvdn_shard_run_host(cytvdn_shard *s, const void *block, void *owned_out, int n_fista, int n_plain)
{
    if (!s || !block || !owned_out || n_fista < 0 || n_plain < 0) return fail(CYTVDN_E_INVALID, "bad argument");
    if (n_fista > 0 && !s->fista) return fail(CYTVDN_E_INVALID, "shard was created without FISTA auxiliaries");
    const int n_it = n_fista + n_plain;
    if (n_it > s->max_iters) return fail(CYTVDN_E_INVALID, "shard was created for %d iterations per load", s->max_iters);
    if (s->has_lo && !s->lo.base) return fail(CYTVDN_E_INVALID, "lower neighbour not connected");
    if (s->has_hi && !s->hi.base) return fail(CYTVDN_E_INVALID, "upper neighbour not connected");
    int nbox = (int)std::min<int64_t>(kMaxBox, s->g[1] / 4);
    { const char *env = getenv("CYTVDN_SHARD_PIPELINE"); if (env && *env) nbox = std::min(nbox, atoi(env)); }
    if (s->periodic || s->n3p != s->n3 || nbox < 2 || n_it < 1) {     // not pipelined: plain load / iterate / store
        if (int rc = cytvdn_shard_load(s, block)) return rc;
        if (int rc = cytvdn_shard_iterate(s, n_fista, n_plain)) return rc;
        return cytvdn_shard_store(s, owned_out);
    }
    DeviceGuard guard(s->device);
    for (cudaStream_t st : {s->comp, s->copy, s->up, s->down}) CYTVDN_CUDA_TRY(cudaStreamSynchronize(st));
    if (int rc = reset_state(s)) return rc;                   // (memsets on the compute stream, before the first sweep)
    auto jlo = [&](int c) -> int64_t { return s->g[1] * c / nbox; };
    const size_t row_b = (size_t)s->g[2] * s->n3 * s->elem;   // one row of scan axis 1 (dense == padded here)
    const size_t dense_plane = (size_t)s->g[1] * row_b;
    // ---- uploads: box by box, every stored plane's rows [j0, j1) in one 2-D copy ----
    for (int c = 0; c < nbox; ++c) {
        const size_t off = (size_t)jlo(c) * row_b, width = (size_t)(jlo(c + 1) - jlo(c)) * row_b;
        CYTVDN_CUDA_TRY(cudaMemcpy2DAsync(s->array(0) + off, s->plane_b, (const char *)block + off, dense_plane, width,
                                          (size_t)s->n_local, cudaMemcpyDefault, s->up));
        CYTVDN_CUDA_TRY(cudaEventRecord(s->ev_up[c], s->up));
    }
    // ---- iterations: wavefront over (box, iteration) at both ends, whole-array sweeps in between ----
    int64_t count = 0;
    if (int rc = cytvdn_pipeline_schedule(nbox, n_it, nullptr, nullptr, 0, &count)) return rc;
    std::vector<int32_t> obox((size_t)count), oit((size_t)count);
    if (int rc = cytvdn_pipeline_schedule(nbox, n_it, obox.data(), oit.data(), count, &count)) return rc;
    std::vector<double> tkr((size_t)n_it, 0.0);
    for (int m = 0; m < n_fista; ++m) tkr[m] = next_tk_ratio(s);
    const int cur_final = (int)((s->it_global_base + n_it) & 1);
    for (int64_t q = 0; q < count; ++q) {
        const int c = obox[q], m = oit[q];
        if (c >= 0 && m == 0) CYTVDN_CUDA_TRY(cudaStreamWaitEvent(s->comp, s->ev_up[std::min(c + 1, nbox - 1)], 0));
        if (c < 0 && m == 0) CYTVDN_CUDA_TRY(cudaStreamWaitEvent(s->comp, s->ev_up[nbox - 1], 0));
        const bool last_box = c == nbox - 1;
        if (int rc = enqueue_step(s, m, m < n_fista, tkr[m], c, nbox, c < 0 ? 0 : jlo(c), c < 0 ? s->g[1] : jlo(c + 1),
                                  last_box && !s->mirror))
            return rc;
        if (m == n_it - 1) {                                  // this box (or everything) is final: copy its owned planes home
            const int c0 = c < 0 ? 0 : c, c1 = c < 0 ? nbox : c + 1;
            const size_t off = (size_t)jlo(c0) * row_b, width = (size_t)(jlo(c1) - jlo(c0)) * row_b;
            CYTVDN_CUDA_TRY(cudaEventRecord(s->ev_box_done[c0], s->comp));
            CYTVDN_CUDA_TRY(cudaStreamWaitEvent(s->down, s->ev_box_done[c0], 0));
            const char *src = s->array(s->idx_recon(cur_final)) + (size_t)s->own_lo * s->plane_b + off;
            CYTVDN_CUDA_TRY(cudaMemcpy2DAsync((char *)owned_out + off, dense_plane, src, s->plane_b, width,
                                              (size_t)(s->own_hi - s->own_lo), cudaMemcpyDefault, s->down));
        }
    }
    s->it_run = n_it;
    CYTVDN_CUDA_TRY(cudaEventRecord(s->ev_down, s->down));
    CYTVDN_CUDA_TRY(cudaStreamWaitEvent(s->comp, s->ev_down, 0));     // "everything this run enqueued" ends on comp
    return CYTVDN_OK;
}

int cytvdn_shard_iterate(cytvdn_shard *s, int n_fista, int n_plain)
{
    if (!s || n_fista < 0 || n_plain < 0) return fail(CYTVDN_E_INVALID, "bad argument");
    if (n_fista > 0 && !s->fista) return fail(CYTVDN_E_INVALID, "shard was created without FISTA auxiliaries");
    DeviceGuard guard(s->device);
    for (int i = 0; i < n_fista; ++i) if (int rc = enqueue_iteration(s, true)) return rc;
    for (int i = 0; i < n_plain; ++i) if (int rc = enqueue_iteration(s, false)) return rc;
    return CYTVDN_OK;
}

int cytvdn_shard_synchronize(cytvdn_shard *s)
{
    if (!s) return fail(CYTVDN_E_INVALID, "NULL argument");
    DeviceGuard guard(s->device);
    for (cudaStream_t st : {s->comp, s->copy, s->up, s->down}) CYTVDN_CUDA_TRY(cudaStreamSynchronize(st));
    uint32_t err = 0;
    CYTVDN_CUDA_TRY(cudaMemcpy(&err, s->flags + kFlagErr, sizeof err, cudaMemcpyDeviceToHost));
    if (err) return fail(CYTVDN_E_CUDA, "halo exchange timed out on rank %d waiting for iteration %u of a neighbour", s->rank, err);
    return CYTVDN_OK;
}

int cytvdn_shard_streams(const cytvdn_shard *s, void **compute_stream, void **copy_stream)
{
    if (!s) return fail(CYTVDN_E_INVALID, "NULL argument");
    if (compute_stream) *compute_stream = s->comp;
    if (copy_stream) *copy_stream = s->copy;
    return CYTVDN_OK;
}

int cytvdn_shard_sums(cytvdn_shard *s, double *out, int n)
{
    if (!s || !out || n < 0 || n > s->it_run) return fail(CYTVDN_E_INVALID, "bad argument (n must be <= iterations run)");
    if (int rc = cytvdn_shard_synchronize(s)) return rc;
    DeviceGuard guard(s->device);
    std::vector<double> h((size_t)n * kSlots * 4);
    if (n) CYTVDN_CUDA_TRY(cudaMemcpy(h.data(), s->sums, h.size() * sizeof(double), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; ++i)
        for (int q = 0; q < 3; ++q) {
            double x = 0.0;
            for (int sl = 0; sl < kSlots; ++sl) x += h[((size_t)i * kSlots + sl) * 4 + q];     // fixed order
            out[(size_t)i * 3 + q] = x;
        }
    return CYTVDN_OK;
}

int cytvdn_shard_store(cytvdn_shard *s, void *owned_block)
{
    if (!s || !owned_block) return fail(CYTVDN_E_INVALID, "NULL argument");
    DeviceGuard guard(s->device);
    const int cur = (int)(s->it_global() & 1);
    const char *src = (s->it_run == 0 ? s->array(0) : s->array(s->idx_recon(cur))) + (size_t)s->own_lo * s->plane_b;
    const size_t rows = (size_t)(s->own_hi - s->own_lo) * s->g[1] * s->g[2];
    if (s->n3p == s->n3) CYTVDN_CUDA_TRY(cudaMemcpyAsync(owned_block, src, rows * s->n3 * s->elem, cudaMemcpyDefault, s->comp));
    else
        CYTVDN_CUDA_TRY(cudaMemcpy2DAsync(owned_block, (size_t)s->n3 * s->elem, src, (size_t)s->n3p * s->elem, (size_t)s->n3 * s->elem,
                                          rows, cudaMemcpyDefault, s->comp));
    CYTVDN_CUDA_TRY(cudaStreamSynchronize(s->comp));
    return CYTVDN_OK;
}

int cytvdn_shard_profile(cytvdn_shard *s, int on)
{
    if (!s) return fail(CYTVDN_E_INVALID, "NULL argument");
    DeviceGuard guard(s->device);
    if (on && s->pev.empty()) {
        s->pev.resize((size_t)kProfileMax * 6);
        for (auto &e : s->pev) CYTVDN_CUDA_TRY(cudaEventCreate(&e));
    }
    s->profile = on != 0;
    return CYTVDN_OK;
}

int cytvdn_shard_timeline(cytvdn_shard *s, double *out, int n)
{
    if (!s || !out || n < 0) return fail(CYTVDN_E_INVALID, "bad argument");
    if (s->pev.empty()) return fail(CYTVDN_E_INVALID, "profiling was not enabled (cytvdn_shard_profile)");
    if (int rc = cytvdn_shard_synchronize(s)) return rc;
    DeviceGuard guard(s->device);
    const int m = (int)std::min<int64_t>(std::min<int64_t>(n, s->it_run), kProfileMax);
    const bool nb = s->has_lo || s->has_hi;
    for (int i = 0; i < m; ++i) {
        cudaEvent_t *pe = &s->pev[(size_t)i * 6];
        float t = 0;
        double *o = out + (size_t)i * 6;
        // [0] start of the iteration relative to iteration 0's start, [1] wait for the neighbours' planes,
        // [2] halo planes, [3] interior, [4] halo-done -> push start (copy stream latency), [5] push (copy engines)
        CYTVDN_CUDA_TRY(cudaEventElapsedTime(&t, s->pev[0], pe[0])); o[0] = t;
        CYTVDN_CUDA_TRY(cudaEventElapsedTime(&t, pe[0], pe[1])); o[1] = t;
        CYTVDN_CUDA_TRY(cudaEventElapsedTime(&t, pe[1], pe[2])); o[2] = t;
        CYTVDN_CUDA_TRY(cudaEventElapsedTime(&t, pe[2], pe[3])); o[3] = t;
        o[4] = o[5] = 0.0;
        if (nb) {
            CYTVDN_CUDA_TRY(cudaEventElapsedTime(&t, pe[2], pe[4])); o[4] = t;
            CYTVDN_CUDA_TRY(cudaEventElapsedTime(&t, pe[4], pe[5])); o[5] = t;
        }
    }
    return CYTVDN_OK;
}

// -------------------------------------------------------------------------------------------------------------
// The whole sharded loop in ONE process: one host thread drives `ndev` devices (cyTVDN/mpi.py:26-438 without MPI).
// -------------------------------------------------------------------------------------------------------------
int cytvdn_denoise_sharded(const cytvdn_denoise_params *p, int ndev, const int *devices, const void *data, void *recon,
                           double *bnorm, double *delta, int32_t *iters_done, double *timing_ms)
{
    if (!p || !data || !recon) return fail(CYTVDN_E_INVALID, "NULL argument");
    if (ndev < 1 || ndev > 64) return fail(CYTVDN_E_INVALID, "ndev must be in 1..64");
    if (p->ndim != 4) return fail(CYTVDN_E_UNSUPPORTED, "sharding exists for 4-D datacubes only (mpi.py:252-255)");
    const int nF = p->iters_fista, nU = p->iters_plain, nIt = nF + nU;
    if (nF < 0 || nU < 0) return fail(CYTVDN_E_INVALID, "negative iteration count");
    if (nIt > 0 && (!bnorm || !delta)) return fail(CYTVDN_E_INVALID, "bnorm / delta is NULL");
    // out of core: the tiles run the two-pass kernels, which also know the half-isotropic pairs
    if (p->schedule == 3) return cytvdn_denoise_sharded_streamed(p, ndev, devices, data, recon, bnorm, delta, iters_done, timing_ms);
    if (p->isotropic_R || p->isotropic_Q)
        return fail(CYTVDN_E_UNSUPPORTED, "the in-core sharded loop is anisotropic (mpi.py:317-358); schedule 3 (out of core) "
                                          "runs half-isotropic pairs on several devices");
    if (p->bc_mode != 0 && p->bc_mode != 2 && p->bc_mode != 3)
        return fail(CYTVDN_E_UNSUPPORTED, "sharded runs support BC_mode 2 (mpi.py:84), 0 and 3");
    const auto t_start = std::chrono::steady_clock::now();
    auto ms_since = [&](std::chrono::steady_clock::time_point t0) {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    };
    const size_t elem = p->dtype == CYTVDN_F32 ? 4 : 8;
    auto is_dev = [](const void *q) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, q) != cudaSuccess) { cudaGetLastError(); return false; }
        return at.type == cudaMemoryTypeDevice;
    };
    std::vector<cytvdn_shard *> sh(ndev, nullptr);
    struct Cleanup {
        std::vector<cytvdn_shard *> &v;
        ~Cleanup()
        {
            // nobody may free an arena a neighbour still pushes into
            for (auto s : v) if (s) { DeviceGuard g(s->device); cudaStreamSynchronize(s->comp); cudaStreamSynchronize(s->copy); }
            for (auto s : v) cytvdn_shard_destroy(s);
        }
    } cleanup{sh};
    cytvdn_shard_params sp;
    memset(&sp, 0, sizeof sp);
    sp.dtype = p->dtype; sp.world = ndev; sp.periodic = p->bc_mode == 0 ? 1 : (p->bc_mode == 3 ? 2 : 0);
    sp.fista = nF > 0; sp.max_iters = std::max(1, nIt);
    for (int k = 0; k < 4; ++k) { sp.gshape[k] = p->shape[k]; sp.clip[k] = p->clip[k]; sp.lambda_mu[k] = p->lambda_mu[k]; }
    std::vector<std::array<unsigned char, 128>> handles(ndev);
    for (int r = 0; r < ndev; ++r) {
        sp.rank = r; sp.device = devices ? devices[r] : r;
        if (int rc = cytvdn_shard_create(&sp, &sh[r])) {
            // the shards do not fit the devices' memory: schedule 0 falls back to the out-of-core tiles
            if (rc == CYTVDN_E_NOMEM && p->schedule == 0 && !p->use_stopping && p->bc_mode != 0 && nIt > 0 && !is_dev(data) && !is_dev(recon)) {
                for (auto &q : sh) { cytvdn_shard_destroy(q); q = nullptr; }
                return cytvdn_denoise_sharded_streamed(p, ndev, devices, data, recon, bnorm, delta, iters_done, timing_ms);
            }
            return rc;
        }
        if (int rc = cytvdn_shard_export(sh[r], handles[r].data())) return rc;
    }
    for (int r = 0; r < ndev; ++r) {
        if (int rc = cytvdn_shard_connect(sh[r], 0, handles[(r + ndev - 1) % ndev].data())) return rc;
        if (int rc = cytvdn_shard_connect(sh[r], 1, handles[(r + 1) % ndev].data())) return rc;
    }
    // ---- load: every shard copies its stored planes (owned + overlap) of the global array ----
    const size_t gplane = (size_t)p->shape[1] * p->shape[2] * p->shape[3] * elem;      // dense plane of the caller's array
    for (int r = 0; r < ndev; ++r) {
        cytvdn_shard *s = sh[r];
        if (s->read_lo >= 0 && s->read_lo + s->n_local <= s->g[0]) {
            if (int rc = cytvdn_shard_load(s, (const char *)data + (size_t)s->read_lo * gplane)) return rc;
        } else {                                              // periodic wrap: the overlap planes come from the other end
            DeviceGuard g(s->device);
            if (int rc = cytvdn_shard_load(s, nullptr)) return rc;
            for (int64_t i = 0; i < s->n_local; ++i) {
                const int64_t gi = ((s->read_lo + i) % s->g[0] + s->g[0]) % s->g[0];
                char *dst = s->array(0) + (size_t)i * s->plane_b;
                const char *src = (const char *)data + (size_t)gi * gplane;
                const size_t rows = (size_t)s->g[1] * s->g[2];
                if (s->n3p == s->n3) CYTVDN_CUDA_TRY(cudaMemcpyAsync(dst, src, gplane, cudaMemcpyDefault, s->comp));
                else {
                    CYTVDN_CUDA_TRY(cudaMemsetAsync(dst, 0, s->plane_b, s->comp));
                    CYTVDN_CUDA_TRY(cudaMemcpy2DAsync(dst, (size_t)s->n3p * elem, src, (size_t)s->n3 * elem, (size_t)s->n3 * elem, rows,
                                                      cudaMemcpyDefault, s->comp));
                }
            }
        }
    }
    const double setup_ms = ms_since(t_start);
    const auto t_loop = std::chrono::steady_clock::now();
    // ---- iterate: iteration by iteration across the devices (every wait finds its push already enqueued) ----
    int done[2] = {0, 0};
    std::vector<double> tot((size_t)std::max(1, nIt) * 3, 0.0), part((size_t)std::max(1, nIt) * 3);
    for (int phase = 0; phase < 2; ++phase) {
        const int n = phase == 0 ? nF : nU;
        for (int it = 0; it < n; ++it) {
            for (int r = 0; r < ndev; ++r)
                if (int rc = cytvdn_shard_iterate(sh[r], phase == 0 ? 1 : 0, phase == 0 ? 0 : 1)) return rc;
            ++done[phase];
            if (p->use_stopping) {                            // cyTVDN.py:189-194: needs the global delta now
                const int i = done[0] + done[1] - 1;
                double s3[3] = {0, 0, 0};
                for (int r = 0; r < ndev; ++r) {
                    if (int rc = cytvdn_shard_sums(sh[r], part.data(), i + 1)) return rc;
                    for (int q = 0; q < 3; ++q) s3[q] += part[(size_t)i * 3 + q];
                }
                const double dl = s3[1] / s3[2];
                const double dl_t = p->dtype == CYTVDN_F32 ? (double)(float)dl : dl;
                if (dl_t < p->stopping_relative_change) break;
            }
        }
    }
    const int ran = done[0] + done[1];
    for (int r = 0; r < ndev; ++r) {
        if (int rc = cytvdn_shard_sums(sh[r], part.data(), ran)) return rc;
        for (int i = 0; i < ran * 3; ++i) tot[i] += part[i];                                   // rank order: deterministic
    }
    const double loop_ms = ms_since(t_loop);
    const auto t_fin = std::chrono::steady_clock::now();
    for (int i = 0; i < nIt; ++i) {
        bnorm[i] = i < ran ? tot[(size_t)i * 3] : 0.0;
        delta[i] = i < ran ? tot[(size_t)i * 3 + 1] / tot[(size_t)i * 3 + 2] : 0.0;
    }
    // ---- result: owned planes go home (the copies of all devices run concurrently) ----
    for (int r = 0; r < ndev; ++r) {
        cytvdn_shard *s = sh[r];
        DeviceGuard g(s->device);
        const int cur = (int)(s->it_global() & 1);
        const char *src = (s->it_run == 0 ? s->array(0) : s->array(s->idx_recon(cur))) + (size_t)s->own_lo * s->plane_b;
        char *dst = (char *)recon + (size_t)s->valid_lo * gplane;
        const size_t rows = (size_t)(s->own_hi - s->own_lo) * s->g[1] * s->g[2];
        if (s->n3p == s->n3) CYTVDN_CUDA_TRY(cudaMemcpyAsync(dst, src, rows * s->n3 * elem, cudaMemcpyDefault, s->comp));
        else
            CYTVDN_CUDA_TRY(cudaMemcpy2DAsync(dst, (size_t)s->n3 * elem, src, (size_t)s->n3p * elem, (size_t)s->n3 * elem, rows,
                                              cudaMemcpyDefault, s->comp));
    }
    for (int r = 0; r < ndev; ++r) if (int rc = cytvdn_shard_synchronize(sh[r])) return rc;
    if (iters_done) { iters_done[0] = done[0]; iters_done[1] = done[1]; iters_done[2] = 2 | (ndev << 8); }
    if (timing_ms) { timing_ms[0] = setup_ms; timing_ms[1] = loop_ms; timing_ms[2] = ms_since(t_fin); }
    return CYTVDN_OK;
}

}  // extern "C"
