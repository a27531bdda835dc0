// cytvdn_b200.cu -- C ABI of libcytvdn_b200.so (see include/cytvdn_b200.h) and the host side of
// the launches: argument validation, sweep tiling, per-stream reduction workspaces and the
// device-resident iteration loop that replaces the reference's Python host loops
// (cyTVDN/cyTVDN.py:127-247, :350-435).
#include "../../include/cytvdn_b200.h"
#include "kernels.cuh"
#include "fused.cuh"
#include "fused_tma.cuh"
#include "internal.hh"

#include <sys/mman.h>

#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <type_traits>
#include <utility>
#include <vector>

using namespace cytvdn;

namespace {

thread_local std::string g_err;
std::atomic<int64_t> g_launches{0};

int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                        \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
            return fail(_e == cudaErrorMemoryAllocation ? CYTVDN_E_NOMEM : CYTVDN_E_CUDA,     \
                        "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

// ---- per (device, stream) reduction workspace ----------------------------------------------
// One small scratch block (per-CTA partial sums + ticket) per (device, stream) pair that has launched a reducing
// kernel: launches on one stream are serialised, launches on different streams may overlap and must not share the
// ticket.  The table is bounded: beyond kMaxWorkspaces entries the least recently used one is freed (cudaFree
// synchronises the device, so a workspace is never freed under a running kernel), and
// cytvdn_workspace_release() drops all of them.  A stream handle the runtime recycles for a new stream simply
// finds the old stream's block (the ticket is zero between launches), which is harmless.
struct Workspace {
    double *partials = nullptr;   // [kMaxGrid][4]
    unsigned *ticket = nullptr;
    uint64_t stamp = 0;           // last use (LRU)
};
constexpr int kMaxGrid = 148 * 16;
constexpr size_t kMaxWorkspaces = 64;
std::mutex g_ws_mutex;
std::map<std::pair<int, cudaStream_t>, Workspace> g_ws;
uint64_t g_ws_clock = 0;

void free_workspace(Workspace &w)
{
    if (w.partials) cudaFree(w.partials);
    if (w.ticket) cudaFree(w.ticket);
    w.partials = nullptr; w.ticket = nullptr;
}

int get_workspace(cudaStream_t st, Workspace *out)
{
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_ws_mutex);
    auto key = std::make_pair(dev, st);
    auto it = g_ws.find(key);
    if (it == g_ws.end()) {
        if (g_ws.size() >= kMaxWorkspaces) {            // evict the least recently used block of THIS device
            auto victim = g_ws.end();
            for (auto j = g_ws.begin(); j != g_ws.end(); ++j)
                if (j->first.first == dev && (victim == g_ws.end() || j->second.stamp < victim->second.stamp)) victim = j;
            if (victim != g_ws.end()) { free_workspace(victim->second); g_ws.erase(victim); }
        }
        Workspace w;
        CUDA_TRY(cudaMalloc(&w.partials, sizeof(double) * 4 * kMaxGrid));
        CUDA_TRY(cudaMalloc(&w.ticket, sizeof(unsigned) * 4));
        CUDA_TRY(cudaMemsetAsync(w.ticket, 0, sizeof(unsigned) * 4, st));
        it = g_ws.emplace(key, w).first;
    }
    it->second.stamp = ++g_ws_clock;
    *out = it->second;
    return CYTVDN_OK;
}

// drop every reduction workspace of device `dev` (the caller has synchronised it)
void drop_workspaces(int dev)
{
    std::lock_guard<std::mutex> lk(g_ws_mutex);
    for (auto it = g_ws.begin(); it != g_ws.end();) {
        if (it->first.first == dev) { free_workspace(it->second); it = g_ws.erase(it); }
        else ++it;
    }
}

// ---- device properties / occupancy cache ------------------------------------------------------
struct DevInfo {
    int sms = 0;
    size_t l2 = 0;
};
int dev_info(DevInfo *d)
{
    static std::mutex m;
    static std::map<int, DevInfo> cache;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(m);
    auto it = cache.find(dev);
    if (it == cache.end()) {
        DevInfo di;
        int v = 0;
        CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
        di.sms = v;
        CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, dev));
        di.l2 = (size_t)v;
        it = cache.emplace(dev, di).first;
    }
    *d = it->second;
    return CYTVDN_OK;
}

template <typename K>
int grid_for(K kernel, int ntiles, int *grid, size_t dyn_smem = 0)
{
    static std::mutex m;
    static std::map<std::pair<int, const void *>, int> cache;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    DevInfo di;
    if (int rc = dev_info(&di)) return rc;
    int per_sm = 0;
    {
        std::lock_guard<std::mutex> lk(m);
        auto key = std::make_pair(dev, (const void *)kernel);
        auto it = cache.find(key);
        if (it == cache.end()) {
            if (dyn_smem > 0)
                CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem));
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kBlock, dyn_smem));
            if (per_sm < 1) per_sm = 1;
            static const char *env = getenv("CYTVDN_CTAS_PER_SM");
            if (env && atoi(env) > 0 && atoi(env) < per_sm) per_sm = atoi(env);
            cache.emplace(key, per_sm);
        } else {
            per_sm = it->second;
        }
    }
    int g = di.sms * per_sm;
    if (g > kMaxGrid) g = kMaxGrid;
    if (g > ntiles) g = ntiles;
    if (g < 1) g = 1;
    *grid = g;
    return CYTVDN_OK;
}

// ---- geometry ----------------------------------------------------------------------------------
struct Dims {
    int ndim;
    int64_t n[4];     // embedded 4-D extents ([N0,N1,1,N2] for 3-D)
    int axmap[4];     // user axis -> embedded axis
    int64_t pitch;    // elements between rows of the fast axis (>= n[3])
};

int make_dims(int ndim, const int64_t *shape, Dims *D, const cytvdn_step_opts *opts = nullptr)
{
    if (ndim != 3 && ndim != 4) return fail(CYTVDN_E_INVALID, "ndim must be 3 or 4 (got %d)", ndim);
    if (!shape) return fail(CYTVDN_E_INVALID, "shape is NULL");
    for (int k = 0; k < ndim; ++k)
        if (shape[k] < 1 || shape[k] > 0x7fffffff)
            return fail(CYTVDN_E_INVALID, "shape[%d]=%lld out of range", k, (long long)shape[k]);
    D->ndim = ndim;
    if (ndim == 4) {
        for (int k = 0; k < 4; ++k) { D->n[k] = shape[k]; D->axmap[k] = k; }
    } else {
        D->n[0] = shape[0]; D->n[1] = shape[1]; D->n[2] = 1; D->n[3] = shape[2];
        D->axmap[0] = 0; D->axmap[1] = 1; D->axmap[2] = 3; D->axmap[3] = -1;
    }
    D->pitch = D->n[3];
    if (opts && opts->row_pitch > 0) {
        if (opts->row_pitch < D->n[3] || opts->row_pitch > 0x7fffffff)
            return fail(CYTVDN_E_INVALID, "row_pitch %lld is smaller than the row length %lld", (long long)opts->row_pitch,
                        (long long)D->n[3]);
        D->pitch = opts->row_pitch;
    }
    return CYTVDN_OK;
}

// bytes of L2 the strips of one sweep may occupy
int64_t l2_budget(const cytvdn_step_opts *o)
{
    if (o && o->l2_budget_bytes > 0) return o->l2_budget_bytes;
    const char *env = getenv("CYTVDN_L2_BUDGET_MB");      // read per call: tests shrink it to force strips
    if (env && atof(env) > 0) return (int64_t)(atof(env) * 1048576.0);
    return 24ll << 20;
}

// footprint_arrays: number of distinct arrays whose lines live in L2 during the sweep
int make_sweep(const Dims &D, int vw, size_t elem, const cytvdn_step_opts *o, int footprint_arrays, Sweep *S)
{
    memset(S, 0, sizeof *S);
    S->n0 = (int32_t)D.n[0]; S->n1 = (int32_t)D.n[1]; S->n2 = (int32_t)D.n[2]; S->n3 = (int32_t)D.n[3];
    S->n3p = (int32_t)D.pitch;
    if (D.pitch % vw) return fail(CYTVDN_E_INVALID, "internal: row pitch %lld not a multiple of the vector width %d", (long long)D.pitch, vw);
    S->st1 = D.n[2] * D.pitch;
    S->st0 = D.n[1] * S->st1;
    if (S->st1 / vw > 0x3fffffff) return fail(CYTVDN_E_INVALID, "inner plane too large");
    S->n3v = (int32_t)(D.pitch / vw);
    S->mv = (int32_t)(S->st1 / vw);
    int64_t lo[2] = {0, 0}, hi[2] = {D.n[0], D.n[1]}, olo[2] = {0, 0}, ohi[2] = {D.n[0], D.n[1]};
    if (o) {
        for (int k = 0; k < 2; ++k) {
            lo[k] = o->box_lo[k];
            if (o->box_hi[k] > 0) hi[k] = o->box_hi[k];
            olo[k] = o->own_lo[k];
            if (o->own_hi[k] > 0) ohi[k] = o->own_hi[k];
            if (lo[k] < 0 || hi[k] > D.n[k] || lo[k] > hi[k])
                return fail(CYTVDN_E_INVALID, "opts box on axis %d is [%lld,%lld) for extent %lld", k,
                            (long long)lo[k], (long long)hi[k], (long long)D.n[k]);
        }
    }
    S->dynamic = (o && (o->flags & 1)) ? 1 : 0;
    S->own_store = (o && (o->flags & 2)) ? 1 : 0;
    S->i0 = (int32_t)lo[0]; S->ni = (int32_t)(hi[0] - lo[0]);
    S->j0 = (int32_t)lo[1];
    const int64_t nj = hi[1] - lo[1];
    S->oi0 = (int32_t)olo[0]; S->oi1 = (int32_t)ohi[0]; S->oj0 = (int32_t)olo[1]; S->oj1 = (int32_t)ohi[1];

    // strip width: keep TJ inner planes of every array of the sweep inside the L2 budget
    const int64_t plane_bytes = S->st1 * (int64_t)elem * footprint_arrays;
    int64_t tjmax = l2_budget(o) / (plane_bytes > 0 ? plane_bytes : 1);
    const int64_t cap = (1ll << 30) / (S->mv > 0 ? S->mv : 1);       // slab vector count must fit int32
    if (tjmax > cap) tjmax = cap;
    if (tjmax < 1) tjmax = 1;
    int64_t tj = nj;
    if (nj > tjmax) {
        const int64_t nstrips = (nj + tjmax - 1) / tjmax;
        tj = (nj + nstrips - 1) / nstrips;
    }
    if (tj < 1) tj = 1;
    S->tj = (int32_t)tj;
    S->nfull = (int32_t)(nj / tj);
    S->tl = (int32_t)(nj - (int64_t)S->nfull * tj);
    S->cps_full = (int32_t)((tj * S->mv + kBlock - 1) / kBlock);
    S->cps_last = (int32_t)(((int64_t)S->tl * S->mv + kBlock - 1) / kBlock);
    const int64_t tiles_full = (int64_t)S->ni * S->cps_full;
    const int64_t ntiles = tiles_full * S->nfull + (int64_t)S->ni * S->cps_last;
    if (tiles_full > 0x7fffffff || ntiles > 0x7fffffff) return fail(CYTVDN_E_INVALID, "array too large for one sweep");
    S->tiles_full = (int32_t)tiles_full;
    S->ntiles = (int32_t)ntiles;
    S->d_tiles_full = make_fastdiv((uint32_t)(S->tiles_full > 0 ? S->tiles_full : 1));
    S->d_cps_full = make_fastdiv((uint32_t)(S->cps_full > 0 ? S->cps_full : 1));
    S->d_cps_last = make_fastdiv((uint32_t)(S->cps_last > 0 ? S->cps_last : 1));
    S->d_mv = make_fastdiv((uint32_t)S->mv);
    S->d_n3v = make_fastdiv((uint32_t)S->n3v);
    return CYTVDN_OK;
}

template <typename T> constexpr int vec_width() { return 16 / (int)sizeof(T); }

// The reference's comparison-based clip leaves a value unchanged when the THRESHOLD is NaN (both comparisons are
// false); +inf does the same and lets the float kernels clip with max.NaN / min.NaN.
template <typename T> T clip_for_kernel(double c) { return std::isnan(c) ? (T)INFINITY : (T)c; }

// Elements per thread: 16-byte vectors when every row start is 16-byte aligned (extent of the fast axis a
// multiple of the vector width and all base pointers aligned); 8-byte vectors for float rows of even length;
// else scalar.  `ptr_bits` = OR of all base pointers of the launch.
template <typename T>
int pick_vw(int64_t n3, uintptr_t ptr_bits)
{
    constexpr int full = vec_width<T>();
    if (n3 % full == 0 && (ptr_bits & 15u) == 0) return full;
    if (sizeof(T) == 4 && n3 % 2 == 0 && (ptr_bits & 7u) == 0) return 2;
    return 1;
}
inline uintptr_t bits(const void *p) { return reinterpret_cast<uintptr_t>(p); }

// call f(std::integral_constant<int, VW>) for the run-time vector width
template <typename T, typename F>
int dispatch_vw(int vw, F &&f)
{
    if (vw == vec_width<T>()) return f(std::integral_constant<int, vec_width<T>()>{});
    if constexpr (sizeof(T) == 4) {
        if (vw == 2) return f(std::integral_constant<int, 2>{});
    }
    return f(std::integral_constant<int, 1>{});
}

// ---- half-step A -------------------------------------------------------------------------------
template <typename T, int VW, bool FISTA>
int launch_acc_mode(int mode, const AccParams<T> &P, cudaStream_t st)
{
    int grid = 1;
    if (P.S.ntiles <= 0) return CYTVDN_OK;
    switch (mode) {
    case ACC_ALL4: {
        auto k = tv_accumulator_kernel<T, VW, FISTA, ACC_ALL4>;
        if (int rc = grid_for(k, P.S.ntiles, &grid)) return rc;
        k<<<grid, kBlock, 0, st>>>(P);
        break;
    }
    case ACC_ALL3: {
        auto k = tv_accumulator_kernel<T, VW, FISTA, ACC_ALL3>;
        if (int rc = grid_for(k, P.S.ntiles, &grid)) return rc;
        k<<<grid, kBlock, 0, st>>>(P);
        break;
    }
    default: {
        auto k = tv_accumulator_kernel<T, VW, FISTA, ACC_GEN>;
        if (int rc = grid_for(k, P.S.ntiles, &grid)) return rc;
        k<<<grid, kBlock, 0, st>>>(P);
        break;
    }
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CUDA_TRY(cudaGetLastError());
    return CYTVDN_OK;
}

struct AccCall {
    Dims D;
    const void *a;
    void *b[4];          // embedded axis order
    void *d[4];
    double clip[4];
    int bc[4];
    bool fista;
    double tk;
    int mode;
    int axmask;
    int iso_p, iso_q, iso_mask;
    double *norm_dev;
    const cytvdn_step_opts *opts;
    cudaStream_t st;
};

template <typename T>
int run_acc(const AccCall &c)
{
    AccParams<T> P;
    memset(&P, 0, sizeof P);
    uintptr_t pb = bits(c.a);
    int nax = 0;
    for (int k = 0; k < 4; ++k) {
        const bool on = c.mode == ACC_ALL4 ? true : c.mode == ACC_ALL3 ? (k != 2) : ((c.axmask >> k) & 1);
        P.clip[k] = clip_for_kernel<T>(c.clip[k]);
        P.bc[k] = c.bc[k];
        if (!on) continue;
        ++nax;
        if (!c.b[k]) return fail(CYTVDN_E_INVALID, "accumulator for axis slot %d is NULL", k);
        if (c.fista && !c.d[k]) return fail(CYTVDN_E_INVALID, "FISTA auxiliary for axis slot %d is NULL", k);
        pb |= bits(c.b[k]) | (c.fista ? bits(c.d[k]) : 0);
        P.b[k] = (T *)c.b[k];
        P.d[k] = (T *)c.d[k];
    }
    const int vw = pick_vw<T>(c.D.pitch, pb);
    if (int rc = make_sweep(c.D, vw, sizeof(T), c.opts, 1 + nax * (c.fista ? 2 : 1), &P.S)) return rc;
    P.u = (const T *)c.a;
    P.tk = (T)c.tk;
    P.axmask = c.axmask;
    P.iso_p = c.iso_p; P.iso_q = c.iso_q; P.iso_mask = c.iso_mask;
    Workspace w;
    if (int rc = get_workspace(c.st, &w)) return rc;
    P.W.partials = w.partials; P.W.ticket = w.ticket; P.W.out = c.norm_dev;
    if (P.S.ntiles <= 0) {       // empty box: the sum is 0
        CUDA_TRY(cudaMemsetAsync(c.norm_dev, 0, sizeof(double), c.st));
        return CYTVDN_OK;
    }
    return dispatch_vw<T>(vw, [&](auto VWc) {
        constexpr int VW = decltype(VWc)::value;
        return c.fista ? launch_acc_mode<T, VW, true>(c.mode, P, c.st) : launch_acc_mode<T, VW, false>(c.mode, P, c.st);
    });
}

int check_common(int dtype, const void *a, const double *out_dev)
{
    if (dtype != CYTVDN_F32 && dtype != CYTVDN_F64) return fail(CYTVDN_E_INVALID, "dtype must be CYTVDN_F32 or CYTVDN_F64");
    if (!a) return fail(CYTVDN_E_INVALID, "input array pointer is NULL");
    if (!out_dev) return fail(CYTVDN_E_INVALID, "reduction output pointer is NULL");
    return CYTVDN_OK;
}

// ---- half-step B -------------------------------------------------------------------------------
template <typename T>
int run_dcu(const Dims &D, const void *orig, const void *uin, void *uout, const void *const *b,
            const double *w, int zero_wrap, bool self_wrap, double *sums_dev, const cytvdn_step_opts *opts,
            cudaStream_t st)
{
    DcuParams<T> P;
    memset(&P, 0, sizeof P);
    uintptr_t pb = bits(orig) | bits(uin) | bits(uout);
    for (int k = 0; k < D.ndim; ++k) {
        const int s = D.axmap[k];
        if (!b[k]) return fail(CYTVDN_E_INVALID, "b[%d] is NULL", k);
        pb |= bits(b[k]);
        P.b[s] = (const T *)b[k];
        P.w[s] = (T)w[k];
        if ((zero_wrap >> k) & 1) P.zero_wrap |= 1 << s;
        if (self_wrap) P.self_wrap |= 1 << s;
    }
    const int vw = pick_vw<T>(D.pitch, pb);
    if (int rc = make_sweep(D, vw, sizeof(T), opts, 2 + D.ndim, &P.S)) return rc;
    P.f = (const T *)orig; P.uin = (const T *)uin; P.uout = (T *)uout;
    P.ref = opts ? (const T *)opts->sse_reference : nullptr;
    Workspace ws;
    if (int rc = get_workspace(st, &ws)) return rc;
    P.W.partials = ws.partials; P.W.ticket = ws.ticket; P.W.out = sums_dev;
    if (P.S.ntiles <= 0) {
        CUDA_TRY(cudaMemsetAsync(sums_dev, 0, (P.ref ? 3 : 2) * sizeof(double), st));
        return CYTVDN_OK;
    }
    if (P.ref && vw != vec_width<T>())
        return fail(CYTVDN_E_UNSUPPORTED, "sse_reference needs 16-byte aligned rows (use cytvdn_sum_square_error otherwise)");
    if (int rc = dispatch_vw<T>(vw, [&](auto VWc) -> int {
            constexpr int VW = decltype(VWc)::value;
            int grid = 1;
            auto go = [&](auto k) -> int {
                if (int rc = grid_for(k, P.S.ntiles, &grid)) return rc;
                k<<<grid, kBlock, 0, st>>>(P);
                return CYTVDN_OK;
            };
            if constexpr (VW == vec_width<T>()) {
                if (P.ref) return D.ndim == 4 ? go(tv_datacube_kernel<T, VW, true, true>) : go(tv_datacube_kernel<T, VW, false, true>);
            }
            return D.ndim == 4 ? go(tv_datacube_kernel<T, VW, true>) : go(tv_datacube_kernel<T, VW, false>);
        }))
        return rc;
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CUDA_TRY(cudaGetLastError());
    return CYTVDN_OK;
}

template <typename T>
int run_sse(int64_t n, const void *a, const void *b, double *out, cudaStream_t st, int64_t n3 = 1, int64_t n3p = 1)
{
    Workspace ws;
    if (int rc = get_workspace(st, &ws)) return rc;
    RedWork W{ws.partials, ws.ticket, out};
    if (n <= 0) { CUDA_TRY(cudaMemsetAsync(out, 0, sizeof(double), st)); return CYTVDN_OK; }
    int grid = 1;
    auto k = tv_sse_kernel<T>;
    const int64_t blocks = (n + kBlock - 1) / kBlock;
    if (int rc = grid_for(k, (int)(blocks > 0x7fffffff ? 0x7fffffff : blocks), &grid)) return rc;
    k<<<grid, kBlock, 0, st>>>((const T *)a, (const T *)b, n, (int32_t)n3, (int32_t)n3p, W);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CUDA_TRY(cudaGetLastError());
    return CYTVDN_OK;
}

// ---- fused iteration ------------------------------------------------------------------------------
struct FusedCall {
    const void *lo_u, *hi_u, *hi_b0, *hi_d0;     // peer pointers (axis-0 halo on neighbouring GPUs) or NULL
    Dims D;
    const void *orig, *uin;
    void *uout;
    const void *bin[4], *din[4];     // embedded axis order
    void *bout[4], *dout[4];
    double clip[4], w[4];
    int bc[4];
    bool fista;
    int iso_mask;                                // bit 0: pair (0,1), bit 1: pair (2,3) half-isotropic
    bool mirror;                                 // BC_mode 3 (all axes)
    double tk;
    int zero_wrap;
    double *sums_dev;
    const cytvdn_step_opts *opts;
    cudaStream_t st;
};

template <typename T>
int run_fused(const FusedCall &c)
{
    FusedParams<T> P;
    memset(&P, 0, sizeof P);
    uintptr_t pb = bits(c.orig) | bits(c.uin) | bits(c.uout) | bits(c.lo_u) | bits(c.hi_u) | bits(c.hi_b0) | bits(c.hi_d0);
    int nax = 0;
    for (int k = 0; k < 4; ++k) {
        if (c.D.ndim == 3 && k == 2) continue;
        ++nax;
        if (!c.bin[k] || !c.bout[k]) return fail(CYTVDN_E_INVALID, "accumulator in/out for axis slot %d is NULL", k);
        if (c.fista && (!c.din[k] || !c.dout[k])) return fail(CYTVDN_E_INVALID, "FISTA auxiliary in/out for axis slot %d is NULL", k);
        if (c.bin[k] == c.bout[k] || (c.fista && c.din[k] == c.dout[k]))
            return fail(CYTVDN_E_INVALID, "the fused iteration is out of place: in and out arrays must differ");
        pb |= bits(c.bin[k]) | bits(c.bout[k]) | (c.fista ? (bits(c.din[k]) | bits(c.dout[k])) : 0);
        P.bin[k] = (const T *)c.bin[k]; P.bout[k] = (T *)c.bout[k];
        P.din[k] = (const T *)c.din[k]; P.dout[k] = (T *)c.dout[k];
        P.clip[k] = clip_for_kernel<T>(c.clip[k]); P.w[k] = (T)c.w[k]; P.bc[k] = c.bc[k];
    }
    if (c.uin == c.uout) return fail(CYTVDN_E_INVALID, "the fused iteration is out of place: recon_in == recon_out");
    const int vw = pick_vw<T>(c.D.pitch, pb);
    const bool vec = vw == vec_width<T>();
    // every array of the sweep transits L2 and recon must survive 2*TJ planes (read as x+e0, x, x-e0)
    const int arrays = 3 + nax * (c.fista ? 4 : 2);
    if (int rc = make_sweep(c.D, vw, sizeof(T), c.opts, 2 * arrays, &P.S)) return rc;
    P.f = (const T *)c.orig; P.uin = (const T *)c.uin; P.uout = (T *)c.uout;
    P.tk = (T)c.tk; P.zero_wrap = c.zero_wrap;
    P.lo_u = (const T *)c.lo_u; P.hi_u = (const T *)c.hi_u; P.hi_b0 = (const T *)c.hi_b0; P.hi_d0 = (const T *)c.hi_d0;
    P.ref = c.opts ? (const T *)c.opts->sse_reference : nullptr;
    if (c.mirror) {
        if (!vec || c.iso_mask || c.lo_u || c.hi_u || P.ref)
            return fail(CYTVDN_E_UNSUPPORTED, "BC_mode=3 in the fused iteration: anisotropic, 16-byte aligned rows, no peer "
                                              "pointers, no sse_reference (the two-pass kernels cover the rest)");
        for (int k = 0; k < 4; ++k)
            if ((c.D.ndim == 4 || k != 2) && c.D.n[k] < 2) return fail(CYTVDN_E_INVALID, "mirror boundary needs extent >= 2 on every axis");
    }
    if (P.ref && (c.iso_mask || !vec || c.lo_u || c.hi_u))
        return fail(CYTVDN_E_UNSUPPORTED, "sse_reference in the fused iteration: anisotropic, 16-byte aligned rows, no peer "
                                          "pointers (use cytvdn_sum_square_error otherwise)");
    Workspace ws;
    if (int rc = get_workspace(c.st, &ws)) return rc;
    P.W.partials = ws.partials; P.W.ticket = ws.ticket; P.W.out = c.sums_dev;
    if (P.S.ntiles <= 0) {
        CUDA_TRY(cudaMemsetAsync(c.sums_dev, 0, (P.ref ? 4 : 3) * sizeof(double), c.st));
        return CYTVDN_OK;
    }
    int grid = 1;
    const bool ax2 = c.D.ndim == 4;
    // experimental TMA-staged variant (measured slower, see fused_tma.cuh): opt-in, vector path, static tile order
    bool use_tma = false;
    { const char *env = getenv("CYTVDN_FUSED_TMA"); if (env && !strcmp(env, "1")) use_tma = vec && !P.S.dynamic && P.S.n3p == P.S.n3; }
    if (use_tma) {
        const int nself = 2 + ((ax2 ? 3 : 2) + 1) * (c.fista ? 2 : 1);
        const size_t smem = (size_t)2 * nself * kBlock * 16;
#define LAUNCH_TMA(FV, AX2V)                                                           \
    do {                                                                               \
        auto k = tv_fused_tma_kernel<T, vec_width<T>(), FV, AX2V>;                     \
        if (int rc = grid_for(k, P.S.ntiles, &grid, smem)) return rc;                  \
        k<<<grid, kBlock, smem, c.st>>>(P);                                            \
    } while (0)
        if (c.fista) { if (ax2) LAUNCH_TMA(true, true); else LAUNCH_TMA(true, false); }
        else         { if (ax2) LAUNCH_TMA(false, true); else LAUNCH_TMA(false, false); }
#undef LAUNCH_TMA
        g_launches.fetch_add(1, std::memory_order_relaxed);
        CUDA_TRY(cudaGetLastError());
        return CYTVDN_OK;
    }
    if (c.iso_mask) {
        // half-isotropic pairs: 4-D, full vector width (cytvdn_denoise and the sharded drivers pad rows to it)
        if (!ax2 || !vec)
            return fail(CYTVDN_E_UNSUPPORTED, "the fused half-isotropic iteration needs a 4-D array whose rows are 16-byte "
                                              "aligned (row length or cytvdn_step_opts.row_pitch a multiple of %d elements)",
                        vec_width<T>());
        if (c.lo_u || c.hi_u) return fail(CYTVDN_E_UNSUPPORTED, "peer pointers are not supported with half-isotropic pairs");
        auto go = [&](auto k) -> int {
            if (int rc = grid_for(k, P.S.ntiles, &grid)) return rc;
            k<<<grid, kBlock, 0, c.st>>>(P);
            return CYTVDN_OK;
        };
        int rc = CYTVDN_OK;
        const bool R = c.iso_mask & 1, Q = c.iso_mask & 2;
        if (c.fista) rc = R ? (Q ? go(tv_fused_iso_kernel<T, true, true, true>) : go(tv_fused_iso_kernel<T, true, true, false>))
                            : go(tv_fused_iso_kernel<T, true, false, true>);
        else rc = R ? (Q ? go(tv_fused_iso_kernel<T, false, true, true>) : go(tv_fused_iso_kernel<T, false, true, false>))
                    : go(tv_fused_iso_kernel<T, false, false, true>);
        if (rc) return rc;
        g_launches.fetch_add(1, std::memory_order_relaxed);
        CUDA_TRY(cudaGetLastError());
        return CYTVDN_OK;
    }
    if (int rc = dispatch_vw<T>(vw, [&](auto VWc) -> int {
            constexpr int VW = decltype(VWc)::value;
            auto go = [&](auto k) -> int {
                if (int rc = grid_for(k, P.S.ntiles, &grid)) return rc;
                k<<<grid, kBlock, 0, c.st>>>(P);
                return CYTVDN_OK;
            };
            if constexpr (VW == vec_width<T>()) {
                if (c.mirror) {                                 // BC_mode 3
                    if (c.fista) return ax2 ? go(tv_fused_kernel<T, VW, true, true, false, false, true>) : go(tv_fused_kernel<T, VW, true, false, false, false, true>);
                    return ax2 ? go(tv_fused_kernel<T, VW, false, true, false, false, true>) : go(tv_fused_kernel<T, VW, false, false, false, false, true>);
                }
                if (P.ref) {                                    // sum (ref - recon')^2 in the same pass
                    if (c.fista) return ax2 ? go(tv_fused_kernel<T, VW, true, true, false, true>) : go(tv_fused_kernel<T, VW, true, false, false, true>);
                    return ax2 ? go(tv_fused_kernel<T, VW, false, true, false, true>) : go(tv_fused_kernel<T, VW, false, false, false, true>);
                }
            }
            auto pick = [&](auto PEERc) -> int {
                constexpr bool PEER = decltype(PEERc)::value;
                if (c.fista) return ax2 ? go(tv_fused_kernel<T, VW, true, true, PEER>) : go(tv_fused_kernel<T, VW, true, false, PEER>);
                return ax2 ? go(tv_fused_kernel<T, VW, false, true, PEER>) : go(tv_fused_kernel<T, VW, false, false, PEER>);
            };
            const bool peer = c.lo_u || c.hi_u;
            return peer ? pick(std::true_type{}) : pick(std::false_type{});
        }))
        return rc;
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CUDA_TRY(cudaGetLastError());
    return CYTVDN_OK;
}

bool is_device_ptr(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

}  // namespace

// =================================================================================================
extern "C" {

int cytvdn_version(void) { return CYTVDN_VERSION; }
const char *cytvdn_last_error(void) { return g_err.c_str(); }
int64_t cytvdn_launch_count(void) { return g_launches.load(); }

int cytvdn_device_count(int *count)
{
    if (!count) return fail(CYTVDN_E_INVALID, "count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
    *count = n;
    return CYTVDN_OK;
}

int cytvdn_accumulator_update(int ndim, const int64_t *shape, int dtype, const void *a, void *b, void *d,
                              double tk, int ax, double clip, int bc_mode, double *norm_dev,
                              const cytvdn_step_opts *opts, void *stream)
{
    AccCall c;
    memset(&c, 0, sizeof c);
    if (int rc = make_dims(ndim, shape, &c.D, opts)) return rc;
    if (int rc = check_common(dtype, a, norm_dev)) return rc;
    if (ax < 0 || ax >= ndim) return fail(CYTVDN_E_INVALID, "ax=%d out of range for ndim=%d", ax, ndim);
    if (bc_mode < 0 || bc_mode > 3) return fail(CYTVDN_E_INVALID, "BC_mode must be 0, 1, 2 or 3");
    if (bc_mode == 3) bc_mode = 1;                    // half-step A of the clamped mirror is the reference's mirror
    if (bc_mode == 1 && shape[ax] < 2) return fail(CYTVDN_E_INVALID, "mirror boundary needs extent >= 2 on axis %d", ax);
    if (!b) return fail(CYTVDN_E_INVALID, "b is NULL");
    const int s = c.D.axmap[ax];
    c.a = a; c.b[s] = b; c.d[s] = d; c.clip[s] = clip; c.bc[s] = bc_mode;
    c.fista = d != nullptr; c.tk = tk; c.mode = ACC_GEN; c.axmask = 1 << s;
    c.iso_p = c.iso_q = -1; c.norm_dev = norm_dev; c.opts = opts; c.st = (cudaStream_t)stream;
    return dtype == CYTVDN_F32 ? run_acc<float>(c) : run_acc<double>(c);
}

int cytvdn_iso_accumulator_update(const int64_t *shape, int dtype, const void *a, void *b1, void *b2, void *d1,
                                  void *d2, double tk, int ax1, int ax2, double clip, double *norm_dev,
                                  const cytvdn_step_opts *opts, void *stream)
{
    AccCall c;
    memset(&c, 0, sizeof c);
    if (int rc = make_dims(4, shape, &c.D, opts)) return rc;
    if (int rc = check_common(dtype, a, norm_dev)) return rc;
    if (ax1 < 0 || ax1 > 3 || ax2 < 0 || ax2 > 3 || ax1 == ax2)
        return fail(CYTVDN_E_INVALID, "ax1/ax2 must be two different axes in 0..3 (got %d, %d)", ax1, ax2);
    if (!b1 || !b2) return fail(CYTVDN_E_INVALID, "b1/b2 is NULL");
    if ((d1 == nullptr) != (d2 == nullptr)) return fail(CYTVDN_E_INVALID, "d1 and d2 must both be given or both be NULL");
    c.a = a; c.b[ax1] = b1; c.b[ax2] = b2; c.d[ax1] = d1; c.d[ax2] = d2;
    for (int k = 0; k < 4; ++k) c.clip[k] = clip;     // the pair shares one threshold (clip[0] in the kernel)
    c.bc[ax1] = c.bc[ax2] = 2;
    c.fista = d1 != nullptr; c.tk = tk; c.mode = ACC_GEN; c.axmask = (1 << ax1) | (1 << ax2);
    c.iso_p = ax1; c.iso_q = ax2; c.norm_dev = norm_dev; c.opts = opts; c.st = (cudaStream_t)stream;
    return dtype == CYTVDN_F32 ? run_acc<float>(c) : run_acc<double>(c);
}

int cytvdn_accumulator_update_all(int ndim, const int64_t *shape, int dtype, const void *a, void *const *b,
                                  void *const *d, double tk, const double *clip, int iso_R, int iso_Q,
                                  int bc_mode, double *norm_dev, const cytvdn_step_opts *opts, void *stream)
{
    AccCall c;
    memset(&c, 0, sizeof c);
    if (int rc = make_dims(ndim, shape, &c.D, opts)) return rc;
    if (int rc = check_common(dtype, a, norm_dev)) return rc;
    if (!b || !clip) return fail(CYTVDN_E_INVALID, "b / clip is NULL");
    if (bc_mode < 0 || bc_mode > 3) return fail(CYTVDN_E_INVALID, "BC_mode must be 0, 1, 2 or 3");
    if (bc_mode == 3) bc_mode = 1;
    if (ndim == 3 && (iso_R || iso_Q)) return fail(CYTVDN_E_INVALID, "half-isotropic update exists for 4-D only");
    for (int k = 0; k < ndim; ++k) {
        const int s = c.D.axmap[k];
        c.b[s] = b[k]; c.d[s] = d ? d[k] : nullptr; c.clip[s] = clip[k];
        const bool iso = (k < 2 && iso_R) || (k >= 2 && iso_Q);
        const bool force_jz = opts && ((opts->flags >> (8 + k)) & 1);   // split axis of a periodic sharded run
        c.bc[s] = (iso || force_jz) ? 2 : bc_mode;         // iso_* kernels know Jia-Zhao only
        if (c.bc[s] == 1 && shape[k] < 2) return fail(CYTVDN_E_INVALID, "mirror boundary needs extent >= 2 on axis %d", k);
    }
    c.a = a; c.fista = d != nullptr; c.tk = tk;
    c.mode = ndim == 4 ? ACC_ALL4 : ACC_ALL3;
    c.axmask = ndim == 4 ? 15 : 11;
    c.iso_p = c.iso_q = -1; c.iso_mask = (iso_R ? 1 : 0) | (iso_Q ? 2 : 0);
    c.norm_dev = norm_dev; c.opts = opts; c.st = (cudaStream_t)stream;
    return dtype == CYTVDN_F32 ? run_acc<float>(c) : run_acc<double>(c);
}

int cytvdn_datacube_update(int ndim, const int64_t *shape, int dtype, const void *orig, const void *recon_in,
                           void *recon_out, const void *const *b, const double *lambda_mu, int bc_mode,
                           double *sums_dev, const cytvdn_step_opts *opts, void *stream)
{
    Dims D;
    if (int rc = make_dims(ndim, shape, &D, opts)) return rc;
    if (int rc = check_common(dtype, orig, sums_dev)) return rc;
    if (!recon_in || !recon_out || !b || !lambda_mu) return fail(CYTVDN_E_INVALID, "recon / b / lambda_mu is NULL");
    if (bc_mode == 1)
        return fail(CYTVDN_E_UNSUPPORTED, "BC_mode=1 (mirror) is undefined behaviour in the reference's "
                                          "datacube_update (utils.pyx:117-120) and is not implemented; "
                                          "BC_mode=3 is the well-defined (clamped-index) mirror");
    if (bc_mode != 0 && bc_mode != 2 && bc_mode != 3) return fail(CYTVDN_E_INVALID, "BC_mode must be 0, 2 or 3");
    const int zw = opts ? opts->zero_wrap_mask : 0;
    const bool sw = bc_mode == 3;
    return dtype == CYTVDN_F32
               ? run_dcu<float>(D, orig, recon_in, recon_out, b, lambda_mu, zw, sw, sums_dev, opts, (cudaStream_t)stream)
               : run_dcu<double>(D, orig, recon_in, recon_out, b, lambda_mu, zw, sw, sums_dev, opts, (cudaStream_t)stream);
}

int cytvdn_fused_iteration(int ndim, const int64_t *shape, int dtype, const void *orig, const void *recon_in,
                           void *recon_out, const void *const *b_in, void *const *b_out, const void *const *d_in,
                           void *const *d_out, double tk, const double *clip, const double *lambda_mu, int bc_mode,
                           double *sums_dev, const cytvdn_step_opts *opts, void *stream)
{
    FusedCall c;
    memset(&c, 0, sizeof c);
    if (int rc = make_dims(ndim, shape, &c.D, opts)) return rc;
    if (int rc = check_common(dtype, orig, sums_dev)) return rc;
    if (!recon_in || !recon_out || !b_in || !b_out || !clip || !lambda_mu)
        return fail(CYTVDN_E_INVALID, "recon / b / clip / lambda_mu is NULL");
    if ((d_in == nullptr) != (d_out == nullptr)) return fail(CYTVDN_E_INVALID, "d_in and d_out must both be given or both be NULL");
    if (bc_mode == 1)
        return fail(CYTVDN_E_UNSUPPORTED, "BC_mode=1 (mirror) is undefined behaviour in the reference's "
                                          "datacube_update (utils.pyx:117-120) and is not implemented");
    if (bc_mode != 0 && bc_mode != 2 && bc_mode != 3) return fail(CYTVDN_E_INVALID, "BC_mode must be 0, 2 or 3");
    c.mirror = bc_mode == 3;
    if (c.mirror && opts && ((opts->flags >> 8) & 15))
        return fail(CYTVDN_E_UNSUPPORTED, "per-axis Jia-Zhao overrides (flags bits 8..11) do not combine with BC_mode=3");
    c.iso_mask = opts ? ((opts->flags >> 4) & 3) : 0;
    if (c.iso_mask && ndim != 4) return fail(CYTVDN_E_INVALID, "half-isotropic update exists for 4-D only");
    for (int k = 0; k < ndim; ++k) {
        const int s = c.D.axmap[k];
        c.bin[s] = b_in[k]; c.bout[s] = b_out[k];
        c.din[s] = d_in ? d_in[k] : nullptr; c.dout[s] = d_out ? d_out[k] : nullptr;
        c.clip[s] = clip[k]; c.w[s] = lambda_mu[k];
        const bool iso = (k < 2 && (c.iso_mask & 1)) || (k >= 2 && (c.iso_mask & 2));   // iso_* kernels know Jia-Zhao only
        c.bc[s] = (iso || (opts && ((opts->flags >> (8 + k)) & 1))) ? 2 : bc_mode;
        if (opts && ((opts->zero_wrap_mask >> k) & 1)) c.zero_wrap |= 1 << s;
    }
    c.orig = orig; c.uin = recon_in; c.uout = recon_out;
    if (opts) {
        c.lo_u = opts->peer_lo_recon; c.hi_u = opts->peer_hi_recon; c.hi_b0 = opts->peer_hi_b0; c.hi_d0 = opts->peer_hi_d0;
        if (c.hi_u && (!c.hi_b0 || (d_in && !c.hi_d0)))
            return fail(CYTVDN_E_INVALID, "peer_hi_recon needs peer_hi_b0 (and peer_hi_d0 with FISTA)");
    }
    c.fista = d_in != nullptr; c.tk = tk; c.sums_dev = sums_dev; c.opts = opts; c.st = (cudaStream_t)stream;
    return dtype == CYTVDN_F32 ? run_fused<float>(c) : run_fused<double>(c);
}

int cytvdn_sum_square_error(int64_t n, int dtype, const void *a, const void *b, double *sse_dev, void *stream)
{
    if (int rc = check_common(dtype, a, sse_dev)) return rc;
    if (!b) return fail(CYTVDN_E_INVALID, "b is NULL");
    if (n < 0) return fail(CYTVDN_E_INVALID, "n < 0");
    return dtype == CYTVDN_F32 ? run_sse<float>(n, a, b, sse_dev, (cudaStream_t)stream)
                               : run_sse<double>(n, a, b, sse_dev, (cudaStream_t)stream);
}

int cytvdn_synth_counts(const int64_t *gshape, int64_t offset0, int64_t lshape0, int dtype,
                        const float *scan_mod_dev, const float *templ_dev, double counts, uint64_t seed,
                        void *out_dev, void *stream)
{
    if (!gshape || !scan_mod_dev || !templ_dev || !out_dev) return fail(CYTVDN_E_INVALID, "NULL argument");
    if (dtype != CYTVDN_F32 && dtype != CYTVDN_F64) return fail(CYTVDN_E_INVALID, "bad dtype");
    if (offset0 < 0 || lshape0 < 0 || offset0 + lshape0 > gshape[0]) return fail(CYTVDN_E_INVALID, "bad block range");
    const int64_t m = gshape[2] * gshape[3];
    const int64_t st0 = gshape[1] * m;
    if (m > 0x7fffffff) return fail(CYTVDN_E_INVALID, "inner plane too large");
    const int64_t nloc = lshape0 * st0;
    if (nloc == 0) return CYTVDN_OK;
    DevInfo di;
    if (int rc = dev_info(&di)) return rc;
    int64_t blocks = (nloc + kBlock - 1) / kBlock;
    const int grid = (int)(blocks < (int64_t)di.sms * 8 ? blocks : (int64_t)di.sms * 8);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == CYTVDN_F32)
        tv_synth_kernel<float><<<grid, kBlock, 0, st>>>((float *)out_dev, nloc, offset0 * st0, m,
                                                       scan_mod_dev, templ_dev, (float)counts, seed);
    else
        tv_synth_kernel<double><<<grid, kBlock, 0, st>>>((double *)out_dev, nloc, offset0 * st0, m,
                                                        scan_mod_dev, templ_dev, (float)counts, seed);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CUDA_TRY(cudaGetLastError());
    return CYTVDN_OK;
}

// ---- the device-resident loop -------------------------------------------------------------------
static int validate_params(const cytvdn_denoise_params *p, Dims *D)
{
    if (!p) return fail(CYTVDN_E_INVALID, "params is NULL");
    if (int rc = make_dims(p->ndim, p->shape, D)) return rc;
    if (p->dtype != CYTVDN_F32 && p->dtype != CYTVDN_F64) return fail(CYTVDN_E_INVALID, "bad dtype");
    if (p->iters_fista < 0 || p->iters_plain < 0) return fail(CYTVDN_E_INVALID, "negative iteration count");
    if (p->bc_mode == 1)
        return fail(CYTVDN_E_UNSUPPORTED, "BC_mode=1 (mirror) is undefined behaviour in the reference's "
                                          "datacube_update (utils.pyx:117-120) and is not implemented; "
                                          "BC_mode=3 is the well-defined (clamped-index) mirror");
    if (p->bc_mode != 0 && p->bc_mode != 2 && p->bc_mode != 3) return fail(CYTVDN_E_INVALID, "BC_mode must be 0, 2 or 3");
    if (p->bc_mode == 3) {
        if (p->isotropic_R || p->isotropic_Q)
            return fail(CYTVDN_E_INVALID, "BC_mode=3 (mirror) exists for the anisotropic update only "
                                          "(the half-isotropic kernels know the Jia-Zhao boundary only, halfisotropic.pyx:63-95)");
        for (int k = 0; k < p->ndim; ++k)
            if (p->shape[k] < 2) return fail(CYTVDN_E_INVALID, "mirror boundary needs extent >= 2 on axis %d", k);
    }
    if (p->ndim == 3 && (p->isotropic_R || p->isotropic_Q))
        return fail(CYTVDN_E_INVALID, "half-isotropic update exists for 4-D only");
    return CYTVDN_OK;
}

namespace {
// Which schedule a run uses: 2 = fused single pass (76 B/voxel, needs a second set of b/d arrays),
// 1 = two passes (96 B/voxel, in place).  params->schedule: 0 auto, 1, 2; env CYTVDN_SCHEDULE overrides.
bool fused_possible(const cytvdn_denoise_params *p)
{
    return p->bc_mode == 0 || p->bc_mode == 2 || p->bc_mode == 3;    // (the mirror is anisotropic only: validate_params)
}
int requested_schedule(const cytvdn_denoise_params *p)
{
    int want = p->schedule;
    const char *env = getenv("CYTVDN_SCHEDULE");
    if (env && *env) {
        if (!strcmp(env, "fused") || !strcmp(env, "2")) want = 2;
        else if (!strcmp(env, "two_pass") || !strcmp(env, "1")) want = 1;
        else if (!strcmp(env, "streamed") || !strcmp(env, "3")) want = 3;
    }
    return want;
}
int64_t arrays_needed(const cytvdn_denoise_params *p, bool fused, bool data_dev, bool recon_dev, bool ref_host)
{
    const int nd = p->ndim;
    const bool any = p->iters_fista + p->iters_plain > 0;
    int64_t a = 0;
    if (any) a += (int64_t)nd * (p->iters_fista > 0 ? 2 : 1) * (fused ? 2 : 1);   // b (+d), ping-pong when fused
    if (!data_dev) a += 1;
    if (!recon_dev) a += 1;
    if (fused && any) a += 1;                                                       // second recon buffer
    if (ref_host) a += 1;
    return a;
}
}  // namespace

int cytvdn_denoise_workspace_bytes(const cytvdn_denoise_params *p, int data_on_device, int recon_on_device,
                                   int64_t *bytes)
{
    Dims D;
    if (int rc = validate_params(p, &D)) return rc;
    if (!bytes) return fail(CYTVDN_E_INVALID, "bytes is NULL");
    const int64_t elem = p->dtype == CYTVDN_F32 ? 4 : 8, vwf = 16 / elem;
    const bool padded = D.n[3] % vwf != 0;             // internal rows are padded: caller arrays are always copied
    const int64_t nb = D.n[0] * D.n[1] * D.n[2] * ((D.n[3] + vwf - 1) / vwf * vwf) * elem;
    const bool fused = fused_possible(p) && requested_schedule(p) != 1 && (requested_schedule(p) == 2 || !p->isotropic_R);
    const int64_t nbp = (int64_t)((nb + 255) & ~(int64_t)255);
    const int64_t scratch = (((int64_t)(p->iters_fista + p->iters_plain + 1) * 4 * 16 * 8 + 255) & ~(int64_t)255) + 8192;
    *bytes = arrays_needed(p, fused, data_on_device != 0 && !padded, recon_on_device != 0 && !padded, false) * nbp + scratch;
    return CYTVDN_OK;
}

namespace {
// All device state of one cytvdn_denoise call lives in ONE allocation that is carved up: on B200 twenty
// 4 GiB cudaMalloc/cudaFree pairs cost ~250 ms, one 80 GiB pair ~35 ms (tools/alloc_timing.py) -- and now and then
// several hundred ms (round 1: 0.66 s of a 2.1 s call).  A caller that denoises more than once can take even that
// out of its calls: cytvdn_workspace_reserve() makes the library hold one device block per device, and every Arena
// whose request fits carves from it instead of calling cudaMalloc (bump allocation, reset when the last borrower
// returns).  Nothing is ever cached behind the caller's back: without a reservation each call allocates and frees.
struct Reserved {
    char *base = nullptr;
    size_t size = 0, used = 0;
    int refs = 0;
};
std::mutex g_res_mutex;
std::map<int, Reserved> g_reserved;                       // device -> block held for the caller

struct Arena {
    char *base = nullptr;
    size_t size = 0, used = 0;
    int borrowed_dev = -1;                                // >= 0: carved from the reserved block of that device
    ~Arena() { release(); }
    void release()
    {
        if (base && borrowed_dev >= 0) {
            std::lock_guard<std::mutex> lk(g_res_mutex);
            auto it = g_reserved.find(borrowed_dev);
            if (it != g_reserved.end() && --it->second.refs == 0) it->second.used = 0;
        } else if (base) {
            cudaFree(base);
        }
        base = nullptr; size = used = 0; borrowed_dev = -1;
    }
    static size_t padded(size_t bytes) { return (bytes + 255) & ~(size_t)255; }
    int reserve(size_t bytes)
    {
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) {
            std::lock_guard<std::mutex> lk(g_res_mutex);
            auto it = g_reserved.find(dev);
            const size_t need = padded(bytes);
            if (it != g_reserved.end() && it->second.base && it->second.used + need <= it->second.size) {
                base = it->second.base + it->second.used;
                it->second.used += need;
                ++it->second.refs;
                size = bytes; borrowed_dev = dev;
                return CYTVDN_OK;
            }
        }
        cudaError_t e = cudaMalloc((void **)&base, bytes);
        if (e != cudaSuccess) {
            cudaGetLastError();
            base = nullptr;
            return fail(CYTVDN_E_NOMEM, "cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
        }
        size = bytes;
        return CYTVDN_OK;
    }
    int alloc(void **p, size_t bytes)
    {
        const size_t need = padded(bytes);
        if (used + need > size) return fail(CYTVDN_E_NOMEM, "internal: arena exhausted (%zu + %zu > %zu)", used, need, size);
        *p = base + used;
        used += need;
        return CYTVDN_OK;
    }
};

// host-clock milestones of the calling thread's last cytvdn_denoise (cytvdn_last_trace)
constexpr int kTraceMax = 12;
thread_local double g_trace_ms[kTraceMax];
thread_local const char *g_trace_what[kTraceMax];
thread_local int g_trace_n = 0;

// device memory a call can count on: what the driver reports free plus the unused part of the caller's reservation
int available_bytes(size_t *out)
{
    size_t free_b = 0, tot_b = 0;
    CUDA_TRY(cudaMemGetInfo(&free_b, &tot_b));
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_res_mutex);
    auto it = g_reserved.find(dev);
    if (it != g_reserved.end() && it->second.base) free_b += it->second.size - it->second.used;
    *out = free_b;
    return CYTVDN_OK;
}
}  // namespace

// ------------------------------------------------------------------------------------------------
// Host-side plans of the two PCIe schedules.  Pure functions (no CUDA): cytvdn_denoise runs exactly what they
// return, and the CPU tests check their invariants through cytvdn_pipeline_schedule / cytvdn_stream_plan.
// ------------------------------------------------------------------------------------------------
namespace {
// PCIe pipeline: the order of (box, iteration) launches for `n_iter` iterations over `nbox` boxes of axis-0 planes;
// box -1 = a sweep of the whole array.  The first and the last `nbox` iterations run as wavefronts: step t runs
// iteration m0 + (t - c) of box c, higher boxes first, so that box c+1 has finished iteration m-1 before box c
// starts iteration m (and box c-1, one iteration ahead, has not yet overwritten what box c still reads).
void pipeline_schedule(int nbox, int n_iter, std::vector<std::pair<int, int>> &out)
{
    out.clear();
    auto wave = [&](int m0, int m1) {
        const int depth = m1 - m0;
        for (int t = 0; t < nbox + depth - 1; ++t)
            for (int c = std::min(t, nbox - 1); c >= 0 && t - c < depth; --c) out.push_back({c, m0 + (t - c)});
    };
    const int depth = nbox;
    if (n_iter <= 2 * depth + 2) wave(0, n_iter);
    else {
        wave(0, depth);
        for (int m = depth; m < n_iter - depth; ++m) out.push_back({-1, m});
        wave(n_iter - depth, n_iter);
    }
}

// Out-of-core schedule: tile geometry for a device budget.
struct StreamPlan {
    int64_t planes_per_slot;   // P: axis-0 planes one tile slot holds
    int64_t iters_per_pass;    // K
    int64_t core_planes;       // planes a tile advances by K iterations (= P - 2K, or the whole axis)
    int64_t tiles, passes;
    int64_t arrays_per_slot;   // f, recon, b (+ d)
    int64_t plane_bytes;       // one axis-0 plane of an internal (row-padded) array
    int64_t host_state_bytes;  // page-locked host arrays the call allocates (b, d between passes)
};
int make_stream_plan(const cytvdn_denoise_params *p, const Dims &D, size_t budget, StreamPlan *sp, int ndev = 1)
{
    const int nd = p->ndim, nF = p->iters_fista, M = p->iters_fista + p->iters_plain;
    const int64_t elem = p->dtype == CYTVDN_F32 ? 4 : 8, full_vw = 16 / elem;
    const int64_t n0 = D.n[0], n3p = (D.n[3] + full_vw - 1) / full_vw * full_vw;
    sp->plane_bytes = D.n[1] * D.n[2] * n3p * elem;
    sp->arrays_per_slot = 2 + nd * (nF > 0 ? 2 : 1);
    if (M <= 0) return fail(CYTVDN_E_INVALID, "the out-of-core schedule needs at least one iteration");
    int64_t P = (int64_t)(budget / ((size_t)2 * sp->arrays_per_slot * sp->plane_bytes));
    if (ndev == 1 && P >= n0) { P = n0; sp->iters_per_pass = M; sp->core_planes = n0; }      // one tile: nothing is recomputed
    else {
        // two slots of P planes + the carry buffer (2K = P/2 planes): 2.5 P planes of every array; several devices
        // additionally snapshot the K planes above their last tile (edge buffer): 2.75 P
        P = ndev == 1 ? (int64_t)(budget / ((size_t)5 * sp->arrays_per_slot * sp->plane_bytes / 2))
                      : (int64_t)(budget / ((size_t)11 * sp->arrays_per_slot * sp->plane_bytes / 4));
        if (P < 4)
            return fail(CYTVDN_E_NOMEM, "out-of-core schedule: two tiles of 4 planes and their carry buffer (%lld bytes) do "
                                        "not fit in the device budget of %zu bytes",
                        (long long)(10 * sp->arrays_per_slot * sp->plane_bytes), budget);
        sp->iters_per_pass = std::min<int64_t>(M, std::max<int64_t>(1, P / 4));     // P/4 minimises the traffic per iteration
        sp->core_planes = P - 2 * sp->iters_per_pass;
        if (ndev > 1) {     // a multiple of ndev tiles of equal size, so that every device gets the same work
            const int64_t per_dev = (n0 + (int64_t)ndev * sp->core_planes - 1) / ((int64_t)ndev * sp->core_planes);
            sp->core_planes = std::max<int64_t>(1, (n0 + per_dev * ndev - 1) / (per_dev * ndev));
        }
    }
    sp->planes_per_slot = P;
    sp->tiles = (n0 + sp->core_planes - 1) / sp->core_planes;
    sp->passes = (M + sp->iters_per_pass - 1) / sp->iters_per_pass;
    sp->host_state_bytes = 0;
    if (sp->passes > 1)
        sp->host_state_bytes = (int64_t)nd * (1 + ((nF > 0 && nF > sp->iters_per_pass) ? 1 : 0)) * n0 * sp->plane_bytes;
    return CYTVDN_OK;
}
}  // namespace

// ------------------------------------------------------------------------------------------------
// Out-of-core schedule: the arrays do not fit in HBM, host arrays in and out.
//
// Temporal blocking with overlapped tiles of axis-0 planes.  One PASS advances every voxel by K iterations: a tile
// = `core` planes + K halo planes towards each neighbouring tile is copied in (state after m iterations), iterated K
// times on a box that shrinks by one plane per iteration and side (a plane can be advanced only while both
// neighbours hold the previous iterate), which leaves exactly the core at state m+K; the core is copied back.
// Between passes the state (recon in the caller's `recon` array; b and d in pinned host arrays owned by the call)
// lives on the host.  The first pass needs the input only (b = d = 0, recon = input), the last pass returns the
// reconstruction only.  Two tile slots on the device: tile t+1 is copied in while tile t iterates.  The 2K planes
// two neighbouring tiles share are not copied in twice: they are saved (device to device) into a carry buffer
// before tile t starts to iterate and handed to tile t+1, so every plane crosses the bus once per pass and the copy
// in of tile t+1 never reads host planes that the copy back of tile t overwrites.  The schedule is PCIe bound by a wide margin (per pass and voxel the bus carries the whole
// state in and out, the kernels need 1/7 of that time for K ~ 20 iterations), so what counts is K, i.e. planes per
// slot: the tiles are iterated with the IN-PLACE two-pass kernels (10 arrays per slot for 4-D FISTA) rather than
// the fused one (19) -- measured 1.9x faster end to end (tools/stream_bench.py).  Half-step A sweeps one plane more
// than half-step B at the upper end (B reads the forward neighbour of b).  The kernels are the in-core ones (boxes,
// owned-range sums, zero_wrap), so the reconstruction is bit-identical to the in-core run.
// Jia-Zhao boundary, fixed iteration counts, no reference_data; anisotropic and half-isotropic.
// ------------------------------------------------------------------------------------------------
namespace {
// Page-locked host memory.  cudaMallocHost pins 4 KB pages at ~2.4 GB/s (14 s for the 34 GB host state of a
// config-3 sized out-of-core run).  Instead: an anonymous mapping advised to use transparent huge pages, faulted in
// from all cores, then cudaHostRegister -- 1.9 s for the same 34 GB, same copy rates.  Falls back to cudaMallocHost
// when the mapping or the registration fails; CYTVDN_HOST_ALLOC=cuda selects cudaMallocHost outright.
std::mutex g_host_mu;
std::map<void *, size_t> g_host_mapped;                  // allocations made by mmap + cudaHostRegister

int pinned_alloc(void **out, size_t bytes)
{
    *out = nullptr;
    const char *env = getenv("CYTVDN_HOST_ALLOC");
    if (!(env && !strcmp(env, "cuda")) && bytes >= ((size_t)8 << 20)) {
        const size_t len = (bytes + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
        void *q = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (q != MAP_FAILED) {
            madvise(q, len, MADV_HUGEPAGE);
            {   // fault the pages in (the kernel zeroes them) from all cores; registering then only pins
                const unsigned hw = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
                std::vector<std::thread> th;
                const size_t chunk = ((len / hw) + 4095) & ~(size_t)4095;
                for (unsigned i = 0; i < hw; ++i)
                    th.emplace_back([=] {
                        const size_t lo = (size_t)i * chunk, hi = std::min(len, lo + chunk);
                        for (size_t x = lo; x < hi; x += 4096) ((volatile char *)q)[x] = 0;
                    });
                for (auto &t : th) t.join();
            }
            if (cudaHostRegister(q, len, cudaHostRegisterPortable) == cudaSuccess) {
                std::lock_guard<std::mutex> lk(g_host_mu);
                g_host_mapped[q] = len;
                *out = q;
                return CYTVDN_OK;
            }
            cudaGetLastError();
            munmap(q, len);
        }
    }
    cudaError_t e = cudaMallocHost(out, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *out = nullptr;
        return fail(CYTVDN_E_NOMEM, "page-locked host allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    }
    return CYTVDN_OK;
}

int pinned_free(void *q)
{
    if (!q) return CYTVDN_OK;
    size_t len = 0;
    {
        std::lock_guard<std::mutex> lk(g_host_mu);
        auto it = g_host_mapped.find(q);
        if (it != g_host_mapped.end()) { len = it->second; g_host_mapped.erase(it); }
    }
    if (len) {
        cudaError_t e = cudaHostUnregister(q);
        munmap(q, len);
        if (e != cudaSuccess) { cudaGetLastError(); return fail(CYTVDN_E_CUDA, "cudaHostUnregister: %s", cudaGetErrorString(e)); }
        return CYTVDN_OK;
    }
    CUDA_TRY(cudaFreeHost(q));
    return CYTVDN_OK;
}

struct HostPinned {
    std::vector<void *> p;
    ~HostPinned() { for (void *q : p) pinned_free(q); }
    int alloc(void **out, size_t bytes)
    {
        if (int rc = pinned_alloc(out, bytes)) return rc;
        p.push_back(*out);
        return CYTVDN_OK;
    }
};

// What the devices of one out-of-core run share: the tile geometry, the host state between passes, the per-tile sums
// and a barrier (one host thread per device; a single device never waits).
struct StreamShared {
    StreamPlan sp;
    int ndev = 1;
    void *hb[4] = {nullptr, nullptr, nullptr, nullptr}, *hd[4] = {nullptr, nullptr, nullptr, nullptr};
    std::vector<double> sums_h;                       // [M][tiles][4]
    // barrier; wait() returns false once any device has failed (everybody then unwinds)
    std::mutex mu;
    std::condition_variable cv;
    int arrived = 0;
    uint64_t phase = 0;
    bool failed = false;
    bool wait()
    {
        if (ndev == 1) return !failed;
        std::unique_lock<std::mutex> lk(mu);
        if (failed) return false;
        if (++arrived == ndev) { arrived = 0; ++phase; cv.notify_all(); return true; }
        const uint64_t ph = phase;
        cv.wait(lk, [&] { return phase != ph || failed; });
        return !failed;
    }
    void fail_all() { std::lock_guard<std::mutex> lk(mu); failed = true; cv.notify_all(); }
};

// One device's share of the out-of-core schedule: tiles [t_lo, t_hi) of every pass.
int stream_worker(const cytvdn_denoise_params *p, const Dims &D, const void *data, void *recon, StreamShared &sh, int t_lo, int t_hi,
                  cudaStream_t user_stream)
{
    const int nd = p->ndim, nF = p->iters_fista, nU = p->iters_plain, M = nF + nU;
    const size_t elem = p->dtype == CYTVDN_F32 ? 4 : 8;
    const int64_t full_vw = 16 / (int64_t)elem;
    const int64_t n0 = D.n[0], n3 = D.n[3], n3p = (n3 + full_vw - 1) / full_vw * full_vw;
    const bool padded = n3p != n3;
    const int64_t plane_rows = D.n[1] * D.n[2];
    const size_t plane_b = (size_t)plane_rows * n3p * elem;           // one plane of an internal array
    const bool fista = nF > 0;
    const StreamPlan &sp = sh.sp;
    const int64_t P = sp.planes_per_slot, K = sp.iters_per_pass, core = sp.core_planes;
    const int nt = (int)sp.tiles;
    const int ntl = std::max(0, t_hi - t_lo);                         // tiles of this device
    const bool multi = sh.ndev > 1;
    void *const *hb = sh.hb, *const *hd = sh.hd;

    Arena pool;
    const size_t slot_b = (size_t)sp.arrays_per_slot * Arena::padded((size_t)P * plane_b);
    const size_t nsums = (size_t)M * std::max(ntl, 1) * 4;
    // slot[2] is the CARRY buffer: the 2K planes two neighbouring tiles share are saved from tile t's slot before it
    // starts iterating and handed to tile t+1, so that every plane crosses the bus once per pass.
    // slot[3] is the EDGE buffer (several devices only): the K planes above this device's last tile belong to the next
    // device, which writes them back early in the pass -- they are snapshotted before any device writes anything.
    const int64_t carry_planes = ntl > 1 ? 2 * K : 0;
    const int64_t edge_planes = (multi && t_hi < nt && ntl > 0) ? K : 0;
    const size_t carry_b = (size_t)sp.arrays_per_slot * Arena::padded((size_t)carry_planes * plane_b);
    const size_t edge_b = (size_t)sp.arrays_per_slot * Arena::padded((size_t)edge_planes * plane_b);
    if (ntl > 0)
        if (int rc = pool.reserve(2 * slot_b + carry_b + edge_b + Arena::padded(nsums * sizeof(double)) + 4096)) return rc;
    struct Slot { void *f, *r, *b[4], *d[4]; } slot[4];
    memset(slot, 0, sizeof slot);
    for (int q = 0; q < 4 && ntl > 0; ++q) {
        Slot &sl = slot[q];
        const size_t planes = q < 2 ? (size_t)P : q == 2 ? (size_t)carry_planes : (size_t)edge_planes;
        if (planes == 0) continue;
        if (int rc = pool.alloc(&sl.f, planes * plane_b)) return rc;
        if (int rc = pool.alloc(&sl.r, planes * plane_b)) return rc;
        for (int k = 0; k < nd; ++k) {
            if (int rc = pool.alloc(&sl.b[k], planes * plane_b)) return rc;
            if (fista) if (int rc = pool.alloc(&sl.d[k], planes * plane_b)) return rc;
        }
    }
    Slot &carry = slot[2], &edge = slot[3];
    double *sums_d = nullptr;
    if (ntl > 0) if (int rc = pool.alloc((void **)&sums_d, nsums * sizeof(double))) return rc;

    struct Streams {
        cudaStream_t up = nullptr, comp = nullptr, down = nullptr;      // comp may be the caller's stream (not owned)
        bool own_comp = false;
        cudaEvent_t up_done[2] = {0, 0}, comp_done[2] = {0, 0}, down_done[2] = {0, 0};
        ~Streams()
        {
            // error paths leave copies in flight that reference the host state freed right after this object
            if (up) cudaStreamSynchronize(up);
            if (down) cudaStreamSynchronize(down);
            if (comp && own_comp) cudaStreamSynchronize(comp);
            for (int q = 0; q < 2; ++q) {
                if (up_done[q]) cudaEventDestroy(up_done[q]);
                if (comp_done[q]) cudaEventDestroy(comp_done[q]);
                if (down_done[q]) cudaEventDestroy(down_done[q]);
            }
            if (up) cudaStreamDestroy(up);
            if (down) cudaStreamDestroy(down);
            if (comp && own_comp) cudaStreamDestroy(comp);
        }
    } S;
    CUDA_TRY(cudaStreamCreateWithFlags(&S.up, cudaStreamNonBlocking));
    if (multi) { CUDA_TRY(cudaStreamCreateWithFlags(&S.comp, cudaStreamNonBlocking)); S.own_comp = true; }
    else S.comp = user_stream;                        // kernels run on the caller's stream (its reduction workspace is cached)
    CUDA_TRY(cudaStreamCreateWithFlags(&S.down, cudaStreamNonBlocking));
    for (int q = 0; q < 2; ++q) {
        CUDA_TRY(cudaEventCreateWithFlags(&S.up_done[q], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&S.comp_done[q], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&S.down_done[q], cudaEventDisableTiming));
    }
    if (ntl > 0) {
        CUDA_TRY(cudaMemsetAsync(sums_d, 0, nsums * sizeof(double), S.comp));
        CUDA_TRY(cudaStreamSynchronize(S.comp));
    }

    // planes [g0, g0 + np) of a dense caller array <-> planes [l0, ...) of an internal (padded) tile array
    auto dense_to_tile = [&](void *tile, int64_t l0, const void *host, int64_t g0, int64_t np, cudaStream_t st) -> int {
        if (np <= 0) return CYTVDN_OK;
        char *dp = (char *)tile + (size_t)l0 * plane_b;
        const char *sp_ = (const char *)host + (size_t)g0 * plane_rows * n3 * elem;
        if (!padded) { CUDA_TRY(cudaMemcpyAsync(dp, sp_, (size_t)np * plane_b, cudaMemcpyDefault, st)); return CYTVDN_OK; }
        CUDA_TRY(cudaMemsetAsync(dp, 0, (size_t)np * plane_b, st));
        CUDA_TRY(cudaMemcpy2DAsync(dp, (size_t)n3p * elem, sp_, (size_t)n3 * elem, (size_t)n3 * elem,
                                   (size_t)(np * plane_rows), cudaMemcpyDefault, st));
        return CYTVDN_OK;
    };
    auto tile_to_dense = [&](void *host, int64_t g0, const void *tile, int64_t l0, int64_t np, cudaStream_t st) -> int {
        if (np <= 0) return CYTVDN_OK;
        char *dp = (char *)host + (size_t)g0 * plane_rows * n3 * elem;
        const char *sp_ = (const char *)tile + (size_t)l0 * plane_b;
        if (!padded) { CUDA_TRY(cudaMemcpyAsync(dp, sp_, (size_t)np * plane_b, cudaMemcpyDefault, st)); return CYTVDN_OK; }
        CUDA_TRY(cudaMemcpy2DAsync(dp, (size_t)n3 * elem, sp_, (size_t)n3p * elem, (size_t)n3 * elem,
                                   (size_t)(np * plane_rows), cudaMemcpyDefault, st));
        return CYTVDN_OK;
    };

    std::vector<double> tkr(M, 0.0);
    {
        double tk = 1.0;
        for (int i = 0; i < nF; ++i) {                          // cyTVDN.py:154-156
            const double tk_new = (1.0 + std::sqrt(1.0 + 4.0 * tk * tk)) / 2.0;
            tkr[i] = (tk - 1.0) / tk_new;
            tk = tk_new;
        }
    }

    for (int pass = 0, m0 = 0; m0 < M; ++pass, m0 += (int)K) {
        const int Kp = (int)std::min<int64_t>(K, M - m0);
        const bool first = m0 == 0, last = m0 + Kp == M;
        const bool need_d_in = fista && m0 < nF;                // a FISTA iteration of this pass reads d ...
        const bool need_d_out = fista && m0 + Kp < nF;          // ... and one of a later pass will
        auto ext_lo = [&](int t) { return std::max<int64_t>(0, (int64_t)t * core - Kp); };
        auto ext_hi = [&](int t) { return std::min<int64_t>(n0, std::min<int64_t>(n0, (int64_t)(t + 1) * core) + Kp); };
        // planes of this device's last tile that lie above its core and belong to the next device
        const bool use_edge = edge_planes > 0 && !first;
        const int64_t edge_g0 = std::min<int64_t>(n0, (int64_t)t_hi * core);      // first plane of the next device

        auto upload = [&](int t) -> int {
            Slot &sl = slot[(t - t_lo) & 1];
            const int64_t e0 = ext_lo(t), e1 = ext_hi(t), np = e1 - e0;
            // leading planes that tile t-1 also held: they wait in the carry buffer (saved below, same stream)
            const int64_t have = t > t_lo ? std::min(ext_hi(t - 1), e1) - e0 : 0;
            // trailing planes of the range's last tile: from the edge snapshot, not from the (already rewritten) host
            const int64_t tail = (use_edge && t == t_hi - 1) ? e1 - edge_g0 : 0;
            const int64_t fresh = np - have - tail;             // planes read from the host state
            const size_t hb_ = (size_t)have * plane_b, rest_b = (size_t)fresh * plane_b, tail_off = (size_t)(have + fresh) * plane_b,
                         tail_b = (size_t)tail * plane_b;
            CUDA_TRY(cudaStreamWaitEvent(S.up, S.down_done[(t - t_lo) & 1], 0));      // the slot's previous tile has left
            auto d2d = [&](void *dst, const void *src, size_t bytes) -> int {
                if (bytes) CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, S.up));
                return CYTVDN_OK;
            };
            if (int rc = d2d(sl.f, carry.f, hb_)) return rc;
            if (int rc = dense_to_tile(sl.f, have, data, e0 + have, np - have, S.up)) return rc;      // the input never changes
            if (!first) {
                if (int rc = d2d(sl.r, carry.r, hb_)) return rc;
                if (int rc = dense_to_tile(sl.r, have, recon, e0 + have, fresh, S.up)) return rc;
                if (int rc = d2d((char *)sl.r + tail_off, edge.r, tail_b)) return rc;
            }
            for (int k = 0; k < nd; ++k) {
                if (first) {
                    CUDA_TRY(cudaMemsetAsync(sl.b[k], 0, (size_t)np * plane_b, S.up));
                    if (fista) CUDA_TRY(cudaMemsetAsync(sl.d[k], 0, (size_t)np * plane_b, S.up));
                } else {
                    if (int rc = d2d(sl.b[k], carry.b[k], hb_)) return rc;
                    if (rest_b)
                        CUDA_TRY(cudaMemcpyAsync((char *)sl.b[k] + hb_, (char *)hb[k] + (size_t)(e0 + have) * plane_b, rest_b,
                                                 cudaMemcpyHostToDevice, S.up));
                    if (int rc = d2d((char *)sl.b[k] + tail_off, edge.b[k], tail_b)) return rc;
                    if (need_d_in) {
                        if (int rc = d2d(sl.d[k], carry.d[k], hb_)) return rc;
                        if (rest_b)
                            CUDA_TRY(cudaMemcpyAsync((char *)sl.d[k] + hb_, (char *)hd[k] + (size_t)(e0 + have) * plane_b, rest_b,
                                                     cudaMemcpyHostToDevice, S.up));
                        if (int rc = d2d((char *)sl.d[k] + tail_off, edge.d[k], tail_b)) return rc;
                    }
                }
            }
            if (t + 1 < t_hi) {     // save what tile t+1 shares with this one before the iterations change it
                const int64_t s0 = ext_lo(t + 1) - e0, cnt = std::min(e1, ext_hi(t + 1)) - ext_lo(t + 1);
                const size_t off = (size_t)s0 * plane_b, cb = (size_t)cnt * plane_b;
                if (int rc = d2d(carry.f, (char *)sl.f + off, cb)) return rc;
                if (!first) {
                    if (int rc = d2d(carry.r, (char *)sl.r + off, cb)) return rc;
                    for (int k = 0; k < nd; ++k) {
                        if (int rc = d2d(carry.b[k], (char *)sl.b[k] + off, cb)) return rc;
                        if (need_d_in) if (int rc = d2d(carry.d[k], (char *)sl.d[k] + off, cb)) return rc;
                    }
                }
            }
            CUDA_TRY(cudaEventRecord(S.up_done[(t - t_lo) & 1], S.up));
            return CYTVDN_OK;
        };
        // the K planes above this device's range, as they are at the START of the pass (state m0)
        auto snapshot_edge = [&]() -> int {
            if (!use_edge) return CYTVDN_OK;
            const int64_t cnt = std::min<int64_t>(n0, edge_g0 + Kp) - edge_g0;
            if (int rc = dense_to_tile(edge.r, 0, recon, edge_g0, cnt, S.up)) return rc;
            for (int k = 0; k < nd; ++k) {
                CUDA_TRY(cudaMemcpyAsync(edge.b[k], (char *)hb[k] + (size_t)edge_g0 * plane_b, (size_t)cnt * plane_b,
                                         cudaMemcpyHostToDevice, S.up));
                if (need_d_in)
                    CUDA_TRY(cudaMemcpyAsync(edge.d[k], (char *)hd[k] + (size_t)edge_g0 * plane_b, (size_t)cnt * plane_b,
                                             cudaMemcpyHostToDevice, S.up));
            }
            return CYTVDN_OK;
        };
        auto compute = [&](int t) -> int {
            Slot &sl = slot[(t - t_lo) & 1];
            const int64_t e0 = ext_lo(t), e1 = ext_hi(t), np = e1 - e0;
            const int64_t c0 = (int64_t)t * core, c1 = std::min<int64_t>(n0, c0 + core);
            int64_t shape[4];
            for (int k = 0; k < nd; ++k) shape[k] = p->shape[k];
            shape[0] = np;
            CUDA_TRY(cudaStreamWaitEvent(S.comp, S.up_done[(t - t_lo) & 1], 0));
            for (int k = 0; k < Kp; ++k) {
                const int m = m0 + k;
                cytvdn_step_opts o;
                memset(&o, 0, sizeof o);
                o.row_pitch = n3p;
                o.box_lo[0] = e0 > 0 ? k + 1 : 0;
                o.box_hi[0] = e1 < n0 ? np - (k + 1) : np;
                o.own_lo[0] = c0 - e0;
                o.own_hi[0] = c1 - e0;
                if (e1 == n0 && e0 > 0 && p->bc_mode == 2) o.zero_wrap_mask = 1;   // plane 0 of b_0 is identically 0 under Jia-Zhao
                const bool fi = m < nF;
                const void *uin = (first && k == 0) ? sl.f : sl.r;      // recon = datacube.copy(), cyTVDN.py:145
                double *sm = sums_d + ((size_t)m * ntl + (t - t_lo)) * 4;
                cytvdn_step_opts oa = o;                                // A: one plane more at the upper end
                if (e1 < n0) oa.box_hi[0] = o.box_hi[0] + 1;
                oa.zero_wrap_mask = 0;
                if (int rc = cytvdn_accumulator_update_all(nd, shape, p->dtype, uin, sl.b, fi ? sl.d : nullptr, tkr[m], p->clip,
                                                           p->isotropic_R, p->isotropic_Q, p->bc_mode, sm, &oa, S.comp))
                    return rc;
                if (int rc = cytvdn_datacube_update(nd, shape, p->dtype, sl.f, uin, sl.r, sl.b, p->lambda_mu, p->bc_mode,
                                                    sm + 1, &o, S.comp))
                    return rc;
            }
            CUDA_TRY(cudaEventRecord(S.comp_done[(t - t_lo) & 1], S.comp));
            return CYTVDN_OK;
        };
        auto download = [&](int t) -> int {
            Slot &sl = slot[(t - t_lo) & 1];
            const int64_t e0 = ext_lo(t);
            const int64_t c0 = (int64_t)t * core, c1 = std::min<int64_t>(n0, c0 + core);
            CUDA_TRY(cudaStreamWaitEvent(S.down, S.comp_done[(t - t_lo) & 1], 0));
            if (int rc = tile_to_dense(recon, c0, sl.r, c0 - e0, c1 - c0, S.down)) return rc;
            if (!last)
                for (int k = 0; k < nd; ++k) {
                    CUDA_TRY(cudaMemcpyAsync((char *)hb[k] + (size_t)c0 * plane_b, (char *)sl.b[k] + (size_t)(c0 - e0) * plane_b,
                                             (size_t)(c1 - c0) * plane_b, cudaMemcpyDeviceToHost, S.down));
                    if (need_d_out)
                        CUDA_TRY(cudaMemcpyAsync((char *)hd[k] + (size_t)c0 * plane_b, (char *)sl.d[k] + (size_t)(c0 - e0) * plane_b,
                                                 (size_t)(c1 - c0) * plane_b, cudaMemcpyDeviceToHost, S.down));
                }
            CUDA_TRY(cudaEventRecord(S.down_done[(t - t_lo) & 1], S.down));
            return CYTVDN_OK;
        };

        if (ntl > 0) {
            if (int rc = snapshot_edge()) return rc;           // first: a device with ONE tile consumes it in upload(t_lo)
            if (int rc = upload(t_lo)) return rc;
        }
        if (multi && !first) {
            // every device has read what it needs from the neighbouring ranges (the K planes below its first tile, in
            // upload(t_lo), and the K planes above its last one) before anybody writes this pass's results back
            CUDA_TRY(cudaStreamSynchronize(S.up));
            if (!sh.wait()) return fail(CYTVDN_E_CUDA, "another device of the out-of-core run failed");
        }
        for (int t = t_lo; t < t_hi; ++t) {
            if (t + 1 < t_hi) if (int rc = upload(t + 1)) return rc;
            if (int rc = compute(t)) return rc;
            if (int rc = download(t)) return rc;
        }
        // the next pass reads what this one wrote.  (Letting the passes run into each other -- per-tile events instead
        // of this barrier -- was measured: no gain, the bus is busy either way.)
        CUDA_TRY(cudaStreamSynchronize(S.down));
        CUDA_TRY(cudaStreamSynchronize(S.up));
        if (multi && !last && !sh.wait()) return fail(CYTVDN_E_CUDA, "another device of the out-of-core run failed");
    }
    if (ntl > 0) {
        std::vector<double> loc(nsums, 0.0);
        CUDA_TRY(cudaMemcpy(loc.data(), sums_d, nsums * sizeof(double), cudaMemcpyDeviceToHost));
        for (int i = 0; i < M; ++i)
            for (int t = t_lo; t < t_hi; ++t)
                for (int q = 0; q < 4; ++q) sh.sums_h[((size_t)i * nt + t) * 4 + q] = loc[((size_t)i * ntl + (t - t_lo)) * 4 + q];
    }
    return CYTVDN_OK;
}

// host state between passes (internal, padded layout); recon's host state is the caller's array
int stream_host_state(const cytvdn_denoise_params *p, const Dims &D, StreamShared &sh, HostPinned &hostmem)
{
    const int nd = p->ndim, nF = p->iters_fista;
    const size_t elem = p->dtype == CYTVDN_F32 ? 4 : 8;
    const int64_t full_vw = 16 / (int64_t)elem, n3p = (D.n[3] + full_vw - 1) / full_vw * full_vw;
    const size_t plane_b = (size_t)D.n[1] * D.n[2] * n3p * elem;
    if (sh.sp.passes > 1)
        for (int k = 0; k < nd; ++k) {
            if (int rc = hostmem.alloc(&sh.hb[k], (size_t)D.n[0] * plane_b)) return rc;
            if (nF > 0 && nF > sh.sp.iters_per_pass) if (int rc = hostmem.alloc(&sh.hd[k], (size_t)D.n[0] * plane_b)) return rc;
        }
    sh.sums_h.assign((size_t)(p->iters_fista + p->iters_plain) * sh.sp.tiles * 4, 0.0);
    return CYTVDN_OK;
}

void stream_finish(const cytvdn_denoise_params *p, const StreamShared &sh, double *bnorm, double *delta)
{
    const int M = p->iters_fista + p->iters_plain, nt = (int)sh.sp.tiles;
    for (int i = 0; i < M; ++i) {
        double s3[3] = {0.0, 0.0, 0.0};
        for (int t = 0; t < nt; ++t)                            // fixed order: deterministic
            for (int q = 0; q < 3; ++q) s3[q] += sh.sums_h[((size_t)i * nt + t) * 4 + q];
        bnorm[i] = s3[0];
        delta[i] = s3[1] / s3[2];
    }
}

int denoise_streamed(const cytvdn_denoise_params *p, const Dims &D, const void *data, void *recon, double *bnorm,
                     double *delta, int32_t *iters_done, double *timing_ms, size_t budget)
{
    const auto t_start = std::chrono::steady_clock::now();
    auto ms_since = [&](std::chrono::steady_clock::time_point t0) {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    };
    StreamShared sh;
    if (int rc = make_stream_plan(p, D, budget, &sh.sp)) return rc;
    HostPinned hostmem;
    if (int rc = stream_host_state(p, D, sh, hostmem)) return rc;
    const double setup_ms = ms_since(t_start);
    const auto t_loop = std::chrono::steady_clock::now();
    if (int rc = stream_worker(p, D, data, recon, sh, 0, (int)sh.sp.tiles, (cudaStream_t)p->stream)) return rc;
    const double loop_ms = ms_since(t_loop);
    const auto t_fin = std::chrono::steady_clock::now();
    stream_finish(p, sh, bnorm, delta);
    if (iters_done) { iters_done[0] = p->iters_fista; iters_done[1] = p->iters_plain; iters_done[2] = 3 | ((int)sh.sp.tiles << 8); }
    if (timing_ms) { timing_ms[0] = setup_ms; timing_ms[1] = loop_ms; timing_ms[2] = ms_since(t_fin); }
    return CYTVDN_OK;
}
}  // namespace

// Sharded AND out of core (include/cytvdn_b200.h): the tiles of every pass dealt to `ndev` devices, one host thread each.
int cytvdn_denoise_sharded_streamed(const cytvdn_denoise_params *p, int ndev, const int *devices, const void *data, void *recon,
                                    double *bnorm, double *delta, int32_t *iters_done, double *timing_ms)
{
    Dims D;
    if (int rc = validate_params(p, &D)) return rc;
    if (!data || !recon || data == recon) return fail(CYTVDN_E_INVALID, "data / recon is NULL or aliased");
    if (ndev < 1 || ndev > 64) return fail(CYTVDN_E_INVALID, "ndev must be in 1..64");
    const int M = p->iters_fista + p->iters_plain;
    if (M <= 0 || !bnorm || !delta) return fail(CYTVDN_E_INVALID, "the out-of-core schedule needs iterations and bnorm / delta");
    if ((p->bc_mode != 2 && p->bc_mode != 3) || p->use_stopping || is_device_ptr(data) || is_device_ptr(recon))
        return fail(CYTVDN_E_UNSUPPORTED, "the out-of-core schedule needs host arrays in and out, BC_mode 2 or 3 and no stopping test");
    const auto t_start = std::chrono::steady_clock::now();
    auto ms_since = [&](std::chrono::steady_clock::time_point t0) {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    };
    int prev_dev = 0;
    CUDA_TRY(cudaGetDevice(&prev_dev));
    // one geometry for all devices: the smallest budget decides (CYTVDN_STREAM_BUDGET_MB: per device, for tests)
    size_t budget = 0;
    { const char *env = getenv("CYTVDN_STREAM_BUDGET_MB"); if (env && atof(env) > 0) budget = (size_t)(atof(env) * 1048576.0); }
    std::vector<int> devs(ndev);
    for (int r = 0; r < ndev; ++r) devs[r] = devices ? devices[r] : r;
    if (!budget) {
        // devices that appear more than once (tests: several shards on one GPU) share that GPU's memory
        for (int r = 0; r < ndev; ++r) {
            CUDA_TRY(cudaSetDevice(devs[r]));
            size_t free_b = 0, tot_b = 0;
            CUDA_TRY(cudaMemGetInfo(&free_b, &tot_b));
            int share = 0;
            for (int q = 0; q < ndev; ++q) share += devs[q] == devs[r];
            size_t b = (free_b > ((size_t)1 << 30) ? free_b - ((size_t)1 << 30) : free_b / 2) / (size_t)share;
            if (!budget || b < budget) budget = b;
        }
        CUDA_TRY(cudaSetDevice(prev_dev));
    }
    StreamShared sh;
    sh.ndev = ndev;
    if (int rc = make_stream_plan(p, D, budget, &sh.sp, ndev)) return rc;
    HostPinned hostmem;
    if (int rc = stream_host_state(p, D, sh, hostmem)) return rc;
    const double setup_ms = ms_since(t_start);
    const auto t_loop = std::chrono::steady_clock::now();
    const int nt = (int)sh.sp.tiles, per = (nt + ndev - 1) / ndev;
    std::vector<int> rcs(ndev, CYTVDN_OK);
    std::vector<std::string> msgs(ndev);
    std::vector<std::thread> th;
    for (int r = 0; r < ndev; ++r)
        th.emplace_back([&, r] {
            int rc = CYTVDN_OK;
            if (cudaSetDevice(devs[r]) != cudaSuccess) rc = fail(CYTVDN_E_CUDA, "cudaSetDevice(%d) failed", devs[r]);
            if (!rc) rc = stream_worker(p, D, data, recon, sh, std::min(nt, r * per), std::min(nt, (r + 1) * per), nullptr);
            if (rc) { msgs[r] = g_err; sh.fail_all(); }
            rcs[r] = rc;
        });
    for (auto &t : th) t.join();
    CUDA_TRY(cudaSetDevice(prev_dev));
    for (int r = 0; r < ndev; ++r)
        if (rcs[r] && msgs[r].find("another device") == std::string::npos) return fail(rcs[r], "device %d: %s", devs[r], msgs[r].c_str());
    for (int r = 0; r < ndev; ++r) if (rcs[r]) return fail(rcs[r], "device %d: %s", devs[r], msgs[r].c_str());
    const double loop_ms = ms_since(t_loop);
    const auto t_fin = std::chrono::steady_clock::now();
    stream_finish(p, sh, bnorm, delta);
    if (iters_done) { iters_done[0] = p->iters_fista; iters_done[1] = p->iters_plain; iters_done[2] = 3 | (ndev << 8); }
    if (timing_ms) { timing_ms[0] = setup_ms; timing_ms[1] = loop_ms; timing_ms[2] = ms_since(t_fin); }
    return CYTVDN_OK;
}

int cytvdn_denoise(const cytvdn_denoise_params *p, const void *data, void *recon, const void *reference_data,
                   double *bnorm, double *delta, double *mse, int32_t *iters_done, double *timing_ms)
{
    Dims D;
    if (int rc = validate_params(p, &D)) return rc;
    if (!data || !recon) return fail(CYTVDN_E_INVALID, "data / recon is NULL");
    // CYTVDN_TRACE=1: host-clock milestones of the call on stderr (where the wall time outside the CUDA events goes)
    const bool trace = [] { const char *e = getenv("CYTVDN_TRACE"); return e && *e && strcmp(e, "0"); }();
    const auto t_start = std::chrono::steady_clock::now();
    g_trace_n = 0;
    auto mark = [&](const char *what) {
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count();
        if (g_trace_n < kTraceMax) { g_trace_ms[g_trace_n] = ms; g_trace_what[g_trace_n] = what; ++g_trace_n; }
        if (trace) fprintf(stderr, "[cytvdn_denoise] %8.2f ms  %s\n", ms, what);
    };
    if (data == recon) return fail(CYTVDN_E_INVALID, "recon must not alias data (the input is read in every iteration)");
    const int nF = p->iters_fista, nU = p->iters_plain, nIt = nF + nU;
    if (nIt > 0 && (!bnorm || !delta)) return fail(CYTVDN_E_INVALID, "bnorm / delta is NULL");
    if (reference_data && !mse) return fail(CYTVDN_E_INVALID, "mse is NULL but reference_data was given");
    const size_t elem = p->dtype == CYTVDN_F32 ? 4 : 8;
    const int nd = p->ndim;
    // Rows whose length is not a multiple of the 16-byte vector width are PADDED in the internal state (pitch
    // rounded up), so every shape runs on the vector path; the pad voxels are inert (cytvdn_step_opts.row_pitch).
    // Caller arrays stay dense: they are converted by 2-D copies on the way in and out.
    const int64_t full_vw = 16 / (int64_t)elem;
    bool padded = D.n[3] % full_vw != 0;
    { const char *env = getenv("CYTVDN_PAD_ROWS"); if (env && !strcmp(env, "0")) padded = false; }
    const int64_t n3 = D.n[3], n3p = padded ? (n3 + full_vw - 1) / full_vw * full_vw : n3;
    const int64_t rows = D.n[0] * D.n[1] * D.n[2];
    const int64_t nvox = rows * n3p;                    // voxels of one internal array (pads included)
    const size_t nb = (size_t)nvox * elem;
    cytvdn_step_opts sopts;
    memset(&sopts, 0, sizeof sopts);
    sopts.row_pitch = n3p;
    sopts.flags = (p->isotropic_R ? 16 : 0) | (p->isotropic_Q ? 32 : 0);     // read by the fused iteration only

    // residency of the caller's arrays (decides which schedules apply) vs "can be used in place" (dense rows only)
    const bool data_on_dev = is_device_ptr(data), recon_on_dev = is_device_ptr(recon);
    const bool data_dev = data_on_dev && !padded, recon_dev = recon_on_dev && !padded;
    const bool ref_dev = reference_data ? (is_device_ptr(reference_data) && !padded) : false;
    int prev_dev = -1;
    CUDA_TRY(cudaGetDevice(&prev_dev));
    if (!data_on_dev && p->device >= 0) CUDA_TRY(cudaSetDevice(p->device));
    struct Restore {
        int d;
        void now() { int c = -1; if (d >= 0 && (cudaGetDevice(&c) != cudaSuccess || c != d)) cudaSetDevice(d); d = -1; }
        ~Restore() { now(); }
    } restore{prev_dev};
    cudaStream_t st = (cudaStream_t)p->stream;

    // ---- out of core: host arrays whose state does not fit in HBM (or schedule 3 / CYTVDN_STREAM_BUDGET_MB) ----
    {
        bool want = requested_schedule(p) == 3;
        size_t budget = 0;
        { const char *env = getenv("CYTVDN_STREAM_BUDGET_MB"); if (env && atof(env) > 0) { budget = (size_t)(atof(env) * 1048576.0); want = true; } }
        const bool can = !data_on_dev && !recon_on_dev && (p->bc_mode == 2 || p->bc_mode == 3) && !p->use_stopping &&
                         !reference_data && nIt > 0;
        size_t free_b = 0;
        // (cudaMemGetInfo costs 15 - 20 ms on a B200 with tens of GB allocated: asked only when the answer matters)
        if (want || (can && requested_schedule(p) == 0)) if (int rc = available_bytes(&free_b)) return rc;
        if (!want && can && requested_schedule(p) == 0) {
            const int64_t in_core = arrays_needed(p, false, false, false, false) * (int64_t)nb;     // two-pass, in place
            want = in_core + (int64_t)(512ll << 20) > (int64_t)free_b;
        }
        if (want) {
            if (!can)
                return fail(CYTVDN_E_UNSUPPORTED, "the out-of-core schedule needs host arrays in and out, BC_mode 2 or 3, at "
                                                  "least one iteration, no stopping test and no reference_data");
            if (!budget) budget = free_b > ((size_t)1 << 30) ? free_b - ((size_t)1 << 30) : free_b / 2;
            return denoise_streamed(p, D, data, recon, bnorm, delta, iters_done, timing_ms, budget);
        }
    }

    // ---- schedule: fused single pass when it applies and the second state set fits ----------------
    bool fused = false;
    {
        const int want = requested_schedule(p);
        if (want == 2 && !fused_possible(p))
            return fail(CYTVDN_E_INVALID, "the fused schedule does not cover this boundary mode; use schedule 0 or 1");
        // The pair (0,1) costs the fused kernel three joint shrinks per voxel instead of one (the forward
        // neighbours on both far axes are recomputed): measured on config 4, 20.6 ms fused against 16.5 ms in two
        // passes; (2,3) alone is a wash (16.1 / 16.3 ms).  Auto therefore runs isotropic_R in two passes.
        const bool fused_pays = !p->isotropic_R;
        // the mirror variant of the fused kernel exists for the full vector width only
        const bool mirror_ok = p->bc_mode != 3 || (n3p % full_vw == 0 &&
                                                   (!data_dev || (reinterpret_cast<uintptr_t>(data) & 15u) == 0) &&
                                                   (!recon_dev || (reinterpret_cast<uintptr_t>(recon) & 15u) == 0));
        if (want == 2 && !mirror_ok)
            return fail(CYTVDN_E_UNSUPPORTED, "BC_mode=3 on the fused schedule needs 16-byte aligned rows and arrays");
        if (want == 2 && nIt > 0) fused = true;                 // asked for: an allocation failure is reported, not hidden
        else if (want != 1 && fused_possible(p) && fused_pays && mirror_ok && nIt > 0) {
            size_t free_b = 0;
            if (int rc = available_bytes(&free_b)) return rc;
            const int64_t need = arrays_needed(p, true, data_dev, recon_dev, reference_data && !ref_dev) * (int64_t)nb;
            fused = need + (int64_t)(512ll << 20) <= (int64_t)free_b;
        }
    }

    // ---- pipelining over PCIe (host arrays only) ----------------------------------------------------
    // The array is cut into `nbox` boxes of axis-0 planes.  The first and the last `depth` iterations run box by
    // box in wavefront order (box c may be at most one iteration ahead of box c+1, which is all the stencil's
    // dependence cone asks for): box c starts iterating as soon as boxes c and c+1 have arrived, and is copied
    // back as soon as ITS last iteration is done while the boxes behind it still iterate.  One host->device copy
    // costs ~6.7 iterations (4 B/voxel at ~55 GB/s against 76 B/voxel at ~7 TB/s), so 16 boxes hide it.
    // Needs the Jia-Zhao boundary (under the periodic one box 0 depends on the last box), fixed iteration
    // counts (no stopping test, no per-iteration MSE: both need the whole array at one iteration).
    // The last box takes its wrap term as 0 (zero_wrap): plane 0 of b_0 is identically 0 under Jia-Zhao.
    int nbox = 1;
    // sum (reference - recon')^2 can ride along in the pass that produces recon' when that kernel has the variant:
    // full vector width (rows and caller pointers 16-byte aligned), and not the fused half-isotropic kernel
    auto aligned16 = [](const void *q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    const bool sse_in_pass = reference_data && !(fused && (p->isotropic_R || p->isotropic_Q || p->bc_mode == 3)) && n3p % full_vw == 0 &&
                             (!data_dev || aligned16(data)) && (!recon_dev || aligned16(recon)) &&
                             (!ref_dev || aligned16(reference_data));
    {
        const bool can = fused && (p->bc_mode == 2 || p->bc_mode == 3) && !p->use_stopping && (!reference_data || sse_in_pass) && nIt > 0 &&
                         (!data_on_dev || !recon_on_dev);
        int want = nb >= ((size_t)256 << 20) ? 16 : 1;
        const char *env = getenv("CYTVDN_PIPELINE");
        if (env && *env) want = atoi(env);
        if (can && want > 1) nbox = (int)std::min<int64_t>(want, D.n[0] / 2);       // at least two planes per box
        if (nbox < 2) nbox = 1;
    }
    const bool pipe = nbox > 1;
    auto box_lo = [&](int c) -> int64_t { return D.n[0] * c / nbox; };
    const int64_t plane_rows = D.n[1] * D.n[2];

    cudaEvent_t ev[4];
    for (auto &e : ev) CUDA_TRY(cudaEventCreate(&e));
    struct EvFree {
        cudaEvent_t *e;
        void now() { if (e) for (int k = 0; k < 4; ++k) cudaEventDestroy(e[k]); e = nullptr; }
        ~EvFree() { now(); }
    } evfree{ev};
    CUDA_TRY(cudaEventRecord(ev[0], st));
    mark("entry checks done");
    struct Side {                                   // copy stream + per-box events of the pipelined path
        cudaStream_t s = nullptr;
        std::vector<cudaEvent_t> up, down;
        void now()
        {
            for (auto e : up) cudaEventDestroy(e);
            for (auto e : down) cudaEventDestroy(e);
            if (s) cudaStreamDestroy(s);
            up.clear(); down.clear(); s = nullptr;
        }
        ~Side() { now(); }
    } side;
    if (pipe) {
        CUDA_TRY(cudaStreamCreateWithFlags(&side.s, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamWaitEvent(side.s, ev[0], 0));
        auto mk = [&](std::vector<cudaEvent_t> &v) -> int {
            v.reserve(nbox);
            for (int c = 0; c < nbox; ++c) { cudaEvent_t e; CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); v.push_back(e); }
            return CYTVDN_OK;
        };
        if (!data_dev) if (int rc = mk(side.up)) return rc;
        if (!recon_dev) if (int rc = mk(side.down)) return rc;
    }

    const size_t nsums = (size_t)(nIt + 1) * 4 * (size_t)nbox;
    void *b[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}}, *d[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
    void *orig_d = nullptr, *rbuf[2] = {nullptr, nullptr}, *ref_d = nullptr;
    double *sums_d = nullptr;
    // dense caller array -> internal array, rows [r0, r0 + nr) (any direction the pointers imply; pads zeroed)
    auto copy_in = [&](void *dst, const void *src, int64_t r0, int64_t nr, cudaStream_t s) -> int {
        char *dp = (char *)dst + (size_t)r0 * n3p * elem;
        const char *sp = (const char *)src + (size_t)r0 * n3 * elem;
        if (!padded) { CUDA_TRY(cudaMemcpyAsync(dp, sp, (size_t)nr * n3 * elem, cudaMemcpyDefault, s)); return CYTVDN_OK; }
        CUDA_TRY(cudaMemsetAsync(dp, 0, (size_t)nr * n3p * elem, s));
        CUDA_TRY(cudaMemcpy2DAsync(dp, (size_t)n3p * elem, sp, (size_t)n3 * elem, (size_t)n3 * elem, (size_t)nr,
                                   cudaMemcpyDefault, s));
        return CYTVDN_OK;
    };
    auto copy_out = [&](void *dst, const void *src, int64_t r0, int64_t nr, cudaStream_t s) -> int {
        char *dp = (char *)dst + (size_t)r0 * n3 * elem;
        const char *sp = (const char *)src + (size_t)r0 * n3p * elem;
        if (!padded) { CUDA_TRY(cudaMemcpyAsync(dp, sp, (size_t)nr * n3 * elem, cudaMemcpyDefault, s)); return CYTVDN_OK; }
        CUDA_TRY(cudaMemcpy2DAsync(dp, (size_t)n3 * elem, sp, (size_t)n3p * elem, (size_t)n3 * elem, (size_t)nr,
                                   cudaMemcpyDefault, s));
        return CYTVDN_OK;
    };
    auto upload_box = [&](int c) -> int {
        const int64_t r0 = box_lo(c) * plane_rows, r1 = box_lo(c + 1) * plane_rows;
        if (int rc = copy_in(orig_d, data, r0, r1 - r0, side.s)) return rc;
        CUDA_TRY(cudaEventRecord(side.up[c], side.s));
        return CYTVDN_OK;
    };
    // pinned (or registered) host memory: asynchronous copies, all enqueued up front; pageable memory: a copy
    // blocks the host while it is staged, so the boxes are interleaved with the kernel launches instead
    auto host_async = [&](const void *q) -> bool {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, q) != cudaSuccess) { cudaGetLastError(); return false; }
        return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
    };
    int uploaded = 0;                                // boxes whose upload has been enqueued

    // The input gets its own allocation so that its upload can start before (and run while) the big arena of
    // the state arrays is being allocated (~30 ms for 80 GB).
    Arena pool0, pool;
    if (data_dev) orig_d = const_cast<void *>(data);
    else {
        if (int rc = pool0.reserve(Arena::padded(nb))) return rc;
        if (int rc = pool0.alloc(&orig_d, nb)) return rc;
        if (!pipe) { if (int rc = copy_in(orig_d, data, 0, rows, st)) return rc; }
        else if (host_async(data)) { for (; uploaded < nbox; ++uploaded) if (int rc = upload_box(uploaded)) return rc; }
    }
    mark("input allocated, upload enqueued");
    {
        const int64_t arrays = arrays_needed(p, fused, true, recon_dev, reference_data && !ref_dev);
        if (int rc = pool.reserve((size_t)arrays * Arena::padded(nb) + Arena::padded(nsums * sizeof(double)) + 4096))
            return rc;
    }
    mark("state arena allocated");
    if (recon_dev) rbuf[0] = recon;
    else if (int rc = pool.alloc(&rbuf[0], nb)) return rc;
    if (fused) if (int rc = pool.alloc(&rbuf[1], nb)) return rc;
    if (reference_data) {
        if (ref_dev) ref_d = const_cast<void *>(reference_data);
        else { if (int rc = pool.alloc(&ref_d, nb)) return rc; if (int rc = copy_in(ref_d, reference_data, 0, rows, st)) return rc; }
    }
    for (int k = 0; k < nd && nIt > 0; ++k) {
        for (int s = 0; s < (fused ? 2 : 1); ++s) {
            if (int rc = pool.alloc(&b[s][k], nb)) return rc;
            if (nF > 0) if (int rc = pool.alloc(&d[s][k], nb)) return rc;
        }
        CUDA_TRY(cudaMemsetAsync(b[0][k], 0, nb, st));            // only the set that is read first
        if (nF > 0) CUDA_TRY(cudaMemsetAsync(d[0][k], 0, nb, st));
    }
    // per iteration (and box): [0] sum|b|, [1] sum|delta|, [2] sum|old|, [3] sse ; slot nIt holds MSE[0]
    if (int rc = pool.alloc((void **)&sums_d, nsums * sizeof(double))) return rc;
    CUDA_TRY(cudaMemsetAsync(sums_d, 0, nsums * sizeof(double), st));
    std::vector<double> sums_h(nsums, 0.0);
    // early stopping: the per-iteration sums come back through a small page-locked buffer (two slots) that the
    // thread keeps for its lifetime (cudaMallocHost costs ~1 ms), one event per slot
    static thread_local double *pinned = nullptr;
    cudaEvent_t stop_ev[2] = {nullptr, nullptr};
    struct StopEvFree {
        cudaEvent_t *e;
        void now() { for (int k = 0; k < 2; ++k) if (e[k]) { cudaEventDestroy(e[k]); e[k] = nullptr; } }
        ~StopEvFree() { now(); }
    } stopfree{stop_ev};
    if (p->use_stopping) {
        if (!pinned) CUDA_TRY(cudaMallocHost(&pinned, 8 * sizeof(double)));
        for (auto &e : stop_ev) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    auto below = [&](const double *s4) {                       // cyTVDN.py:189-194 / :236-242
        const double dl = s4[1] / s4[2];
        const double dl_t = p->dtype == CYTVDN_F32 ? (double)(float)dl : dl;       // stored in the array dtype
        return dl_t < p->stopping_relative_change;
    };
    // Fused schedule: iteration i+1 never touches what iteration i produced (ping-pong), so from iteration
    // kSpeculateFrom on it is launched BEFORE the host has seen delta[i]; if delta[i] turns out to be below the
    // threshold the speculative iteration is discarded -- same result as "stop right after iteration i"
    // (cyTVDN.py:189-194) without draining the stream once per iteration.  Early iterations, where a stop is
    // likely and a discarded sweep costs more than a bubble, and the in-place two-pass schedule test synchronously.
    constexpr int kSpeculateFrom = 8;

    auto sse = [&](const void *a, const void *b, double *out) -> int {
        return p->dtype == CYTVDN_F32 ? run_sse<float>(nvox, a, b, out, st, n3, n3p)
                                      : run_sse<double>(nvox, a, b, out, st, n3, n3p);
    };
    // MSE[i+1] = sum (reference - recon_{i+1})^2 rides along in the pass that produces recon_{i+1} (fused kernel /
    // half-step B) unless that kernel has no such variant (fused half-isotropic): then it is a sweep of its own
    if (sse_in_pass) sopts.sse_reference = ref_d;

    CUDA_TRY(cudaEventRecord(ev[1], st));
    // iteration 0 reads the reconstruction straight from the input (recon = datacube.copy(),
    // cyTVDN.py:145), which saves the device-to-device copy.
    const void *u_cur = orig_d;
    int cur = 0;                 // fused: state set that holds the current b/d ; recon buffer to write next
    int rnext = 0;
    double tk = 1.0;
    int done[2] = {0, 0};
    std::vector<char> ran(nIt > 0 ? nIt : 1, 0);
    if (!pipe) {
    for (int phase = 0; phase < 2; ++phase) {
        const int n = phase == 0 ? nF : nU;
        for (int it = 0; it < n; ++it) {
            const int i = phase == 0 ? it : nF + it;
            double tkr = 0.0;
            if (phase == 0) {                                   // cyTVDN.py:154-156
                const double tk_new = (1.0 + std::sqrt(1.0 + 4.0 * tk * tk)) / 2.0;
                tkr = (tk - 1.0) / tk_new;
                tk = tk_new;
            }
            double *s = sums_d + (size_t)i * 4;
            void *u_out = rbuf[rnext];
            const void *u_before = u_cur;
            if (fused) {
                if (int rc = cytvdn_fused_iteration(nd, p->shape, p->dtype, orig_d, u_cur, u_out, b[cur], b[1 - cur],
                                                    phase == 0 ? d[cur] : nullptr, phase == 0 ? d[1 - cur] : nullptr,
                                                    tkr, p->clip, p->lambda_mu, p->bc_mode, s, &sopts, st))
                    return rc;
                cur = 1 - cur;
                rnext = 1 - rnext;
            } else {
                if (int rc = cytvdn_accumulator_update_all(nd, p->shape, p->dtype, u_cur, b[0], phase == 0 ? d[0] : nullptr,
                                                           tkr, p->clip, p->isotropic_R, p->isotropic_Q, p->bc_mode, s,
                                                           &sopts, st))
                    return rc;
                if (int rc = cytvdn_datacube_update(nd, p->shape, p->dtype, orig_d, u_cur, u_out, b[0], p->lambda_mu,
                                                    p->bc_mode, s + 1, &sopts, st))
                    return rc;
            }
            u_cur = u_out;
            if (reference_data && !sse_in_pass)
                if (int rc = sse(ref_d, u_cur, s + 3)) return rc;
            ran[i] = 1;
            ++done[phase];
            if (p->use_stopping) {
                const int slot = it & 1;
                CUDA_TRY(cudaMemcpyAsync(pinned + 4 * slot, s, 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
                CUDA_TRY(cudaEventRecord(stop_ev[slot], st));
                if (fused && it > kSpeculateFrom) {             // this launch was speculative: now look at iteration i-1
                    CUDA_TRY(cudaEventSynchronize(stop_ev[slot ^ 1]));
                    if (below(pinned + 4 * (slot ^ 1))) {       // it met the criterion: iteration i never happened
                        ran[i] = 0;
                        --done[phase];
                        cur = 1 - cur;
                        rnext = 1 - rnext;
                        u_cur = u_before;
                        break;
                    }
                } else if (!(fused && it == kSpeculateFrom)) {  // (at kSpeculateFrom the test moves one iteration back)
                    CUDA_TRY(cudaEventSynchronize(stop_ev[slot]));
                    if (below(pinned + 4 * slot)) break;
                }
            }
        }
    }
    } else {
        // ---- pipelined: wavefront over (box, iteration) at both ends, whole-array sweeps in between ----------
        std::vector<double> tkr(nIt, 0.0);
        for (int i = 0; i < nF; ++i) {                          // cyTVDN.py:154-156
            const double tk_new = (1.0 + std::sqrt(1.0 + 4.0 * tk * tk)) / 2.0;
            tkr[i] = (tk - 1.0) / tk_new;
            tk = tk_new;
        }
        // iteration m reads state set m%2 and recon buffer (m-1)%2 (the input for m = 0), writes set (m+1)%2 and
        // recon buffer m%2
        auto iterate = [&](int m, int c) -> int {               // c < 0: the whole array
            cytvdn_step_opts o = sopts;
            if (c >= 0) {
                o.box_lo[0] = box_lo(c); o.box_hi[0] = box_lo(c + 1);
                if (c == nbox - 1 && p->bc_mode == 2) o.zero_wrap_mask |= 1;    // (mirror: the last plane's term is b' - b')
            }
            const bool fi = m < nF;
            const int in = m & 1, out = in ^ 1;
            return cytvdn_fused_iteration(nd, p->shape, p->dtype, orig_d, m == 0 ? orig_d : rbuf[(m - 1) & 1], rbuf[m & 1],
                                          b[in], b[out], fi ? d[in] : nullptr, fi ? d[out] : nullptr, tkr[m], p->clip,
                                          p->lambda_mu, p->bc_mode, sums_d + ((size_t)m * nbox + (c < 0 ? 0 : c)) * 4, &o, st);
        };
        // first and last `nbox` iterations box by box in wavefront order, whole-array sweeps in between
        std::vector<std::pair<int, int>> order;
        pipeline_schedule(nbox, nIt, order);
        for (const auto &cm : order) {
            const int c = cm.first, m = cm.second;
            if (c >= 0 && m == 0 && !data_dev) {                // needs the planes of boxes c and c+1
                const int need = std::min(c + 1, nbox - 1);
                for (; uploaded <= need; ++uploaded) if (int rc = upload_box(uploaded)) return rc;
                CUDA_TRY(cudaStreamWaitEvent(st, side.up[need], 0));
            }
            if (int rc = iterate(m, c)) return rc;
            if (c >= 0 && m == nIt - 1 && !recon_dev) CUDA_TRY(cudaEventRecord(side.down[c], st));
        }
        for (int i = 0; i < nIt; ++i) ran[i] = 1;
        done[0] = nF; done[1] = nU;
        u_cur = rbuf[(nIt - 1) & 1];
        if (!recon_dev) {                                       // boxes go home as they finish
            for (int c = 0; c < nbox; ++c) {
                CUDA_TRY(cudaStreamWaitEvent(side.s, side.down[c], 0));
                const int64_t r0 = box_lo(c) * plane_rows, r1 = box_lo(c + 1) * plane_rows;
                if (int rc = copy_out(recon, u_cur, r0, r1 - r0, side.s)) return rc;
            }
            CUDA_TRY(cudaEventRecord(side.down[0], side.s));     // reuse: "all copies done"
            CUDA_TRY(cudaStreamWaitEvent(st, side.down[0], 0));
        }
    }
    if (reference_data)                        // MSE[0] = sum (datacube - reference)^2 (cyTVDN.py:122-125); the input is
        if (int rc = sse(orig_d, ref_d, sums_d + (size_t)nIt * 4 * nbox + 3)) return rc;      // complete on the device by now
    if (recon_dev && u_cur != rbuf[0])         // zero iterations, or the fused ping-pong ended in the spare buffer
        CUDA_TRY(cudaMemcpyAsync(rbuf[0], u_cur, nb, cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaEventRecord(ev[2], st));
    mark("iterations enqueued");

    if (!recon_dev && !pipe) if (int rc = copy_out(recon, u_cur, 0, rows, st)) return rc;
    CUDA_TRY(cudaMemcpyAsync(sums_h.data(), sums_d, nsums * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    mark("device work complete");
    for (int i = 0; i < nIt; ++i) {
        double s4[4] = {0.0, 0.0, 0.0, 0.0};
        for (int c = 0; c < nbox; ++c)                          // fixed order: deterministic
            for (int q = 0; q < 4; ++q) s4[q] += sums_h[((size_t)i * nbox + c) * 4 + q];
        bnorm[i] = ran[i] ? s4[0] : 0.0;
        delta[i] = ran[i] ? s4[1] / s4[2] : 0.0;
        if (mse) mse[i + 1] = ran[i] ? s4[3] : 0.0;
    }
    if (mse) mse[0] = sums_h[(size_t)nIt * 4 * nbox + 3];
    if (iters_done) { iters_done[0] = done[0]; iters_done[1] = done[1]; iters_done[2] = (fused ? 2 : 1) | (pipe ? nbox << 8 : 0); }
    CUDA_TRY(cudaEventRecord(ev[3], st));
    CUDA_TRY(cudaEventSynchronize(ev[3]));
    if (timing_ms) {
        float t = 0;
        CUDA_TRY(cudaEventElapsedTime(&t, ev[0], ev[1])); timing_ms[0] = t;
        CUDA_TRY(cudaEventElapsedTime(&t, ev[1], ev[2])); timing_ms[1] = t;
        CUDA_TRY(cudaEventElapsedTime(&t, ev[2], ev[3])); timing_ms[2] = t;
    }
    side.now();
    evfree.now();
    stopfree.now();
    restore.now();
    mark("results delivered");
    // cudaFree of the ~80 GB arena: 25-40 ms, now and then several hundred (tools/e2e_trace.py).  Handing it to a
    // helper thread was tried and bought nothing: the caller's next allocations stall for as long as the free runs.
    pool.release();
    pool0.release();
    mark("arenas freed");
    return CYTVDN_OK;
}


// ---- host-side plans (no GPU needed) ------------------------------------------------------------
int cytvdn_pipeline_schedule(int nbox, int n_iter, int32_t *box, int32_t *iter, int64_t capacity, int64_t *count)
{
    if (nbox < 1 || n_iter < 0 || !count) return fail(CYTVDN_E_INVALID, "bad argument");
    std::vector<std::pair<int, int>> order;
    pipeline_schedule(nbox, n_iter, order);
    *count = (int64_t)order.size();
    if (box && iter) {
        if (capacity < *count) return fail(CYTVDN_E_INVALID, "capacity %lld < %lld launches", (long long)capacity, (long long)*count);
        for (size_t i = 0; i < order.size(); ++i) { box[i] = order[i].first; iter[i] = order[i].second; }
    }
    return CYTVDN_OK;
}

int cytvdn_stream_plan_sharded(const cytvdn_denoise_params *p, int64_t budget_bytes, int ndev, int64_t *out8)
{
    Dims D;
    if (int rc = validate_params(p, &D)) return rc;
    if (!out8 || budget_bytes <= 0 || ndev < 1) return fail(CYTVDN_E_INVALID, "bad argument");
    StreamPlan sp;
    if (int rc = make_stream_plan(p, D, (size_t)budget_bytes, &sp, ndev)) return rc;
    out8[0] = sp.planes_per_slot; out8[1] = sp.iters_per_pass; out8[2] = sp.core_planes; out8[3] = sp.tiles;
    out8[4] = sp.passes; out8[5] = sp.arrays_per_slot; out8[6] = sp.plane_bytes; out8[7] = sp.host_state_bytes;
    return CYTVDN_OK;
}

int cytvdn_stream_plan(const cytvdn_denoise_params *p, int64_t budget_bytes, int64_t *out8)
{
    Dims D;
    if (int rc = validate_params(p, &D)) return rc;
    if (!out8 || budget_bytes <= 0) return fail(CYTVDN_E_INVALID, "bad argument");
    StreamPlan sp;
    if (int rc = make_stream_plan(p, D, (size_t)budget_bytes, &sp)) return rc;
    out8[0] = sp.planes_per_slot; out8[1] = sp.iters_per_pass; out8[2] = sp.core_planes; out8[3] = sp.tiles;
    out8[4] = sp.passes; out8[5] = sp.arrays_per_slot; out8[6] = sp.plane_bytes; out8[7] = sp.host_state_bytes;
    return CYTVDN_OK;
}

// ---- small CUDA helpers -----------------------------------------------------------------------
int cytvdn_workspace_reserve(int64_t bytes)
{
    if (bytes < 0) return fail(CYTVDN_E_INVALID, "bytes < 0");
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_res_mutex);
    Reserved &r = g_reserved[dev];
    if (r.refs > 0) return fail(CYTVDN_E_INVALID, "the reserved workspace is in use by a running call");
    if (r.base && r.size >= (size_t)bytes) return CYTVDN_OK;          // big enough already
    if (r.base) { cudaFree(r.base); r = Reserved(); }
    if (bytes == 0) return CYTVDN_OK;
    cudaError_t e = cudaMalloc((void **)&r.base, (size_t)bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        r = Reserved();
        return fail(CYTVDN_E_NOMEM, "cudaMalloc of %lld bytes failed: %s", (long long)bytes, cudaGetErrorString(e));
    }
    r.size = (size_t)bytes;
    return CYTVDN_OK;
}

int cytvdn_workspace_release(void)
{
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceSynchronize());
    {
        std::lock_guard<std::mutex> lk(g_res_mutex);
        auto it = g_reserved.find(dev);
        if (it != g_reserved.end()) {
            if (it->second.refs > 0) return fail(CYTVDN_E_INVALID, "the reserved workspace is in use by a running call");
            if (it->second.base) cudaFree(it->second.base);
            g_reserved.erase(it);
        }
    }
    drop_workspaces(dev);
    return CYTVDN_OK;
}

int cytvdn_workspace_info(int64_t *reserved_bytes, int64_t *in_use_bytes)
{
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_res_mutex);
    auto it = g_reserved.find(dev);
    if (reserved_bytes) *reserved_bytes = it == g_reserved.end() ? 0 : (int64_t)it->second.size;
    if (in_use_bytes) *in_use_bytes = it == g_reserved.end() ? 0 : (int64_t)it->second.used;
    return CYTVDN_OK;
}

int cytvdn_last_trace(double *ms, const char **what, int capacity, int *count)
{
    if (!count) return fail(CYTVDN_E_INVALID, "count is NULL");
    *count = g_trace_n;
    for (int k = 0; k < g_trace_n && k < capacity; ++k) {
        if (ms) ms[k] = g_trace_ms[k];
        if (what) what[k] = g_trace_what[k];
    }
    return CYTVDN_OK;
}

int cytvdn_malloc(void **ptr, int64_t bytes)
{
    if (!ptr || bytes < 0) return fail(CYTVDN_E_INVALID, "bad argument");
    *ptr = nullptr;
    if (bytes == 0) return CYTVDN_OK;
    CUDA_TRY(cudaMalloc(ptr, (size_t)bytes));
    return CYTVDN_OK;
}
int cytvdn_free(void *ptr) { if (ptr) CUDA_TRY(cudaFree(ptr)); return CYTVDN_OK; }
int cytvdn_host_alloc(void **ptr, int64_t bytes)
{
    if (!ptr || bytes < 0) return fail(CYTVDN_E_INVALID, "bad argument");
    *ptr = nullptr;
    if (bytes == 0) return CYTVDN_OK;
    return pinned_alloc(ptr, (size_t)bytes);
}
int cytvdn_host_free(void *ptr) { return pinned_free(ptr); }
int cytvdn_memcpy(void *dst, const void *src, int64_t bytes, void *stream)
{
    if (bytes <= 0) return CYTVDN_OK;
    CUDA_TRY(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, (cudaStream_t)stream));
    return CYTVDN_OK;
}
int cytvdn_memset(void *dst, int value, int64_t bytes, void *stream)
{
    if (bytes <= 0) return CYTVDN_OK;
    CUDA_TRY(cudaMemsetAsync(dst, value, (size_t)bytes, (cudaStream_t)stream));
    return CYTVDN_OK;
}
int cytvdn_stream_synchronize(void *stream) { CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream)); return CYTVDN_OK; }
int cytvdn_set_device(int device) { CUDA_TRY(cudaSetDevice(device)); return CYTVDN_OK; }
int cytvdn_get_device(int *device) { if (!device) return fail(CYTVDN_E_INVALID, "NULL"); CUDA_TRY(cudaGetDevice(device)); return CYTVDN_OK; }
int cytvdn_ipc_get_handle(void *ptr, unsigned char handle[64])
{
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    if (!ptr || !handle) return fail(CYTVDN_E_INVALID, "NULL argument");
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, ptr));
    memcpy(handle, &h, 64);
    return CYTVDN_OK;
}
int cytvdn_ipc_open(const unsigned char handle[64], void **peer_ptr)
{
    if (!handle || !peer_ptr) return fail(CYTVDN_E_INVALID, "NULL argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CUDA_TRY(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return CYTVDN_OK;
}
int cytvdn_ipc_close(void *peer_ptr)
{
    if (peer_ptr) CUDA_TRY(cudaIpcCloseMemHandle(peer_ptr));
    return CYTVDN_OK;
}
int cytvdn_mem_info(int64_t *free_bytes, int64_t *total_bytes)
{
    size_t f = 0, t = 0;
    CUDA_TRY(cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = (int64_t)f;
    if (total_bytes) *total_bytes = (int64_t)t;
    return CYTVDN_OK;
}

}  // extern "C"

// ---- helpers for the other translation units (internal.hh) -----------------------------------------
namespace cytvdn_internal {
int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
int warm_workspace(cudaStream_t stream) { Workspace w; return get_workspace(stream, &w); }
int pinned_alloc(void **out, size_t bytes) { return ::pinned_alloc(out, bytes); }
int pinned_free(void *p) { return ::pinned_free(p); }
}  // namespace cytvdn_internal
