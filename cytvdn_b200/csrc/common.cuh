// common.cuh -- shared device/host helpers of libcytvdn_b200 (sm_100a only).
//
// Data layout (DESIGN.md "Data layout in HBM"): every array is the caller's C-contiguous
// [N0,N1,N2,N3] block (3-D arrays are embedded as [N0,N1,1,N2]); axis 3 is unit stride.
// A sweep walks the array in "strips": for each strip of TJ consecutive axis-1 indices,
// for each axis-0 index i, the contiguous slab [i, strip, :, :] is cut into chunks of
// kBlock*VW elements, one chunk per CTA iteration.  Consecutive tiles are therefore
// consecutive in memory inside a slab, the axis-1 neighbour is one inner plane away and the
// axis-0 neighbour TJ inner planes away, i.e. both stay L2 resident (126 MB) and are fetched
// from HBM once.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cytvdn {

#ifndef CYTVDN_BLOCK
#define CYTVDN_BLOCK 256
#endif
constexpr int kBlock = CYTVDN_BLOCK;   // threads per CTA (= vectors per tile)
constexpr int kWarps = kBlock / 32;

// ---- division of n < 2^31 by a runtime constant ------------------------------------------
struct FastDiv {
    uint32_t d, mul, shr;
};
inline FastDiv make_fastdiv(uint32_t d)
{
    FastDiv f;
    if (d == 0) d = 1;
    f.d = d;
    uint32_t s = 0;
    while ((1ull << s) < d) ++s;
    f.shr = s;
    f.mul = (uint32_t)(((1ull << 32) * ((1ull << s) - d)) / d + 1);
    return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv &f)
{
    return (__umulhi(n, f.mul) + n) >> f.shr;
}

// ---- geometry and tiling of one launch ---------------------------------------------------
struct Sweep {
    int64_t st0, st1;          // element strides of axes 0 and 1
    int32_t n0, n1, n2, n3;    // extents
    int32_t n3p;               // row pitch in elements (>= n3; > n3 for rows padded to the vector width)
    int32_t n3v, mv;           // vectors per (padded) row / per inner plane (n2*n3v)
    int32_t i0, j0, ni;        // sweep box: axis-0 origin/extent, axis-1 origin
    int32_t tj, tl, nfull;     // strip width, width of the trailing partial strip (0: none), #full strips
    int32_t cps_full, cps_last;// chunks per (i, strip) slab
    int32_t tiles_full;        // ni * cps_full
    int32_t ntiles;
    int32_t oi0, oi1, oj0, oj1;// voxels with i in [oi0,oi1) and j in [oj0,oj1) enter the reductions
    int32_t dynamic;           // 1: tiles are drawn from a global counter, 0: static stride gridDim.x
    int32_t own_store;         // 1: the fused kernels store recon_out only for voxels inside the owned range
    FastDiv d_tiles_full, d_cps_full, d_cps_last, d_mv, d_n3v;
};

struct Coord {                 // what one thread works on
    int64_t e;                 // element offset of the first of its VW voxels
    int32_t i, j, k, l0;
    int32_t vl;                // index of the last real voxel in this vector (VW-1 unless the row ends inside it)
    bool row_end;              // this vector holds the last real voxel of its row
    bool active, owned;
};

template <int VW>
__device__ __forceinline__ Coord locate(const Sweep &S, int32_t t)
{
    int32_t s, ii, c, width;
    const int32_t nfull_tiles = S.nfull * S.tiles_full;
    if (t < nfull_tiles) {
        s = (int32_t)fdiv((uint32_t)t, S.d_tiles_full);
        const int32_t r = t - s * S.tiles_full;
        ii = (int32_t)fdiv((uint32_t)r, S.d_cps_full);
        c = r - ii * S.cps_full;
        width = S.tj;
    } else {
        const int32_t r = t - nfull_tiles;
        s = S.nfull;
        ii = (int32_t)fdiv((uint32_t)r, S.d_cps_last);
        c = r - ii * S.cps_last;
        width = S.tl;
    }
    Coord p;
    const int32_t jb = S.j0 + s * S.tj;
    p.i = S.i0 + ii;
    const int32_t len = width * S.mv;                 // vectors in this slab
    const int32_t q = c * kBlock + (int32_t)threadIdx.x;
    p.active = q < len;
    const int32_t qq = p.active ? q : 0;
    const int32_t jj = (int32_t)fdiv((uint32_t)qq, S.d_mv);
    const int32_t m = qq - jj * S.mv;
    p.k = (int32_t)fdiv((uint32_t)m, S.d_n3v);
    p.l0 = (m - p.k * S.n3v) * VW;
    p.row_end = p.l0 + VW >= S.n3;
    p.vl = p.row_end ? S.n3 - 1 - p.l0 : VW - 1;
    p.j = jb + jj;
    p.e = (int64_t)p.i * S.st0 + (int64_t)jb * S.st1 + (int64_t)qq * VW;
    p.owned = p.active && p.i >= S.oi0 && p.i < S.oi1 && p.j >= S.oj0 && p.j < S.oj1;
    return p;
}

// ---- vectors of VW elements (16 bytes when VW > 1) -------------------------------------------
template <typename T, int VW> struct Raw;
template <> struct Raw<float, 4>  { typedef float4  type; };
template <> struct Raw<float, 2>  { typedef float2  type; };
template <> struct Raw<float, 1>  { typedef float   type; };
template <> struct Raw<double, 2> { typedef double2 type; };
template <> struct Raw<double, 1> { typedef double  type; };

template <typename T, int VW>
struct Vec {
    T v[VW];
};

template <typename T, int VW>
union VecU {
    typename Raw<T, VW>::type r;
    Vec<T, VW> v;
    __device__ VecU() {}
};

// read-only data (never written by the running kernel): non-coherent path, default L2 policy
template <typename T, int VW>
__device__ __forceinline__ Vec<T, VW> ld_ro(const T *p)
{
    VecU<T, VW> u;
    u.r = __ldg(reinterpret_cast<const typename Raw<T, VW>::type *>(p));
    return u.v;
}
// streamed data (touched once per sweep): evict-first
template <typename T, int VW>
__device__ __forceinline__ Vec<T, VW> ld_stream(const T *p)
{
    VecU<T, VW> u;
    u.r = __ldcs(reinterpret_cast<const typename Raw<T, VW>::type *>(p));
    return u.v;
}
template <typename T, int VW>
__device__ __forceinline__ void st_stream(T *p, const Vec<T, VW> &x)
{
    VecU<T, VW> u;
    u.v = x;
    __stcs(reinterpret_cast<typename Raw<T, VW>::type *>(p), u.r);
}
// L2 policy descriptor (only the experimental TMA variant passes one explicitly; the product kernels keep the
// default policy on every read: descriptor-carrying loads and evict-first hints were measured slower)
__device__ __forceinline__ uint64_t l2_policy_evict_normal()
{
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// Neighbour loads that must be ISSUED AFTER the barrier that precedes them in the source: coherent
// (not .nc) volatile loads with a memory clobber, which neither NVVM nor ptxas may hoist above a
// bar.sync (they do hoist ld.global.nc).  See fused.cuh for why the order matters.
__device__ __forceinline__ float4 ldg_nc_ordered(const float4 *p)
{
    float4 r;
    asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ float2 ldg_nc_ordered(const float2 *p)
{
    float2 r;
    asm volatile("ld.global.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ double2 ldg_nc_ordered(const double2 *p)
{
    double2 r;
    asm volatile("ld.global.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ float ldg_nc_ordered(const float *p)
{
    float r;
    asm volatile("ld.global.f32 %0, [%1];" : "=f"(r) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ double ldg_nc_ordered(const double *p)
{
    double r;
    asm volatile("ld.global.f64 %0, [%1];" : "=d"(r) : "l"(p) : "memory");
    return r;
}
// Ordering point for ONE warp: a named barrier that only this warp's 32 threads join (ids 1..8, one per warp of the
// CTA), carrying a predicate computed from x.  It completes as soon as the warp arrives -- no waiting for the other
// warps -- but it is a BAR instruction: ptxas moves no memory operation across it, and it cannot issue before x exists.
// Round 2: replaces the CTA-wide __syncthreads_or of round 1 in the float kernels (12 % of the stall samples were
// warps waiting at that barrier for the CTA's slowest warp; measured +2 % on config 3, +4 % 4-D unaccelerated, +6 %
// config 1).  The float64 kernels keep the CTA barrier: with coherent self loads they spill (measured 34 -> 25).
template <typename T> struct WarpOrder { static constexpr bool value = sizeof(T) == 4; };
#ifdef CYTVDN_CTA_BARRIER_ORDER        // build knob: round 1's ordering everywhere
template <> struct WarpOrder<float> { static constexpr bool value = false; };
#endif
__device__ __forceinline__ unsigned order_after(float x)
{
    unsigned r;
    asm volatile("{ .reg .pred p, q; setp.neu.f32 q, %1, %1; barrier.red.or.pred p, %2, 32, q; selp.u32 %0, 1, 0, p; }"
                 : "=r"(r) : "f"(x), "r"(1u + (threadIdx.x >> 5)) : "memory");
    return r;
}
__device__ __forceinline__ unsigned order_after(double x)
{
    unsigned r;
    asm volatile("{ .reg .pred p, q; setp.neu.f64 q, %1, %1; barrier.red.or.pred p, %2, 32, q; selp.u32 %0, 1, 0, p; }"
                 : "=r"(r) : "d"(x), "r"(1u + (threadIdx.x >> 5)) : "memory");
    return r;
}

template <typename T, int VW>
__device__ __forceinline__ Vec<T, VW> ld_ro_ordered(const T *p)
{
    VecU<T, VW> u;
    u.r = ldg_nc_ordered(reinterpret_cast<const typename Raw<T, VW>::type *>(p));
    return u.v;
}

// data that is read and later re-read as a neighbour / written in place: default policy
template <typename T, int VW>
__device__ __forceinline__ Vec<T, VW> ld_plain(const T *p)
{
    VecU<T, VW> u;
    u.r = *reinterpret_cast<const typename Raw<T, VW>::type *>(p);
    return u.v;
}
template <typename T, int VW>
__device__ __forceinline__ void st_plain(T *p, const Vec<T, VW> &x)
{
    VecU<T, VW> u;
    u.v = x;
    *reinterpret_cast<typename Raw<T, VW>::type *>(p) = u.r;
}

// ---- reductions: per-thread float64, warp shuffle, block smem, last-block fixed-order sum -------
struct RedWork {
    double *partials;     // [grid][NR]
    unsigned *ticket;     // [0] arrival ticket of the reduction, [1] tile counter; both zero between launches
    double *out;          // [NR] device
};

// Tile scheduler of the persistent sweeps.  Static mode: CTA c takes tiles c, c+G, c+2G, ... (no
// synchronisation; right when the kernel has the GPU to itself).  Dynamic mode (cytvdn_step_opts.flags
// bit 0): tiles are handed out in sweep order from one global counter, so the CTAs that are resident share
// all the work -- a sweep that co-runs with NCCL's send/recv CTAs during the halo exchange is not left
// waiting for CTAs that could not be placed.  The next index is requested at the start of a tile and
// published at its end (latency hidden).
struct TileSched {
    unsigned *counter;
    int32_t dynamic;
    int32_t next;         // thread 0 only
    __device__ __forceinline__ int32_t first() const { return (int32_t)blockIdx.x; }
    __device__ __forceinline__ void prefetch()
    {
        if (dynamic && threadIdx.x == 0) next = (int32_t)(gridDim.x + atomicAdd(counter, 1u));
    }
    // all threads; in dynamic mode it contains one barrier.  `body_has_barrier`: the sweep body executes a
    // block barrier between two calls (else one is added here to protect the shared slot).
    template <bool body_has_barrier>
    __device__ __forceinline__ int32_t advance(int32_t t)
    {
        if (!dynamic) return t + (int32_t)gridDim.x;
        __shared__ int32_t s_next;
        if (!body_has_barrier) __syncthreads();
        if (threadIdx.x == 0) s_next = next;
        __syncthreads();
        return s_next;
    }
};

template <int NR>
__device__ __forceinline__ void reduce_finish(double (&acc)[NR], const RedWork &W)
{
    __shared__ double sm[kWarps][NR];
    __shared__ bool last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        double x = acc[r];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        if (lane == 0) sm[warp][r] = x;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            double x = 0;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) x += sm[w][r];
            W.partials[(size_t)blockIdx.x * NR + r] = x;
        }
        __threadfence();
        const unsigned old = atomicAdd(W.ticket, 1u);
        last = (old == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    // the last CTA sums the per-CTA partials in a fixed order -> run-to-run deterministic
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        double x = 0;
        for (unsigned b = threadIdx.x; b < gridDim.x; b += kBlock)
            x += __ldcg(&W.partials[(size_t)b * NR + r]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        if (lane == 0) sm[warp][r] = x;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            double x = 0;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) x += sm[w][r];
            W.out[r] = x;
        }
        W.ticket[0] = 0;
        W.ticket[1] = 0;          // every CTA has drawn its last tile index before it got here
    }
}

// |x| summed over one vector: float data is summed in float over the VW (<= 16) values a
// thread handles per tile and promoted to double before it meets the running sum.
__device__ __forceinline__ float  absval(float x)  { return fabsf(x); }
__device__ __forceinline__ double absval(double x) { return fabs(x); }

}  // namespace cytvdn
