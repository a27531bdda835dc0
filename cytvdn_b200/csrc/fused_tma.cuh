// fused_tma.cuh -- EXPERIMENTAL variant of the fused iteration kernel (off by default, CYTVDN_FUSED_TMA=1):
// the tile's own ("self") lines are staged through shared memory by the TMA unit one tile ahead
// (cp.async.bulk + mbarrier, two stages), the neighbour loads stay ordinary loads.
//
// Idea: fused.cuh exposes two memory round trips per tile (self loads, then -- one latency later -- the
// neighbour loads); prefetching the 10 contiguous 4 KiB self chunks of the NEXT tile with the TMA unit takes
// the HBM latency off the critical path and costs no registers.
//
// Measured on B200 (config 3): 17.2 ms per iteration against 13.2 ms for fused.cuh, bit-identical results.
// ncu: DRAM reads 73.3 GB instead of 44.7 GB, L2 hit rate 10 % instead of 22 % -- 17 array-reads = the 10
// TMA-staged arrays plus every first-touch neighbour load: lines brought in by ordinary loads are not L2 hits
// for the bulk copies that follow (and vice versa), whatever the L2 policy hint on the bulk copy (evict_normal,
// evict_last) and whether the neighbour loads are issued before or after the stage has arrived.  With the
// neighbour reuse gone the variant moves 112 GB per launch.  Kept as a documented negative result; the product
// path is fused.cuh.
#pragma once
#include "fused.cuh"

namespace cytvdn {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
// 1-D bulk copy global -> shared, completion counted in bytes on `bar`, with an explicit L2 policy
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t pol)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}

// first element offset and number of vectors of tile t (uniform; what thread 0 needs to prefetch a tile)
template <int VW>
__device__ __forceinline__ void tile_extent(const Sweep &S, int32_t t, int64_t &e0, int32_t &nvec)
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
    const int32_t len = width * S.mv;
    const int32_t q0 = c * kBlock;
    nvec = len - q0 < kBlock ? len - q0 : kBlock;
    e0 = (int64_t)(S.i0 + ii) * S.st0 + (int64_t)(S.j0 + s * S.tj) * S.st1 + (int64_t)q0 * VW;
}

template <typename T, int VW, bool FISTA, bool AX2>
__global__ void __launch_bounds__(kBlock, 2)
tv_fused_tma_kernel(const FusedParams<T> P)
{
    constexpr int NFAR = AX2 ? 3 : 2;
    constexpr int NSELF = 2 + (NFAR + 1) * (FISTA ? 2 : 1);          // recon, orig, b_d (, d_d)
    constexpr int kTileBytes = kBlock * VW * (int)sizeof(T);          // 4 KiB
    constexpr int kStageBytes = NSELF * kTileBytes;
    extern __shared__ __align__(128) unsigned char smem[];            // 2 stages
    __shared__ uint64_t full[2];

    const Sweep &S = P.S;
    const int lane = threadIdx.x & 31;
    double acc[3] = {0.0, 0.0, 0.0};

    // self arrays in stage order: 0 recon, 1 orig, 2.. b_d (d = 3, 0, 1, 2), then d_d in the same order
    const uint64_t pol = l2_policy_evict_normal();
    auto issue = [&](int32_t t, int stage) {        // thread 0 only
        int64_t e0;
        int32_t nvec;
        tile_extent<VW>(S, t, e0, nvec);
        const uint32_t bytes = (uint32_t)nvec * VW * (uint32_t)sizeof(T);
        unsigned char *dst = smem + stage * kStageBytes;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic reads of this stage are done
        mbar_expect_tx(&full[stage], bytes * NSELF);
        tma_load_1d(dst + 0 * kTileBytes, P.uin + e0, bytes, &full[stage], pol);
        tma_load_1d(dst + 1 * kTileBytes, P.f + e0, bytes, &full[stage], pol);
        const int order[4] = {3, 0, 1, 2};
#pragma unroll
        for (int k = 0; k < NFAR + 1; ++k) {
            const int d = order[k];
            tma_load_1d(dst + (2 + k) * kTileBytes, P.bin[d] + e0, bytes, &full[stage], pol);
            if (FISTA) tma_load_1d(dst + (2 + NFAR + 1 + k) * kTileBytes, P.din[d] + e0, bytes, &full[stage], pol);
        }
    };
    auto lds = [&](int stage, int slot) -> Vec<T, VW> {
        VecU<T, VW> u;
        u.r = *reinterpret_cast<const typename Raw<T, VW>::type *>(smem + stage * kStageBytes + slot * kTileBytes +
                                                                  threadIdx.x * (VW * sizeof(T)));
        return u.v;
    };

    if (threadIdx.x == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    int32_t t = blockIdx.x;
    if (threadIdx.x == 0 && t < S.ntiles) issue(t, 0);

    for (int it = 0; t < S.ntiles; t += gridDim.x, ++it) {
        const int stage = it & 1;
        const int32_t tn = t + (int32_t)gridDim.x;
        if (threadIdx.x == 0 && tn < S.ntiles) issue(tn, stage ^ 1);  // next tile: in flight while we work on this one

        const Coord c = locate<VW>(S, t);
        const int64_t e = c.e;
        const int32_t coord[3] = {c.i, c.j, c.k};
        const int32_t extent[3] = {S.n0, S.n1, S.n2};
        const int64_t stride[3] = {S.st0, S.st1, (int64_t)S.n3p};
        int64_t poff[3], yoff[3];
        bool at_end[3];
#pragma unroll
        for (int d = 0; d < NFAR; ++d) {
            const int64_t span = (int64_t)(extent[d] - 1) * stride[d];
            poff[d] = coord[d] != 0 ? e - stride[d] : (P.bc[d] == 2 ? e : e + span);
            at_end[d] = coord[d] == extent[d] - 1;
            yoff[d] = at_end[d] ? e - span : e + stride[d];
        }
        // own lines: staged by the TMA unit one tile ago.  The neighbour loads are issued only AFTER they have
        // arrived: the neighbours' lines were requested by their owners at the same moment as ours, so "ours are
        // here" is the signal that theirs are in L2 too (requested earlier they would meet fills still in flight
        // and be fetched from HBM a second time -- measured: DRAM reads 73 GB instead of 45 GB).
        mbar_wait(&full[stage], (uint32_t)(it >> 1) & 1u);      // k-th use of a stage completes phase k
        Vec<T, VW> pv[3], uy[3], by[3], dy[3];
#pragma unroll
        for (int d = 0; d < NFAR; ++d) {
            pv[d] = ld_ro_ordered<T, VW>(P.uin + poff[d]);
            uy[d] = ld_ro_ordered<T, VW>(P.uin + yoff[d]);
            by[d] = ld_ro_ordered<T, VW>(P.bin[d] + yoff[d]);
            if (FISTA) dy[d] = ld_ro_ordered<T, VW>(P.din[d] + yoff[d]);
        }
        Vec<T, VW> us = lds(stage, 0);
        Vec<T, VW> f = lds(stage, 1);
        Vec<T, VW> b3 = lds(stage, 2);
        Vec<T, VW> d3;
        if (FISTA) d3 = lds(stage, 2 + NFAR + 1);
        if (!c.active) {            // tail tile: the TMA copied fewer vectors, keep the arithmetic finite
#pragma unroll
            for (int v = 0; v < VW; ++v) { us.v[v] = T(0); f.v[v] = T(0); b3.v[v] = T(0); d3.v[v] = T(0); }
        }

        T left = __shfl_up_sync(0xffffffffu, us.v[VW - 1], 1);
        if (c.l0 == 0) left = (P.bc[3] == 2) ? us.v[0] : __ldg(P.uin + e + (S.n3 - 1));
        else if (lane == 0) left = __ldg(P.uin + e - 1);
        Vec<T, VW> v3, n3s;
#pragma unroll
        for (int v = 0; v < VW; ++v)
            acc_update<T, FISTA>(us.v[v], v == 0 ? left : us.v[v - 1], b3.v[v], FISTA ? d3.v[v] : T(0),
                                 P.clip[3], P.tk, v3.v[v], n3s.v[v]);
        T right3 = __shfl_down_sync(0xffffffffu, n3s.v[0], 1);
        if (c.l0 + VW == S.n3) {
            if (P.zero_wrap & 8) right3 = T(0);
            else {
                const int64_t y = e + VW - S.n3;
                const T uyy = __ldg(P.uin + y);
                const T py = (P.bc[3] == 2) ? uyy : us.v[VW - 1];
                T vy;
                acc_update<T, FISTA>(uyy, py, __ldg(P.bin[3] + y), FISTA ? __ldg(P.din[3] + y) : T(0),
                                     P.clip[3], P.tk, vy, right3);
            }
        } else if (lane == 31) {
            const int64_t y = e + VW;
            T vy;
            acc_update<T, FISTA>(__ldg(P.uin + y), us.v[VW - 1], __ldg(P.bin[3] + y),
                                 FISTA ? __ldg(P.din[3] + y) : T(0), P.clip[3], P.tk, vy, right3);
        }
        if (c.active) {
            st_stream<T, VW>(P.bout[3] + e, n3s);
            if (FISTA) st_stream<T, VW>(P.dout[3] + e, v3);
        }
        T sb = T(0);
        Vec<T, VW> term[4];
#pragma unroll
        for (int v = 0; v < VW; ++v) {
            sb += absval(n3s.v[v]);
            term[3].v[v] = P.w[3] * (n3s.v[v] - (v == VW - 1 ? right3 : n3s.v[v + 1 < VW ? v + 1 : v]));
        }
#pragma unroll
        for (int d = 0; d < NFAR; ++d) {
            Vec<T, VW> bs = lds(stage, 3 + d);
            Vec<T, VW> ds;
            if (FISTA) ds = lds(stage, 3 + NFAR + 1 + d);
            Vec<T, VW> vs, ns;
            const bool zero = at_end[d] && ((P.zero_wrap >> d) & 1);
            const bool jz0 = at_end[d] && P.bc[d] == 2;
#pragma unroll
            for (int v = 0; v < VW; ++v) {
                acc_update<T, FISTA>(us.v[v], pv[d].v[v], bs.v[v], FISTA ? ds.v[v] : T(0), P.clip[d], P.tk,
                                     vs.v[v], ns.v[v]);
                T vy, nf;
                acc_update<T, FISTA>(uy[d].v[v], jz0 ? uy[d].v[v] : us.v[v], by[d].v[v], FISTA ? dy[d].v[v] : T(0),
                                     P.clip[d], P.tk, vy, nf);
                if (zero) nf = T(0);
                sb += absval(ns.v[v]);
                term[d].v[v] = P.w[d] * (ns.v[v] - nf);
            }
            if (c.active) {
                st_stream<T, VW>(P.bout[d] + e, ns);
                if (FISTA) st_stream<T, VW>(P.dout[d] + e, vs);
            }
        }
        Vec<T, VW> un;
        T sd = T(0), so = T(0);
#pragma unroll
        for (int v = 0; v < VW; ++v) {
            T s = term[0].v[v] + term[1].v[v];
            if (AX2) s = s + term[2].v[v];
            s = s + term[3].v[v];
            un.v[v] = f.v[v] - s;
            sd += absval(un.v[v] - us.v[v]);
            so += absval(us.v[v]);
        }
        if (c.active) st_stream<T, VW>(P.uout + e, un);
        if (c.owned) {
            acc[0] += (double)sb;
            acc[1] += (double)sd;
            acc[2] += (double)so;
        }
        __syncthreads();            // every thread is done with this stage before it is refilled
    }
    reduce_finish<3>(acc, P.W);
}

}  // namespace cytvdn
