// fused.cuh -- one whole TV iteration (half-step A for all axes + half-step B) in ONE pass.
//
// Two-pass traffic is 96 B/voxel (4-D FISTA fp32): half-step B re-reads recon and the four
// accumulators that half-step A just wrote.  Here every tile computes, besides its own b'_d[x],
// the forward neighbours b'_d[x+e_d] it needs for the divergence from the OLD state (recon, b, d of
// the neighbouring voxel, all L2 hits in strip order), so each array crosses HBM once:
//   read  f, recon, b_d, d_d        (1 + 1 + 4 + 4)
//   write recon', b'_d, d'_d        (1 + 4 + 4)            = 19 elements = 76 B/voxel.
// Because neighbours are recomputed from the old state, the new state goes to a second set of
// arrays (ping-pong); nothing is updated in place and tiles are independent of each other.
// Per-voxel arithmetic is the reference's, operation for operation, so the
// result is bit-identical to running the two separate half-steps.
#pragma once
#include "kernels.cuh"

namespace cytvdn {

template <typename T>
struct FusedParams {
    Sweep S;
    const T *f;
    const T *uin;
    T *uout;
    const T *bin[4];
    T *bout[4];
    const T *din[4];
    T *dout[4];
    T clip[4];
    T w[4];
    T tk;
    int32_t bc[4];        // 0 periodic, 2 Jia-Zhao (mirror is undefined for half-step B)
    int32_t zero_wrap;    // bit k: b'_k beyond the last index of axis k is 0 (sharded upper edge)
    // Axis-0 halo read straight from the neighbouring GPUs' memory (peer pointers over NVLink; nullptr = none):
    // the lower neighbour's LAST plane of recon stands for index -1, the upper neighbour's FIRST planes of
    // recon / b_0 / d_0 for index n0.  Same in-plane layout (N1, N2, pitch) as the local arrays.
    const T *lo_u;        // already offset to the start of that last plane
    const T *hi_u;
    const T *hi_b0;
    const T *hi_d0;
    RedWork W;            // out[0] = sum|b'|, out[1] = sum|recon' - recon|, out[2] = sum|recon|
};

// v = clip((u - p) + b);  b' = v + tk (v - d)   (anisotropic.pyx:46-54, :127-132)
template <typename T, bool FISTA>
__device__ __forceinline__ void acc_update(T u, T p, T b, T d, T clip, T tk, T &v, T &bn)
{
    v = clipval((u - p) + b, clip);
    bn = FISTA ? (v + tk * (v - d)) : v;
}

// PEER: axis-0 halo planes may live on a neighbouring GPU (pointer selects; costs ~2 % when compiled in, so the
// single-GPU / NCCL schedules use the PEER=false instantiation)
template <typename T, int VW, bool FISTA, bool AX2, bool PEER>
#ifndef FUSED_MINB
#define FUSED_MINB 2
#endif
#ifndef FUSED_MINB_PLAIN
#define FUSED_MINB_PLAIN 3      // 3-D unaccelerated variant: few arrays, 3 CTAs per SM fit without spills (measured
                                // +17..60 %); the 4-D unaccelerated variant would spill and was measured slower
#endif
#ifdef FUSED_MAXNREG           // experiment knob: cap the registers directly instead of through CTAs per SM.
                               // Measured (config 3, 13.4 ms base): 192 threads x 3 CTAs at 112 regs 14.9 ms,
                               // at 104 regs 18.4 ms, 128 threads x 5 CTAs at 96 regs 18.3 ms -- more resident
                               // threads do not help, spills hurt.
__global__ void __maxnreg__(FUSED_MAXNREG)
#else
__global__ void __launch_bounds__(kBlock, ((!FISTA && !AX2) ? FUSED_MINB_PLAIN : FUSED_MINB))
#endif
tv_fused_kernel(const FusedParams<T> P)
{
    const Sweep &S = P.S;
    const int lane = threadIdx.x & 31;
    // All reads keep the default L2 policy: marking the last-use ("self") loads evict-first was measured
    // 25 % slower -- inside a wave a neighbour may still need the line.
    auto ld_self = [&](const T *p) -> Vec<T, VW> { return ld_ro<T, VW>(p); };
    double acc[3] = {0.0, 0.0, 0.0};
    constexpr int NFAR = AX2 ? 3 : 2;                 // far axes 0, 1 (, 2)

    TileSched sched{P.W.ticket + 1, S.dynamic, 0};
    for (int32_t t = sched.first(); t < S.ntiles; t = sched.template advance<true>(t)) {
        sched.prefetch();
        // Inactive threads (tail of a slab) run on valid addresses of the slab start and only skip the
        // stores and the sums: no divergence before the loads, so all of them are issued back to back.
        const Coord c = locate<VW>(S, t);
        const int64_t e = c.e;
        const int32_t coord[3] = {c.i, c.j, c.k};
        const int32_t extent[3] = {S.n0, S.n1, S.n2};
        const int64_t stride[3] = {S.st0, S.st1, (int64_t)S.n3p};

        // ---------------- addresses (selects, no branches) -----------------------------------------
        int64_t poff[3], yoff[3];
        bool at_end[3];
#pragma unroll
        for (int d = 0; d < NFAR; ++d) {
            const int64_t span = (int64_t)(extent[d] - 1) * stride[d];
            // backward neighbour; at index 0: itself (Jia-Zhao, difference 0) or the last index (periodic)
            poff[d] = coord[d] != 0 ? e - stride[d] : (P.bc[d] == 2 ? e : e + span);
            at_end[d] = coord[d] == extent[d] - 1;
            yoff[d] = at_end[d] ? e - span : e + stride[d];          // forward neighbour (wraps to index 0)
        }

        // axis-0 neighbours that live on another GPU: same in-plane offset, other base pointer
        const int64_t inplane = e - (int64_t)c.i * S.st0;
const bool lo_peer = PEER && c.i == 0 && P.lo_u != nullptr;
        const bool hi_peer = PEER && at_end[0] && P.hi_u != nullptr;
        const T *pv0_ptr = lo_peer ? P.lo_u + inplane : P.uin + poff[0];
        const T *uy0_ptr = hi_peer ? P.hi_u + inplane : P.uin + yoff[0];
        const T *by0_ptr = hi_peer ? P.hi_b0 + inplane : P.bin[0] + yoff[0];
        const T *dy0_ptr = FISTA ? (hi_peer ? P.hi_d0 + inplane : P.din[0] + yoff[0]) : nullptr;

        // ---------------- phase 1: this thread's own voxels (first touch of every line: HBM) ----------
        const Vec<T, VW> us = ld_self(P.uin + e);
        const Vec<T, VW> f = ld_self(P.f + e);
        const Vec<T, VW> b3 = ld_self(P.bin[3] + e);
        Vec<T, VW> d3;
        if (FISTA) d3 = ld_self(P.din[3] + e);
        Vec<T, VW> pv[3], bs[3], ds[3], uy[3], by[3], dy[3];
#pragma unroll
        for (int d = 0; d < NFAR; ++d) {
            bs[d] = ld_self(P.bin[d] + e);                            // last use of these lines
            if (FISTA) ds[d] = ld_self(P.din[d] + e);
        }
        // fast axis (3): neighbours live in adjacent lanes.  The shuffle consumes `us`, i.e. the warp
        // waits here until its own lines have arrived ...
        T left = __shfl_up_sync(0xffffffffu, us.v[VW - 1], 1);
        // ---------------- phase 2: neighbours, issued one memory latency later -----------------------
        // ... and only then asks for its neighbours' lines.  Those are the phase-1 lines of other warps /
        // CTAs of the same wave, requested at the same moment as ours: by now they sit in L1/L2.  Asked
        // for concurrently they would each be fetched from HBM again (measured: L2 hit rate 3 %, DRAM
        // reads x2) -- neither L1 nor L2 merges a miss into a fill that is still in flight.
        // The barrier carries a predicate computed from the shuffle result, so it cannot be scheduled
        // before the shuffle, i.e. before this warp's own lines have arrived.
        if (__syncthreads_or(left != left) == 0x5a5a5a5a) return;       // never taken (result is 0 or 1)
#pragma unroll
        for (int d = 0; d < NFAR; ++d) {
            pv[d] = ld_ro_ordered<T, VW>(d == 0 ? pv0_ptr : P.uin + poff[d]);
            uy[d] = ld_ro_ordered<T, VW>(d == 0 ? uy0_ptr : P.uin + yoff[d]);
            by[d] = ld_ro_ordered<T, VW>(d == 0 ? by0_ptr : P.bin[d] + yoff[d]);
            if (FISTA) dy[d] = ld_ro_ordered<T, VW>(d == 0 ? dy0_ptr : P.din[d] + yoff[d]);
        }
        if (c.l0 == 0) left = (P.bc[3] == 2) ? us.v[0] : __ldg(P.uin + e + (S.n3 - 1));
        else if (lane == 0) left = __ldg(P.uin + e - 1);
        Vec<T, VW> v3, n3s;      // clipped value and new accumulator of this thread's voxels, axis 3
#pragma unroll
        for (int v = 0; v < VW; ++v)
            acc_update<T, FISTA>(us.v[v], v == 0 ? left : us.v[v - 1], b3.v[v], FISTA ? d3.v[v] : T(0),
                                 P.clip[3], P.tk, v3.v[v], n3s.v[v]);
        T right3 = __shfl_down_sync(0xffffffffu, n3s.v[0], 1);     // b'_3 of the next voxel on the row
        T wrap3 = T(0);                     // b'_3 beyond the row's last voxel: the row's voxel 0 (or 0)
        if (c.row_end) {
            if (!(P.zero_wrap & 8)) {
                const int64_t y = e - c.l0;
                const T uyy = __ldg(P.uin + y);
                T ulast = us.v[0];                           // the row's last real voxel lives in this vector
#pragma unroll
                for (int v = 1; v < VW; ++v)
                    if (v == c.vl) ulast = us.v[v];
                const T py = (P.bc[3] == 2) ? uyy : ulast;   // periodic: the voxel before index 0 is the last one
                T vy;
                acc_update<T, FISTA>(uyy, py, __ldg(P.bin[3] + y), FISTA ? __ldg(P.din[3] + y) : T(0),
                                     P.clip[3], P.tk, vy, wrap3);
            }
        } else if (lane == 31) {                                    // next vector belongs to another warp
            const int64_t y = e + VW;
            T vy;
            acc_update<T, FISTA>(__ldg(P.uin + y), us.v[VW - 1], __ldg(P.bin[3] + y),
                                 FISTA ? __ldg(P.din[3] + y) : T(0), P.clip[3], P.tk, vy, right3);
        }
        if (c.active) {
            st_stream<T, VW>(P.bout[3] + e, n3s);
            if (FISTA) st_stream<T, VW>(P.dout[3] + e, v3);
        }
        T sb = T(0);                       // sum |b'| of this thread's voxels (<= 16 values)
        Vec<T, VW> term[4];                // w_d (b'_d[x] - b'_d[x+e_d])
#pragma unroll
        for (int v = 0; v < VW; ++v) {
            if (v <= c.vl) sb += absval(n3s.v[v]);           // pad voxels of a padded row do not count
            T fwd = v == VW - 1 ? right3 : n3s.v[v + 1 < VW ? v + 1 : v];
            if (c.row_end && v == c.vl) fwd = wrap3;
            term[3].v[v] = P.w[3] * (n3s.v[v] - fwd);
        }

        // ---------------- far axes: b'_d at x (stored) and at x + e_d (recomputed, used only here) ----
#pragma unroll
        for (int d = 0; d < NFAR; ++d) {
            Vec<T, VW> vs, ns;
            const bool peer_fwd = d == 0 && hi_peer;          // forward neighbour is a real plane of the next GPU
            const bool zero = at_end[d] && ((P.zero_wrap >> d) & 1) && !peer_fwd;
            const bool jz0 = at_end[d] && P.bc[d] == 2 && !peer_fwd;   // neighbour sits at index 0: difference is 0
#pragma unroll
            for (int v = 0; v < VW; ++v) {
                acc_update<T, FISTA>(us.v[v], pv[d].v[v], bs[d].v[v], FISTA ? ds[d].v[v] : T(0), P.clip[d], P.tk,
                                     vs.v[v], ns.v[v]);
                T vy, nf;
                acc_update<T, FISTA>(uy[d].v[v], jz0 ? uy[d].v[v] : us.v[v], by[d].v[v], FISTA ? dy[d].v[v] : T(0),
                                     P.clip[d], P.tk, vy, nf);
                if (zero) nf = T(0);
                if (v <= c.vl) sb += absval(ns.v[v]);
                term[d].v[v] = P.w[d] * (ns.v[v] - nf);
            }
            if (c.active) {
                st_stream<T, VW>(P.bout[d] + e, ns);
                if (FISTA) st_stream<T, VW>(P.dout[d] + e, vs);
            }
        }

        // ---------------- reconstruction update (utils.pyx:96-104) --------------------------------
        Vec<T, VW> un;
        T sd = T(0), so = T(0);
#pragma unroll
        for (int v = 0; v < VW; ++v) {
            T s = term[0].v[v] + term[1].v[v];
            if (AX2) s = s + term[2].v[v];
            s = s + term[3].v[v];
            un.v[v] = f.v[v] - s;
            if (v <= c.vl) {
                sd += absval(un.v[v] - us.v[v]);
                so += absval(us.v[v]);
            }
        }
        if (c.active) st_stream<T, VW>(P.uout + e, un);
        if (c.owned) {
            acc[0] += (double)sb;
            acc[1] += (double)sd;
            acc[2] += (double)so;
        }
    }
    reduce_finish<3>(acc, P.W);
}

}  // namespace cytvdn
