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
    const T *ref;         // reference_data (cyTVDN.py:186-187) or nullptr; SSE instantiations only
    RedWork W;            // out[0] = sum|b'|, out[1] = sum|recon' - recon|, out[2] = sum|recon| (, out[3] = sum (ref - recon')^2)
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
// SSE: also accumulate sum (reference - recon')^2 (sum_square_error_*, utils.pyx:14-49; MSE[i+1] of cyTVDN.py:186-187)
// in the same pass: one more streamed read, no extra sweep.
// MIRROR: BC_mode 3 on every axis (this repo's well-defined mirror, include/cytvdn_b200.h): the backward neighbour of
// index 0 is index 1 (anisotropic.pyx:69-70) and the forward index of the last voxel is clamped (utils.pyx:117-120
// with min instead of max), i.e. that axis' divergence term is w * (b' - b') there.
template <typename T, int VW, bool FISTA, bool AX2, bool PEER, bool SSE = false, bool MIRROR = false>
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
    // (per-warp ordering: coherent loads -- ptxas sinks read-only loads below the cheap per-warp barrier, next to the
    //  neighbour loads, where they would be fetched concurrently with them)
    constexpr bool WO = WarpOrder<T>::value;
    auto ld_self = [&](const T *p) -> Vec<T, VW> { return WO ? ld_ro_ordered<T, VW>(p) : ld_ro<T, VW>(p); };
    double acc[SSE ? 4 : 3] = {};
    constexpr int NFAR = AX2 ? 3 : 2;                 // far axes 0, 1 (, 2)

    TileSched sched{P.W.ticket + 1, S.dynamic, 0};
    for (int32_t t = sched.first(); t < S.ntiles; t = sched.template advance<!WO>(t)) {
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
            if (MIRROR) poff[d] = coord[d] != 0 ? e - stride[d] : e + stride[d];
            else poff[d] = coord[d] != 0 ? e - stride[d] : (P.bc[d] == 2 ? e : e + span);
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
        if (WO) { if (order_after(left) == 0x5a5a5a5au) return; }        // never taken (result is 0 or 1)
        else if (__syncthreads_or(left != left) == 0x5a5a5a5a) return;
#pragma unroll
        for (int d = 0; d < NFAR; ++d) {
            pv[d] = ld_ro_ordered<T, VW>(d == 0 ? pv0_ptr : P.uin + poff[d]);
            uy[d] = ld_ro_ordered<T, VW>(d == 0 ? uy0_ptr : P.uin + yoff[d]);
            by[d] = ld_ro_ordered<T, VW>(d == 0 ? by0_ptr : P.bin[d] + yoff[d]);
            if (FISTA) dy[d] = ld_ro_ordered<T, VW>(d == 0 ? dy0_ptr : P.din[d] + yoff[d]);
        }
        if (c.l0 == 0) {
            if (MIRROR) left = VW > 1 ? us.v[VW > 1 ? 1 : 0] : __ldg(P.uin + e + 1);
            else left = (P.bc[3] == 2) ? us.v[0] : __ldg(P.uin + e + (S.n3 - 1));
        } else if (lane == 0) left = __ldg(P.uin + e - 1);
        Vec<T, VW> v3, n3s;      // clipped value and new accumulator of this thread's voxels, axis 3
#pragma unroll
        for (int v = 0; v < VW; ++v)
            acc_update<T, FISTA>(us.v[v], v == 0 ? left : us.v[v - 1], b3.v[v], FISTA ? d3.v[v] : T(0),
                                 P.clip[3], P.tk, v3.v[v], n3s.v[v]);
        T right3 = __shfl_down_sync(0xffffffffu, n3s.v[0], 1);     // b'_3 of the next voxel on the row
        T wrap3 = T(0);                     // b'_3 beyond the row's last voxel: the row's voxel 0 (or 0)
        if (c.row_end) {
            if (!(P.zero_wrap & 8) && !MIRROR) {
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
            if (c.row_end && v == c.vl) fwd = MIRROR ? n3s.v[v] : wrap3;
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
                if (MIRROR && at_end[d]) nf = ns.v[v];
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
        if (S.own_store ? c.owned : c.active) st_stream<T, VW>(P.uout + e, un);
        if (c.owned) {
            acc[0] += (double)sb;
            acc[1] += (double)sd;
            acc[2] += (double)so;
        }
        if (SSE) {
            const Vec<T, VW> rf = ld_stream<T, VW>(P.ref + e);
            if (c.owned) {
#pragma unroll
                for (int v = 0; v < VW; ++v)
                    if (v <= c.vl) { const T t = rf.v[v] - un.v[v]; acc[SSE ? 3 : 0] += (double)(t * t); }
            }
        }
    }
    reduce_finish<SSE ? 4 : 3>(acc, P.W);
}

// ------------------------------------------------------------------------------------------------------------
// Fused iteration with half-isotropic axis pairs (halfisotropic.pyx:63-95, :146-186 fused with utils.pyx:90-104).
//
// The pair update couples two axes: (dp, dq) = ((u - u[x-e_p]) + b_p, (u - u[x-e_q]) + b_q) is shrunk jointly.  The
// divergence needs b'_p at x + e_p, and that value depends on BOTH differences at x + e_p, so next to what the
// anisotropic fused kernel loads (u, b_p, d_p at x + e_p) a pair needs the partner's data there: b_q[x + e_p] and
// the diagonal neighbour u[x + e_p - e_q] -- two more L2 hits per far axis of a pair; for the fast axis the partner
// data are the thread's own vectors and shuffles.  Pair shrinks per voxel: 3 for the pair (0,1), 2 for (2,3)
// (b'_3[x + e_3] is the own value of the next voxel), against 2 in the two-pass kernel -- affordable only with the
// float-float hypot and the shared-reciprocal divisions of kernels.cuh (the FP64 version made this kernel slower
// than two passes).  Jia-Zhao boundary on the pair's axes (the reference's iso kernels know no other); an axis that
// is not in a pair is updated exactly as in tv_fused_kernel.  4-D arrays, full vector width only.
// ------------------------------------------------------------------------------------------------------------
#ifndef FUSED_ISO_MINB
#define FUSED_ISO_MINB 2
#endif
template <typename T, bool FISTA, bool ISO_R, bool ISO_Q>
__global__ void __launch_bounds__(kBlock, FUSED_ISO_MINB)
tv_fused_iso_kernel(const FusedParams<T> P)
{
    constexpr int VW = 16 / (int)sizeof(T);
    const Sweep &S = P.S;
    const int lane = threadIdx.x & 31;
    double acc[3] = {0.0, 0.0, 0.0};
    const T rclipR = clip_rcp(P.clip[0]), rclipQ = clip_rcp(P.clip[2]);
    auto fista = [&](T v, T d) -> T { return FISTA ? (v + P.tk * (v - d)) : v; };

    constexpr bool WO = WarpOrder<T>::value;
    auto ld_self = [&](const T *p) -> Vec<T, VW> { return WO ? ld_ro_ordered<T, VW>(p) : ld_ro<T, VW>(p); };
    TileSched sched{P.W.ticket + 1, S.dynamic, 0};
    for (int32_t t = sched.first(); t < S.ntiles; t = sched.template advance<!WO>(t)) {
        sched.prefetch();
        const Coord c = locate<VW>(S, t);
        const int64_t e = c.e;
        const int32_t coord[3] = {c.i, c.j, c.k};
        const int32_t extent[3] = {S.n0, S.n1, S.n2};
        const int64_t stride[3] = {S.st0, S.st1, (int64_t)S.n3p};
        int64_t poff[3], yoff[3];
        bool at_end[3], at0[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const int64_t span = (int64_t)(extent[d] - 1) * stride[d];
            at0[d] = coord[d] == 0;
            poff[d] = !at0[d] ? e - stride[d] : (P.bc[d] == 2 ? e : e + span);
            at_end[d] = coord[d] == extent[d] - 1;
            yoff[d] = at_end[d] ? e - span : e + stride[d];
        }

        // ---------------- phase 1: own voxels ------------------------------------------------------------
        const Vec<T, VW> us = ld_self(P.uin + e);
        const Vec<T, VW> f = ld_self(P.f + e);
        const Vec<T, VW> b3 = ld_self(P.bin[3] + e);
        Vec<T, VW> d3;
        if (FISTA) d3 = ld_self(P.din[3] + e);
        Vec<T, VW> bs[3], ds[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            bs[d] = ld_self(P.bin[d] + e);
            if (FISTA) ds[d] = ld_self(P.din[d] + e);
        }
        T left = __shfl_up_sync(0xffffffffu, us.v[VW - 1], 1);
        // ---------------- phase 2: neighbours, one memory latency later (see tv_fused_kernel) -------------
        if (WO) { if (order_after(left) == 0x5a5a5a5au) return; }        // never taken
        else if (__syncthreads_or(left != left) == 0x5a5a5a5a) return;
        Vec<T, VW> pv[3], uy[3], by[3], dy[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            pv[d] = ld_ro_ordered<T, VW>(P.uin + poff[d]);
            uy[d] = ld_ro_ordered<T, VW>(P.uin + yoff[d]);
            by[d] = ld_ro_ordered<T, VW>(P.bin[d] + yoff[d]);
            if (FISTA) dy[d] = ld_ro_ordered<T, VW>(P.din[d] + yoff[d]);
        }
        // partner data of the pairs at the forward neighbours
        Vec<T, VW> ux0, bq0, ux1, bq1, b3y;      // pair (0,1): u[x+e0-e1], b_1[x+e0]; u[x+e1-e0], b_0[x+e1]; pair (2,3): b_3[x+e2]
        if (ISO_R) {
            ux0 = ld_ro_ordered<T, VW>(P.uin + (at0[1] ? yoff[0] : yoff[0] - S.st1));      // Jia-Zhao: difference 0 at index 0
            bq0 = ld_ro_ordered<T, VW>(P.bin[1] + yoff[0]);
            ux1 = ld_ro_ordered<T, VW>(P.uin + (at0[0] ? yoff[1] : yoff[1] - S.st0));
            bq1 = ld_ro_ordered<T, VW>(P.bin[0] + yoff[1]);
        }
        if (ISO_Q) b3y = ld_ro_ordered<T, VW>(P.bin[3] + yoff[2]);

        // ---------------- differences + accumulators: own voxels ---------------------------------------------
        if (c.l0 == 0) left = (P.bc[3] == 2) ? us.v[0] : __ldg(P.uin + e + (S.n3 - 1));
        else if (lane == 0) left = __ldg(P.uin + e - 1);
        Vec<T, VW> vs[4];                                   // (u - u[x - e_d]) + b_d, then shrunk / clipped (= new d)
#pragma unroll
        for (int v = 0; v < VW; ++v) {
            vs[3].v[v] = (us.v[v] - (v == 0 ? left : us.v[v - 1])) + b3.v[v];
#pragma unroll
            for (int d = 0; d < 3; ++d) vs[d].v[v] = (us.v[v] - pv[d].v[v]) + bs[d].v[v];
        }
        // ---------------- the same at the forward neighbours x + e_d of the far axes ---------------------------
        Vec<T, VW> fa[3], fp[3];                            // difference on axis d at x + e_d / its partner's there
        bool zero[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            zero[d] = at_end[d] && ((P.zero_wrap >> d) & 1);
            const bool jz0 = at_end[d] && P.bc[d] == 2;     // neighbour sits at index 0: its own difference is 0
            T left_y = T(0);                                // pair (2,3): element before uy[2].v[0] on the next row
            if (d == 2 && ISO_Q) {
                left_y = __shfl_up_sync(0xffffffffu, uy[2].v[VW - 1], 1);
                if (c.l0 == 0) left_y = uy[2].v[0];                             // Jia-Zhao on axis 3
                else if (lane == 0) left_y = __ldg(P.uin + yoff[2] - 1);
            }
#pragma unroll
            for (int v = 0; v < VW; ++v) {
                fa[d].v[v] = (uy[d].v[v] - (jz0 ? uy[d].v[v] : us.v[v])) + by[d].v[v];
                if (d == 0 && ISO_R) fp[0].v[v] = (uy[0].v[v] - ux0.v[v]) + bq0.v[v];       // axis 1 at x + e0
                if (d == 1 && ISO_R) fp[1].v[v] = (uy[1].v[v] - ux1.v[v]) + bq1.v[v];       // axis 0 at x + e1
                if (d == 2 && ISO_Q) fp[2].v[v] = (uy[2].v[v] - (v == 0 ? left_y : uy[2].v[v > 0 ? v - 1 : 0])) + b3y.v[v];
            }
        }
        // ---------------- one more site on the fast axis: the voxel after this vector, where no lane owns it ----
        // (row end: the row's voxel 0, wrap; lane 31: first voxel of the next warp's vector)
        const bool edge_wrap = c.row_end && !(P.zero_wrap & 8), edge_next = !c.row_end && lane == 31;
        T e3 = T(0), e2 = T(0), ed3 = T(0);                 // differences on axes 3 / 2 there, and d_3 there
        if (edge_wrap || edge_next) {
            const int64_t y = edge_wrap ? e - c.l0 : e + VW;
            const T uyy = __ldg(P.uin + y);
            T ulast = us.v[0];
#pragma unroll
            for (int v = 1; v < VW; ++v)
                if (v == c.vl) ulast = us.v[v];
            const T prev3 = edge_wrap ? ((P.bc[3] == 2) ? uyy : ulast) : us.v[VW - 1];
            e3 = (uyy - prev3) + __ldg(P.bin[3] + y);
            if (ISO_Q) e2 = (uyy - (at0[2] ? uyy : __ldg(P.uin + y - S.n3p))) + __ldg(P.bin[2] + y);
            if (FISTA) ed3 = __ldg(P.din[3] + y);
        }
        // ---------------- clip or joint shrink, everything this thread holds ------------------------------------
        T amax = T(0);
        if (ISO_R || ISO_Q) {
#pragma unroll
            for (int v = 0; v < VW; ++v) {
                if (ISO_R) {
                    pair_track(amax, vs[0].v[v], vs[1].v[v]);
                    pair_track(amax, fa[0].v[v], fp[0].v[v]);
                    pair_track(amax, fa[1].v[v], fp[1].v[v]);
                }
                if (ISO_Q) {
                    pair_track(amax, vs[2].v[v], vs[3].v[v]);
                    pair_track(amax, fa[2].v[v], fp[2].v[v]);
                }
            }
            if (ISO_Q) pair_track(amax, e2, e3);
        }
        const bool fastR = ISO_R && pair_fast_ok(amax, rclipR), fastQ = ISO_Q && pair_fast_ok(amax, rclipQ);
        if (ISO_R) {
            shrink_vec<T, VW>(vs[0], vs[1], P.clip[0], rclipR, fastR);
            shrink_vec<T, VW>(fa[0], fp[0], P.clip[0], rclipR, fastR);
            shrink_vec<T, VW>(fp[1], fa[1], P.clip[0], rclipR, fastR);          // (axis 0, axis 1) order
        } else {
#pragma unroll
            for (int v = 0; v < VW; ++v) {
                vs[0].v[v] = clipval(vs[0].v[v], P.clip[0]); vs[1].v[v] = clipval(vs[1].v[v], P.clip[1]);
                fa[0].v[v] = clipval(fa[0].v[v], P.clip[0]); fa[1].v[v] = clipval(fa[1].v[v], P.clip[1]);
            }
        }
        if (ISO_Q) {
            shrink_vec<T, VW>(vs[2], vs[3], P.clip[2], rclipQ, fastQ);
            shrink_vec<T, VW>(fa[2], fp[2], P.clip[2], rclipQ, fastQ);
            if (edge_wrap || edge_next) {
                if (fastQ) shrink_fast(e2, e3, P.clip[2], rclipQ);
                else shrink_exact(e2, e3, P.clip[2]);
            }
        } else {
#pragma unroll
            for (int v = 0; v < VW; ++v) {
                vs[2].v[v] = clipval(vs[2].v[v], P.clip[2]); vs[3].v[v] = clipval(vs[3].v[v], P.clip[3]);
                fa[2].v[v] = clipval(fa[2].v[v], P.clip[2]);
            }
            e3 = clipval(e3, P.clip[3]);
        }
        // ---------------- FISTA extrapolation, stores, divergence terms ------------------------------------------
        Vec<T, VW> ns[4];
        T sb = T(0);
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            const Vec<T, VW> &dd = d == 3 ? d3 : ds[d < 3 ? d : 0];
#pragma unroll
            for (int v = 0; v < VW; ++v) {
                ns[d].v[v] = fista(vs[d].v[v], FISTA ? dd.v[v] : T(0));
                if (v <= c.vl) sb += absval(ns[d].v[v]);
            }
            if (c.active) {
                st_stream<T, VW>(P.bout[d] + e, ns[d]);
                if (FISTA) st_stream<T, VW>(P.dout[d] + e, vs[d]);
            }
        }
        Vec<T, VW> term[4];
#pragma unroll
        for (int d = 0; d < 3; ++d)
#pragma unroll
            for (int v = 0; v < VW; ++v) {
                T nf = fista(fa[d].v[v], FISTA ? dy[d].v[v] : T(0));
                if (zero[d]) nf = T(0);
                term[d].v[v] = P.w[d] * (ns[d].v[v] - nf);
            }
        T right3 = __shfl_down_sync(0xffffffffu, ns[3].v[0], 1);
        const T edge = fista(e3, ed3);                      // b'_3 at the edge site
        if (edge_next) right3 = edge;
        const T wrap3 = edge_wrap ? edge : T(0);
#pragma unroll
        for (int v = 0; v < VW; ++v) {
            T fwd = v == VW - 1 ? right3 : ns[3].v[v + 1 < VW ? v + 1 : v];
            if (c.row_end && v == c.vl) fwd = wrap3;
            term[3].v[v] = P.w[3] * (ns[3].v[v] - fwd);
        }

        // ---------------- reconstruction update ------------------------------------------------------------
        Vec<T, VW> un;
        T sd = T(0), so = T(0);
#pragma unroll
        for (int v = 0; v < VW; ++v) {
            T s = term[0].v[v] + term[1].v[v];
            s = s + term[2].v[v];
            s = s + term[3].v[v];
            un.v[v] = f.v[v] - s;
            if (v <= c.vl) {
                sd += absval(un.v[v] - us.v[v]);
                so += absval(us.v[v]);
            }
        }
        if (S.own_store ? c.owned : c.active) st_stream<T, VW>(P.uout + e, un);
        if (c.owned) {
            acc[0] += (double)sb;
            acc[1] += (double)sd;
            acc[2] += (double)so;
        }
    }
    reduce_finish<3>(acc, P.W);
}

}  // namespace cytvdn
