"""Scan-axis sharding of the 4-D TV iteration over several GPUs.

Re-expression of the reference's MPI scheme (`cyTVDN/mpi.py:130-210` partition, `:314-438` iteration + exchange;
description `README.md:104-120`):

* tiles over the scan axes 0/1 on a ``wx x wy`` grid, each tile stores its owned block plus ONE overlap plane towards
  every existing neighbour (`mpi.py:165-196`);
* every rank runs the same kernels on its local block;
* the planes that travel are the sender's first / last OWNED planes (corrected indices).

Two implementations share ``ShardPlan``:

**The C-ABI shard engine** (``EngineShard``, ``denoise4D_engine``; `csrc/cytvdn_shard.cu`, round 2) -- the default for
1-D splits of scan axis 0.  The loop lives in the library: fused single-pass sweeps, halo planes first, the COPY ENGINES
push the new first / last owned recon plane into the neighbours' overlap planes through peer pointers (CUDA IPC between
processes) under the interior sweep, a counter in the neighbour's header announces each plane.  torch.distributed only
carries the 128-byte handles and one all-reduce of three doubles per iteration.  Host arrays in and out are pipelined
box by box along scan axis 1 (``cytvdn_shard_run_host``).  From one process: ``tv.denoise4D(..., devices=[0, 1, ...])``.

**The torch.distributed / NCCL schedules of round 1** (``CudaShard``, ``denoise4D_sharded``, ``denoise4D_peer``) -- kept
for the reference's 2-D ``(wx, wy)`` grids (``grid="mpi"``, strided halo planes), for the reference-structured two-pass
iteration (two exchanges per iteration: accumulators right, reconstruction left) and as the cross-check of the engine
(`bench_sharded.py` compares all of them on every multi-GPU run).

Deviations from `mpi.py`, all required for the sharded result to equal the single-process one
(verified there against the compiled reference):
  - `mpi.py:325,344,408,414` send the overlap planes (``acc[-1]``, ``recon[0]``); the owned planes
    ``acc[-2]`` / ``recon[1]`` are sent here;
  - a tile on the global upper edge that has a left neighbour must not wrap its last plane onto the
    received plane 0 (`utils.pyx:98-99`): the kernel gets ``zero_wrap_mask``;
  - FISTA is supported (the reference stops at "haven't done FISTA yet", `mpi.py:310-311`): only the
    accumulators travel, never the auxiliaries;
  - ``b_norm`` / ``delta`` are produced: sums over OWNED voxels, all-reduced (3 doubles / iteration).
`mpi.py:84` knows ``BC_mode=2`` only; here ``ShardPlan(..., periodic=True)`` adds ``BC_mode=0``: the first and the last
tile of a split axis exchange planes too and the kernels treat that axis as Jia-Zhao inside the block
(``cytvdn_step_opts.flags`` bits 8..11), so the wrap is done by the exchange; the engine also knows the clamped mirror
(``periodic=2``, ``BC_mode=3``).
"""
from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np


# ------------------------------------------------------------------------------------------------
# partition (pure host logic, exercised on CPU by tests/test_sharded_cpu.py)
# ------------------------------------------------------------------------------------------------
def mpi_grid(size: Sequence[int], nworkers: int) -> Tuple[int, int]:
    """The reference's choice of (wx, wy): minimise the total edge length, first minimum wins
    (`mpi.py:131-150`)."""
    best, best_edges = None, None
    for wx in range(1, nworkers + 1):
        if nworkers % wx:
            continue
        wy = nworkers // wx
        edges = (nworkers - 1) * (2 * math.ceil(size[0] / wx) + 2 * math.ceil(size[1] / wy))
        if best_edges is None or edges < best_edges:
            best, best_edges = (wx, wy), edges
    return best


@dataclass(frozen=True)
class HaloOp:
    """One plane transfer: local plane ``send_index`` of ``array`` along ``axis`` goes to rank ``peer``
    (``kind='send'``) or plane ``recv_index`` is filled from rank ``peer`` (``kind='recv'``)."""
    kind: str
    array: str          # "b0", "b1" or "recon"
    axis: int
    index: int
    peer: int
    side: str = "lo"    # on which side of this tile the peer sits; a send to "lo" pairs with a receive from "hi"


class ShardPlan:
    """Where one rank's block sits in the global array and what it exchanges."""

    def __init__(self, gshape: Sequence[int], world: int, rank: int, grid=None, periodic: bool = False):
        """``periodic``: BC_mode 0 -- the first and the last tile of a split axis are neighbours too (the wrap
        of that axis is done by the halo exchange; not available in `mpi.py:84`)."""
        gshape = tuple(int(v) for v in gshape)
        assert len(gshape) == 4, "sharding exists for 4-D datacubes only (mpi.py:252-255)"
        if grid is None:
            grid = (world, 1)
        elif grid == "mpi":
            grid = mpi_grid(gshape[:2], world)
        wx, wy = int(grid[0]), int(grid[1])
        assert wx * wy == world, f"grid {grid} does not match world size {world}"
        self.gshape, self.world, self.rank, self.grid = gshape, world, rank, (wx, wy)
        self.periodic = bool(periodic)
        self.tile = (rank // wy, rank % wy)                       # np.unravel_index(rank, (wx, wy)), mpi.py:156
        n = [math.ceil(gshape[0] / wx), math.ceil(gshape[1] / wy)]   # mpi.py:161-162
        self.valid, self.read, self.has_lo, self.has_hi = [], [], [], []
        for k in range(2):
            t, w = self.tile[k], (wx, wy)[k]
            lo, hi = t * n[k], min((t + 1) * n[k], gshape[k])     # mpi.py:165-170
            if hi - lo < 1:
                raise ValueError(f"axis {k}: extent {gshape[k]} cannot be split over {w} tiles of {n[k]} planes "
                                 f"(tile {t} would be empty)")
            has_lo, has_hi = t > 0, t < w - 1                      # mpi.py:183-187
            if self.periodic and w > 1:
                has_lo = has_hi = True
            self.valid.append((lo, hi))
            self.read.append((lo - 1 if has_lo else lo, hi + 1 if has_hi else hi))   # mpi.py:173-180
            self.has_lo.append(has_lo)
            self.has_hi.append(has_hi)
        self.local_shape = (self.read[0][1] - self.read[0][0], self.read[1][1] - self.read[1][0]) + gshape[2:]
        # owned block in local coordinates (mpi.py:195-196)
        self.own_lo = [1 if self.has_lo[k] else 0 for k in range(2)]
        self.own_hi = [self.local_shape[k] - (1 if self.has_hi[k] else 0) for k in range(2)]

    # neighbour ranks (mpi.py:199-210); they wrap around in a periodic plan
    def peer(self, axis: int, step: int) -> int:
        t = list(self.tile)
        t[axis] = (t[axis] + step) % self.grid[axis]
        return t[0] * self.grid[1] + t[1]

    def read_indices(self, axis: int):
        """Global indices of the stored planes on a scan axis (wrapping in a periodic plan)."""
        lo, hi = self.read[axis]
        return [g % self.gshape[axis] for g in range(lo, hi)]

    def extract(self, garray):
        """This rank's block (owned + overlap planes) of a global NumPy array / torch tensor."""
        if not self.periodic:
            return garray[self.read_global]
        i0, i1 = self.read_indices(0), self.read_indices(1)
        if isinstance(garray, np.ndarray):
            return garray[np.ix_(i0, i1)]
        import torch
        a = garray.index_select(0, torch.as_tensor(i0, device=garray.device))
        return a.index_select(1, torch.as_tensor(i1, device=garray.device))

    @property
    def jz_flags(self) -> int:
        """cytvdn_step_opts.flags bits 8+k: split axes of a periodic run use the Jia-Zhao boundary inside
        the block (their wrap is the exchange's job)."""
        if not self.periodic:
            return 0
        return sum(1 << (8 + k) for k in range(2) if self.grid[k] > 1)

    @property
    def owned_local(self):
        return (slice(self.own_lo[0], self.own_hi[0]), slice(self.own_lo[1], self.own_hi[1]))

    @property
    def owned_global(self):
        return (slice(*self.valid[0]), slice(*self.valid[1]))

    @property
    def read_global(self):
        return (slice(*self.read[0]), slice(*self.read[1]))

    @property
    def owned_voxels(self) -> int:
        return ((self.valid[0][1] - self.valid[0][0]) * (self.valid[1][1] - self.valid[1][0])
                * self.gshape[2] * self.gshape[3])

    @property
    def zero_wrap_mask(self) -> int:
        """Axes on which this tile ends at the global upper edge but holds a received plane 0."""
        if self.periodic:
            return 0
        return sum(1 << k for k in range(2) if self.has_lo[k] and not self.has_hi[k])

    def after_a(self) -> List[HaloOp]:
        """Accumulators shift right (corrected plane indices)."""
        ops = []
        for k in range(2):
            if self.has_hi[k]:
                ops.append(HaloOp("send", f"b{k}", k, self.local_shape[k] - 2, self.peer(k, +1), "hi"))
            if self.has_lo[k]:
                ops.append(HaloOp("recv", f"b{k}", k, 0, self.peer(k, -1), "lo"))
        return ops

    def after_b(self) -> List[HaloOp]:
        """Reconstruction shifts left."""
        ops = []
        for k in range(2):
            if self.has_lo[k]:
                ops.append(HaloOp("send", "recon", k, 1, self.peer(k, -1), "lo"))
            if self.has_hi[k]:
                ops.append(HaloOp("recv", "recon", k, self.local_shape[k] - 1, self.peer(k, +1), "hi"))
        return ops

    # sweep boxes (axis-0 ranges) for the overlapped 1-D schedule
    def a_boxes(self):
        """(halo-first, rest) axis-0 ranges of half-step A.  Halo-first: the plane that is sent and
        the plane that will be overwritten by the receive."""
        n = self.local_shape[0]
        first = []
        if self.has_lo[0]:
            first.append((0, 1))
        if self.has_hi[0]:
            first.append((n - 2, n - 1))
        rest = _complement(n, first)
        return first, rest

    def b_boxes(self):
        """Half-step B: the first owned plane (sent left) goes first; the last plane is skipped when
        it is an overlap plane (it is received)."""
        n = self.local_shape[0]
        first = [(1, 2)] if self.has_lo[0] else []
        skip = [(n - 1, n)] if self.has_hi[0] else []
        rest = _complement(n, first + skip)
        return first, rest

    # ---- fused schedule (one pass per iteration): only the reconstruction travels, both ways ----
    def after_fused(self) -> List[HaloOp]:
        """After a fused iteration the new reconstruction's first owned plane goes left and its last
        owned plane goes right; the overlap planes are received.  (Accumulators never travel: the
        forward neighbour b'[last owned + 1] is recomputed locally from the overlap plane's state.)"""
        ops = []
        # Order matters when both neighbours are the same rank (2 tiles on a periodic axis): NCCL pairs the
        # sends and receives of two ranks in posting order, so sends go lo, hi and receives hi, lo.
        for k in range(2):
            n = self.local_shape[k]
            if self.has_lo[k]:
                ops.append(HaloOp("send", "recon", k, 1, self.peer(k, -1), "lo"))
            if self.has_hi[k]:
                ops.append(HaloOp("send", "recon", k, n - 2, self.peer(k, +1), "hi"))
            if self.has_hi[k]:
                ops.append(HaloOp("recv", "recon", k, n - 1, self.peer(k, +1), "hi"))
            if self.has_lo[k]:
                ops.append(HaloOp("recv", "recon", k, 0, self.peer(k, -1), "lo"))
        return ops

    def fused_boxes(self):
        """(halo-first, rest) axis-0 ranges of a fused iteration.  First: the planes that are sent and
        the upper overlap plane (its garbage recon must be written before the receive lands).  The lower
        overlap plane is never swept: nothing owned depends on its accumulators, its recon is received."""
        n = self.local_shape[0]
        first = set()
        if self.has_lo[0]:
            first.add(1)
        if self.has_hi[0]:
            first.update((n - 2, n - 1))
        skip = {0} if self.has_lo[0] else set()
        first -= skip
        fb = _runs(sorted(first))
        rest = _complement(n, fb + _runs(sorted(skip)))
        return fb, rest

    def describe(self) -> str:
        return (f"rank {self.rank}/{self.world} tile {self.tile} of grid {self.grid}: owns "
                f"[{self.valid[0][0]}:{self.valid[0][1]}, {self.valid[1][0]}:{self.valid[1][1]}], "
                f"stores {self.local_shape}")


def _complement(n, boxes):
    out, cur = [], 0
    for lo, hi in sorted(boxes):
        if lo > cur:
            out.append((cur, lo))
        cur = max(cur, hi)
    if cur < n:
        out.append((cur, n))
    return out


def _runs(idx):
    out = []
    for i in idx:
        if out and out[-1][1] == i:
            out[-1] = (out[-1][0], i + 1)
        else:
            out.append((i, i + 1))
    return out


def plane(t, axis: int, index: int):
    """View of one plane of a 4-D tensor / array."""
    return t[index] if axis == 0 else t[:, index]


def fista_ratio(tk: float):
    """cyTVDN.py:154-156 (float64 on the host)."""
    tk_new = (1.0 + math.sqrt(1.0 + 4.0 * tk * tk)) / 2.0
    return (tk - 1.0) / tk_new, tk_new


def halo_exchange(ops: List[HaloOp], arrays: dict, group=None):
    """Post the plane transfers of one half-step with torch.distributed P2P (NCCL on GPUs, gloo in the
    CPU tests).  Returns (works, unpack) -- call ``unpack()`` after the works completed to scatter
    planes that had to be staged (axis-1 planes are strided)."""
    import torch
    import torch.distributed as dist
    p2p, staged = [], []
    for op in ops:
        view = plane(arrays[op.array], op.axis, op.index)
        if op.kind == "send":
            buf = view if view.is_contiguous() else view.contiguous()
            p2p.append(dist.P2POp(dist.isend, buf, op.peer, group))
        else:
            if view.is_contiguous():
                p2p.append(dist.P2POp(dist.irecv, view, op.peer, group))
            else:
                buf = torch.empty_like(view, memory_format=torch.contiguous_format)
                staged.append((view, buf))
                p2p.append(dist.P2POp(dist.irecv, buf, op.peer, group))
    works = dist.batch_isend_irecv(p2p) if p2p else []

    def unpack():
        for view, buf in staged:
            view.copy_(buf)
    return works, unpack


# ------------------------------------------------------------------------------------------------
# CUDA worker: one rank's state and launches
# ------------------------------------------------------------------------------------------------
class CudaShard:
    """Device state of one rank (local block with overlap planes) and its kernel launches."""

    SLOTS = 16          # doubles of reduction scratch per iteration

    def __init__(self, plan: ShardPlan, shard, mu, lam=None, fista=True, n_iter=1, fused=False):
        self.bc_mode = 0 if plan.periodic else 2
        import torch
        from . import _lib
        self.torch, self._lib, self.lib = torch, _lib, _lib.load()
        _lib.require_gpu()
        assert tuple(shard.shape) == tuple(plan.local_shape), (tuple(shard.shape), plan.local_shape)
        assert shard.is_cuda and shard.is_contiguous() and shard.dtype in (torch.float32, torch.float64)
        dt = np.float32 if shard.dtype == torch.float32 else np.float64
        # rows that are not a multiple of the 16-byte vector width are padded in the device state
        # (cytvdn_step_opts.row_pitch), as cytvdn_denoise does, so that every shape runs on the vector path
        self.n3 = int(shard.shape[3])
        vwf = 16 // (4 if dt == np.float32 else 8)
        self.n3p = (self.n3 + vwf - 1) // vwf * vwf
        if self.n3p != self.n3:
            shard = torch.nn.functional.pad(shard, (0, self.n3p - self.n3)).contiguous()
        self.plan, self.orig, self.fista = plan, shard, fista
        mu = np.asarray(mu, dtype=dt)
        lam = (mu * 1.0 / 32.0) if lam is None else np.asarray(lam, dtype=dt)      # cyTVDN.py:67-68
        self.clip = (C.c_double * 4)(*[float(v) for v in (1.0 / lam)])
        self.w = (C.c_double * 4)(*[float(v) for v in (lam / mu).astype(dt)])
        self.code = 0 if dt == np.float32 else 1
        # all state in ONE allocation, carved into views (a B200 spends ~5 ms per multi-GB cudaMalloc)
        n_arrays = 1 + 4 * (2 if fista else 1) + ((1 + 4 * (2 if fista else 1)) if fused else 0)
        numel = shard.numel()
        pitch = (numel + 63) // 64 * 64                      # keep every view 256-byte aligned
        arena = torch.empty(n_arrays * pitch, dtype=shard.dtype, device=shard.device)
        self._arena, cursor = arena, [0]

        def carve(zero):
            v = arena[cursor[0] * pitch: cursor[0] * pitch + numel].view(shard.shape)
            cursor[0] += 1
            if zero:
                v.zero_()
            return v
        self.recon = carve(False)
        self.recon.copy_(shard)
        self.b = [carve(True) for _ in range(4)]
        self.d = [carve(True) for _ in range(4)] if fista else None
        self.bp = (C.c_void_p * 4)(*[t.data_ptr() for t in self.b])
        self.dp = (C.c_void_p * 4)(*[t.data_ptr() for t in self.d]) if fista else None
        self.sh = (C.c_int64 * 4)(*plan.local_shape)
        self.sums = torch.zeros((max(n_iter, 1), self.SLOTS), dtype=torch.float64, device=shard.device)
        self.arrays = {"b0": self.b[0], "b1": self.b[1], "recon": self.recon}
        self.launches = 0
        self.fused = fused
        if fused:       # second state set: the fused iteration is out of place (ping-pong)
            self.recon2 = carve(False)
            # zeroed: the lower overlap plane of these is never swept (fused_boxes skips plane 0) and must not hold
            # uninitialised memory that return_state / checksums would expose
            self.b2 = [carve(True) for _ in range(4)]
            self.d2 = [carve(True) for _ in range(4)] if fista else None
            self.bp2 = (C.c_void_p * 4)(*[t.data_ptr() for t in self.b2])
            self.dp2 = (C.c_void_p * 4)(*[t.data_ptr() for t in self.d2]) if fista else None
            self.first = True       # iteration 0 reads recon == orig

    def fused_step(self, it: int, slot: int, tk_ratio: float, fista: bool, box0=None, dynamic=False):
        """One fused iteration on an axis-0 range: state (recon, b, d) -> (recon2, b2, d2)."""
        st = self.torch.cuda.current_stream(self.orig.device).cuda_stream
        out = self.sums.data_ptr() + 8 * (it * self.SLOTS + slot)
        o = self._opts(box0, dynamic)
        self._lib.check(self.lib.cytvdn_fused_iteration(
            4, self.sh, self.code, self.orig.data_ptr(), self.recon.data_ptr(), self.recon2.data_ptr(),
            self.bp, self.bp2, self.dp if fista else None, self.dp2 if fista else None, float(tk_ratio),
            self.clip, self.w, self.bc_mode, out, C.byref(o), st))
        self.launches += 1

    def fused_swap(self):
        """After all boxes of an iteration (and its exchange) the new state becomes the current one."""
        self.recon, self.recon2 = self.recon2, self.recon
        self.b, self.b2, self.bp, self.bp2 = self.b2, self.b, self.bp2, self.bp
        if self.d is not None:
            self.d, self.d2, self.dp, self.dp2 = self.d2, self.d, self.dp2, self.dp
        self.arrays = {"b0": self.b[0], "b1": self.b[1], "recon": self.recon}

    def fused_local_sums(self):
        """Slots of the fused launches: 3 doubles each (sum|b|, sum|delta|, sum|old|), 4 doubles apart."""
        s = self.sums
        return self.torch.stack([s[:, 0:16:4].sum(1), s[:, 1:16:4].sum(1), s[:, 2:16:4].sum(1)], dim=1)

    def _opts(self, box0=None, dynamic=False):
        o = self._lib.StepOpts()
        o.flags = (1 if dynamic else 0) | self.plan.jz_flags
        o.row_pitch = self.n3p
        n0 = self.plan.local_shape[0]
        o.box_lo[0], o.box_hi[0] = (0, n0) if box0 is None else box0
        o.box_lo[1], o.box_hi[1] = 0, 0
        for k in range(2):
            o.own_lo[k], o.own_hi[k] = self.plan.own_lo[k], self.plan.own_hi[k]
        o.zero_wrap_mask = self.plan.zero_wrap_mask
        return o

    def half_step_a(self, it: int, slot: int, tk_ratio: float, fista: bool, box0=None, dynamic=False):
        st = self.torch.cuda.current_stream(self.orig.device).cuda_stream
        out = self.sums.data_ptr() + 8 * (it * self.SLOTS + slot)
        o = self._opts(box0, dynamic)
        self._lib.check(self.lib.cytvdn_accumulator_update_all(
            4, self.sh, self.code, self.recon.data_ptr(), self.bp, self.dp if fista else None, float(tk_ratio),
            self.clip, 0, 0, self.bc_mode, out, C.byref(o), st))
        self.launches += 1

    def half_step_b(self, it: int, slot: int, box0=None, dynamic=False):
        st = self.torch.cuda.current_stream(self.orig.device).cuda_stream
        out = self.sums.data_ptr() + 8 * (it * self.SLOTS + slot)
        o = self._opts(box0, dynamic)
        self._lib.check(self.lib.cytvdn_datacube_update(
            4, self.sh, self.code, self.orig.data_ptr(), self.recon.data_ptr(), self.recon.data_ptr(), self.bp,
            self.w, self.bc_mode, out, C.byref(o), st))
        self.launches += 1

    def result(self):
        """The current reconstruction of the local block without the row padding."""
        return self.recon if self.n3p == self.n3 else self.recon[..., :self.n3]

    # slot layout per iteration: A launches 0..3 (1 double each), B launches 4.. (2 doubles each)
    def local_sums(self):
        s = self.sums
        return self.torch.stack([s[:, 0:4].sum(1), s[:, 4:16:2].sum(1), s[:, 5:16:2].sum(1)], dim=1)


def _run_iteration_overlapped(sh: CudaShard, it, tkr, fista, group, comm_stream, mode=None):
    """Two-pass schedule, 1-D (axis-0) split: per half-step the halo planes first, then the exchange -- inline on
    the compute stream (default) or on ``comm_stream`` under the interior sweep (``mode="overlap"``, see
    ``_run_iteration_fused``) -- then the rest of the sweep."""
    torch = sh.torch
    main = torch.cuda.current_stream(sh.orig.device)
    plan = sh.plan
    if mode is None:
        mode = os.environ.get("CYTVDN_SHARD_EXCHANGE", "inline")
    overlap = mode == "overlap" and comm_stream is not None
    for phase in ("a", "b"):
        first, rest = plan.a_boxes() if phase == "a" else plan.b_boxes()
        ops = plan.after_a() if phase == "a" else plan.after_b()
        slot = 0 if phase == "a" else 4
        step = 1 if phase == "a" else 2
        for box in first:
            (sh.half_step_a(it, slot, tkr, fista, box) if phase == "a" else sh.half_step_b(it, slot, box))
            slot += step

        def exchange():
            works, unpack = halo_exchange(ops, sh.arrays, group)
            for w_ in works:
                w_.wait()
            unpack()

        if ops:
            if overlap:
                ev = torch.cuda.Event()
                ev.record(main)
                with torch.cuda.stream(comm_stream):
                    comm_stream.wait_event(ev)
                    exchange()
            else:
                exchange()
        for box in rest:          # in overlap mode it co-runs with the exchange: dynamic tile scheduling
            dyn = overlap and bool(ops)
            (sh.half_step_a(it, slot, tkr, fista, box, dyn) if phase == "a" else sh.half_step_b(it, slot, box, dyn))
            slot += step
        if ops and overlap:
            main.wait_stream(comm_stream)


def _run_iteration_fused(sh: CudaShard, it, tkr, fista, group, comm_stream, mode=None):
    """Fused schedule: planes to send first, then ONE exchange of the NEW reconstruction (both directions),
    then the state sets swap roles.  Where the exchange runs (1-D split):

    ``mode="inline"`` (default): on the compute stream, right after the halo planes -- it waits only for the
        neighbours' halo planes (computed at the same moment), costs ~0.3 ms of a ~28 ms iteration and leaves the
        interior sweep alone on the GPU.
    ``mode="overlap"``: on ``comm_stream`` under the interior sweep.  Measured SLOWER on B200 (28.4 vs 27.6 ms):
        NCCL's send/recv CTAs take SMs from the persistent sweep for longer than the transfer itself lasts.
    """
    torch = sh.torch
    main = torch.cuda.current_stream(sh.orig.device)
    plan = sh.plan
    one_d = plan.grid[1] == 1
    if mode is None:
        mode = os.environ.get("CYTVDN_SHARD_EXCHANGE", "inline")
    overlap = mode == "overlap" and one_d and comm_stream is not None
    first, rest = plan.fused_boxes() if one_d else ([], [(0, plan.local_shape[0])])
    ops = plan.after_fused()
    slot = 0
    for box in first:
        sh.fused_step(it, slot, tkr, fista, box)
        slot += 4
    new_arrays = {"recon": sh.recon2}

    def exchange():
        works, unpack = halo_exchange(ops, new_arrays, group)
        for w_ in works:
            w_.wait()
        unpack()

    if ops and one_d:
        if overlap:
            ev = torch.cuda.Event()
            ev.record(main)
            with torch.cuda.stream(comm_stream):
                comm_stream.wait_event(ev)
                exchange()
        else:
            exchange()
    for box in rest:              # in overlap mode it co-runs with the exchange: dynamic tile scheduling
        sh.fused_step(it, slot, tkr, fista, box, overlap and bool(ops))
        slot += 4
    if ops:
        if overlap:
            main.wait_stream(comm_stream)
        elif not one_d:
            exchange()            # 2-D grids: the strided halo planes are exchanged after the full sweep
    sh.fused_swap()


def _run_iteration_simple(sh: CudaShard, it, tkr, fista, group):
    """Any grid: full sweep, then exchange (what `mpi.py` does, minus its barriers)."""
    for phase in ("a", "b"):
        if phase == "a":
            sh.half_step_a(it, 0, tkr, fista)
        else:
            sh.half_step_b(it, 4)
        ops = sh.plan.after_a() if phase == "a" else sh.plan.after_b()
        if ops:
            works, unpack = halo_exchange(ops, sh.arrays, group)
            for w_ in works:
                w_.wait()
            unpack()


# ------------------------------------------------------------------------------------------------
# The C-ABI shard engine (csrc/cytvdn_shard.cu): the sharded loop lives in the library, the exchange is done by the
# copy engines through peer pointers under the interior sweep.  Python only moves the 128-byte handles between the
# ranks (torch.distributed as plumbing) and adds up three doubles per iteration at the end.
# ------------------------------------------------------------------------------------------------
class EngineShard:
    """One rank's ``cytvdn_shard`` (1-D split of scan axis 0, fused schedule)."""

    def __init__(self, gshape, world, rank, mu, lam=None, dtype=np.float32, fista=True, max_iters=1, periodic=False,
                 device=None):
        from . import _lib
        self._lib, self.lib = _lib, _lib.load()
        _lib.require_gpu()
        dt = np.dtype(dtype)
        mu = np.asarray(mu, dtype=dt)
        lam = (mu * 1.0 / 32.0) if lam is None else np.asarray(lam, dtype=dt)      # cyTVDN.py:67-68
        P = _lib.ShardParams()
        P.dtype = 0 if dt == np.float32 else 1
        # periodic: False / True (BC_mode 0 on the split axis), or 2 = the clamped mirror (BC_mode 3) at the global edges
        P.world, P.rank, P.fista = int(world), int(rank), int(bool(fista))
        P.periodic = 2 if (periodic is not True and periodic == 2) else int(bool(periodic))
        P.max_iters = max(1, int(max_iters))
        P.device = -1 if device is None else int(device)
        for k in range(4):
            P.gshape[k] = int(gshape[k])
            P.clip[k] = float((1.0 / lam)[k])
            P.lambda_mu[k] = float((lam / mu).astype(dt)[k])
        if device is None:
            d = C.c_int(0)
            _lib.check(self.lib.cytvdn_get_device(C.byref(d)))
            device = d.value
        self._dev = int(device)
        self.h = C.c_void_p()
        _lib.check(self.lib.cytvdn_shard_create(C.byref(P), C.byref(self.h)))
        self.dtype, self.world, self.rank, self.gshape = dt, int(world), int(rank), tuple(int(v) for v in gshape)
        self.refresh()

    def refresh(self):
        o = (C.c_int64 * 12)()
        self._lib.check(self.lib.cytvdn_shard_info(self.h, o))
        (self.n_local, self.own_lo, self.own_hi, self.valid_lo, self.valid_hi, self.read_lo, self.has_lo, self.has_hi,
         self.arena_bytes, self.n3p, self.launches, self.it_run) = [int(v) for v in o]
        return self

    @property
    def local_shape(self):
        return (self.n_local,) + self.gshape[1:]

    @property
    def owned_shape(self):
        return (self.own_hi - self.own_lo,) + self.gshape[1:]

    def export(self) -> bytes:
        h = (C.c_ubyte * 128)()
        self._lib.check(self.lib.cytvdn_shard_export(self.h, h))
        return bytes(h)

    def connect(self, side: int, handle: bytes):
        self._lib.check(self.lib.cytvdn_shard_connect(self.h, int(side), (C.c_ubyte * 128)(*handle)))

    def connect_all(self, handles):
        """``handles[r]`` = rank r's export; connects the lower and the upper neighbour (wrapping when periodic)."""
        self.connect(0, handles[(self.rank - 1) % self.world])
        self.connect(1, handles[(self.rank + 1) % self.world])

    def array_ptr(self, which, set_=0, axis=0) -> int:
        p = C.c_void_p()
        self._lib.check(self.lib.cytvdn_shard_array(self.h, {"orig": 0, "recon": 1, "b": 2, "d": 3}[which], set_, axis, C.byref(p)))
        return p.value

    def array(self, which, set_=0, axis=0):
        """torch view of an internal device array (rows padded to ``n3p``)."""
        import torch
        shape = (self.n_local,) + self.gshape[1:3] + (self.n3p,)
        typestr = "<f4" if self.dtype == np.float32 else "<f8"
        dev = torch.device("cuda", self.device_index())
        return torch.as_tensor(_DevArray(self.array_ptr(which, set_, axis), shape, typestr), device=dev)

    def device_index(self):
        return self._dev

    def load(self, block=None):
        """``block``: stored planes (owned + overlap) with dense rows -- CUDA tensor, NumPy array (pinned for speed) or
        None when ``orig`` was written in place."""
        if block is None:
            ptr = None
        elif hasattr(block, "data_ptr"):
            import torch
            assert block.is_contiguous() and tuple(block.shape) == self.local_shape, (tuple(block.shape), self.local_shape)
            torch.cuda.current_stream(block.device).synchronize()     # the shard copies on its own stream
            ptr = block.data_ptr()
        else:
            assert block.flags["C_CONTIGUOUS"] and block.dtype == self.dtype and tuple(block.shape) == self.local_shape
            ptr = block.ctypes.data
        self._keep = block
        self._lib.check(self.lib.cytvdn_shard_load(self.h, C.c_void_p(ptr) if ptr else None))

    def load_synth(self, seed=2, counts=500.0):
        """Generate this shard's planes of the synthetic 4D-STEM array straight into ``orig`` (dense rows only)."""
        import torch
        from . import synth
        assert self.n3p == self.gshape[3] and not (self.read_lo < 0 or self.read_lo + self.n_local > self.gshape[0])
        mod, templ = synth._stem_tables(self.gshape)
        tmod, ttem = torch.from_numpy(mod).cuda(), torch.from_numpy(templ).cuda()
        gs = (C.c_int64 * 4)(*self.gshape)
        st = torch.cuda.current_stream().cuda_stream
        self._lib.check(self.lib.cytvdn_synth_counts(gs, self.read_lo, self.n_local, 0 if self.dtype == np.float32 else 1,
                                                     tmod.data_ptr(), ttem.data_ptr(), float(counts), int(seed),
                                                     C.c_void_p(self.array_ptr("orig")), st))
        torch.cuda.current_stream().synchronize()
        self.load(None)

    def run_host(self, block, owned_out, n_fista=0, n_plain=0):
        """``load`` + ``iterate`` + ``store`` with host arrays, PCIe copies overlapped with the iterations (boxes along
        scan axis 1, wavefront over (box, iteration)); asynchronous: ``synchronize()`` / ``sums()`` wait for it."""
        assert block.flags["C_CONTIGUOUS"] and block.dtype == self.dtype and tuple(block.shape) == self.local_shape
        assert owned_out.flags["C_CONTIGUOUS"] and owned_out.dtype == self.dtype and tuple(owned_out.shape) == self.owned_shape
        self._keep = (block, owned_out)
        self._lib.check(self.lib.cytvdn_shard_run_host(self.h, C.c_void_p(block.ctypes.data), C.c_void_p(owned_out.ctypes.data),
                                                       int(n_fista), int(n_plain)))

    def iterate(self, n_fista=0, n_plain=0):
        self._lib.check(self.lib.cytvdn_shard_iterate(self.h, int(n_fista), int(n_plain)))

    def synchronize(self):
        self._lib.check(self.lib.cytvdn_shard_synchronize(self.h))

    def sums(self, n) -> np.ndarray:
        out = np.zeros((max(n, 1), 3), dtype=np.float64)
        self._lib.check(self.lib.cytvdn_shard_sums(self.h, out.ctypes.data_as(C.POINTER(C.c_double)), int(n)))
        return out[:n]

    def store(self, out):
        """Owned planes of the current reconstruction into ``out`` (CUDA tensor or NumPy array of ``owned_shape``)."""
        if hasattr(out, "data_ptr"):
            assert out.is_contiguous() and tuple(out.shape) == self.owned_shape
            ptr = out.data_ptr()
        else:
            assert out.flags["C_CONTIGUOUS"] and out.dtype == self.dtype and tuple(out.shape) == self.owned_shape
            ptr = out.ctypes.data
        self._lib.check(self.lib.cytvdn_shard_store(self.h, C.c_void_p(ptr)))
        return out

    def profile(self, on=True):
        self._lib.check(self.lib.cytvdn_shard_profile(self.h, int(on)))

    def timeline(self, n) -> np.ndarray:
        out = np.zeros((max(n, 1), 6), dtype=np.float64)
        self._lib.check(self.lib.cytvdn_shard_timeline(self.h, out.ctypes.data_as(C.POINTER(C.c_double)), int(n)))
        return out[:n]

    def close(self):
        if self.h:
            self.lib.cytvdn_shard_destroy(self.h)
            self.h = C.c_void_p()

    def close_collective(self, group=None):
        """Tear-down on every rank of a process group: an arena a neighbour still has mapped through CUDA IPC stays
        allocated until that neighbour unmaps it, so all ranks first unmap (disconnect), then free (destroy)."""
        import torch.distributed as dist
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        if self.h:
            self.lib.cytvdn_shard_synchronize(self.h)
        if multi:
            dist.barrier(group=group)          # nobody unmaps / frees an arena a neighbour still pushes into
        if self.h:
            self.lib.cytvdn_shard_disconnect(self.h)
        if multi:
            dist.barrier(group=group)          # every mapping of every arena is gone: cudaFree really frees
        self.close()


def _split_iterations(iterations, FISTA):
    if type(iterations) in (list, tuple):
        return int(iterations[0]), int(iterations[1])
    return int(iterations * bool(FISTA)), int(iterations * (not FISTA))


def denoise4D_engine(block, mu, iterations=10, FISTA=True, stopping_relative_change=None, *, gshape, group=None, lam=None,
                     periodic=False, out=None, engine=None, return_engine=False):
    """Sharded ``denoise4D`` through the C-ABI shard engine -- call it on every rank (one process per GPU).

    ``block``: this rank's stored planes (owned + one overlap plane per neighbour, ``ShardPlan.read_global``) as a
    contiguous CUDA tensor or a (pinned) NumPy array.  Returns ``(recon_owned, b_norm, delta_recon)``: the rank's OWNED
    planes (``out`` or a new array of the input's kind) and the global scalars.  The halo exchange is done inside the
    library by the copy engines over NVLink (CUDA IPC peer pointers); torch.distributed only carries the handles and
    one all-reduce of 3 doubles per iteration (per iteration only with ``stopping_relative_change``)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    nF, nU = _split_iterations(iterations, FISTA)
    n = nF + nU
    is_t = hasattr(block, "data_ptr")
    dt = np.dtype(str(block.dtype).replace("torch.", "")) if is_t else block.dtype
    dev = block.device if is_t else torch.device("cuda", torch.cuda.current_device())
    # the three sums per iteration travel over whatever backend the group has (gloo reduces host tensors)
    red_dev = dev if (world > 1 and dist.get_backend(group) == "nccl") else torch.device("cpu")
    own_engine = engine is None
    if own_engine:
        engine = EngineShard(gshape, world, rank, mu, lam, dt, fista=nF > 0, max_iters=max(n, 1), periodic=periodic,
                             device=dev.index)
        handles = [None] * world
        if world > 1:
            dist.all_gather_object(handles, engine.export(), group=group)
        else:
            handles = [engine.export()]
        engine.connect_all(handles)
    try:
        ran = np.zeros(n, dtype=bool)
        glob = np.zeros((max(n, 1), 3))
        # host arrays in and out, fixed iteration count: the library overlaps the PCIe copies with the iterations
        # (wavefront over boxes of scan axis 1; no-op fall-back inside for periodic runs / padded rows)
        piped = (not is_t) and stopping_relative_change is None and n > 0 and (out is None or not hasattr(out, "data_ptr"))
        if piped:
            if out is None:
                out = np.empty(engine.owned_shape, dtype=dt)
            engine.run_host(block, out, nF, nU)
            ran[:] = True
        else:
            engine.load(block)
            if world > 1:
                dist.barrier(group=group)      # every rank's flags and state are in place before anyone pushes
        if piped:
            pass
        elif stopping_relative_change is None:
            engine.iterate(nF, nU)
            ran[:] = True
        else:
            stop = False
            for phase, count in ((0, nF), (1, nU)):
                for j in range(count):
                    it = j if phase == 0 else nF + j
                    engine.iterate(1 if phase == 0 else 0, 0 if phase == 0 else 1)
                    ran[it] = True
                    k = engine.refresh().it_run                # sums are stored per iteration that RAN
                    s = torch.from_numpy(engine.sums(k)[k - 1].copy()).to(red_dev)
                    if world > 1:
                        dist.all_reduce(s, group=group)
                    s = s.cpu().numpy()
                    dl = float(s[1] / s[2])
                    if dt == np.float32:
                        dl = float(np.float32(dl))
                    if dl < stopping_relative_change:      # cyTVDN.py:189-194: leaves THIS phase only
                        break
        done = int(ran.sum())
        loc = np.zeros((max(n, 1), 3))
        # sums of the iterations that ran are stored consecutively (it_run counts them)
        got = engine.sums(engine.refresh().it_run)
        loc[np.nonzero(ran)[0]] = got[:done]
        g = torch.from_numpy(loc).to(red_dev)
        if world > 1:
            dist.all_reduce(g, group=group)
        g = g.cpu().numpy()
        with np.errstate(all="ignore"):
            b_norm = np.where(ran, g[:n, 0], 0.0).astype(dt)
            delta = np.where(ran, g[:n, 1] / g[:n, 2], 0.0).astype(dt)
        if not piped:
            if out is None:
                out = (torch.empty(engine.owned_shape, dtype=block.dtype, device=dev) if is_t
                       else np.empty(engine.owned_shape, dtype=dt))
            engine.store(out)
        if world > 1:
            dist.barrier(group=group)          # nobody frees (or re-loads) an arena a neighbour still pushes into
        if own_engine and not return_engine:
            engine.close_collective(group)
        if return_engine:
            return out, b_norm, delta, engine
        return out, b_norm, delta
    except BaseException:
        if own_engine:
            engine.close()                     # error path: no collective tear-down (the other ranks may be gone)
        raise


def emulate_engine_on_one_device(gdata, mu, world, iterations=10, FISTA=True, periodic=False, lam=None):
    """All ranks of the shard engine in THIS process on one device (handles resolve to plain pointers): validates the
    engine -- box order, flags, pushes, owned-only stores -- where fewer GPUs than ranks exist.  Needs
    CUDA_DEVICE_MAX_CONNECTIONS >= 2*world so that the streams do not share hardware queues (a spinning wait kernel
    must never sit in front of the work it waits for).  Returns (assembled recon, b_norm, delta)."""
    import torch
    nF, nU = _split_iterations(iterations, FISTA)
    n = nF + nU
    dt = np.float32 if gdata.dtype == torch.float32 else np.float64
    eng = [EngineShard(gdata.shape, world, r, mu, lam, dt, fista=nF > 0, max_iters=max(n, 1), periodic=periodic)
           for r in range(world)]
    try:
        handles = [e.export() for e in eng]
        n0 = gdata.shape[0]
        for e in eng:
            e.connect_all(handles)
            idx = [(e.read_lo + i) % n0 for i in range(e.n_local)]
            e.load(gdata.index_select(0, torch.as_tensor(idx, device=gdata.device)).contiguous())
        for phase, cnt in ((0, nF), (1, nU)):
            for _ in range(cnt):
                for e in eng:                    # iteration by iteration across the ranks, like cytvdn_denoise_sharded
                    e.iterate(1 if phase == 0 else 0, 0 if phase == 0 else 1)
        out = torch.empty_like(gdata)
        tot = np.zeros((max(n, 1), 3))
        for e in eng:
            e.store(out[e.valid_lo:e.valid_hi])
            tot[:n] += e.sums(n)
        with np.errstate(all="ignore"):
            return out, tot[:n, 0].copy(), (tot[:n, 1] / tot[:n, 2]).copy()
    finally:
        for e in eng:
            e.synchronize() if e.h else None
        for e in eng:
            e.close()


def denoise4D_sharded(shard, mu, iterations=10, FISTA=True, stopping_relative_change=None, *, plan: ShardPlan,
                      group=None, lam=None, overlap=True, return_state=False, schedule="fused", engine=False):
    """Sharded counterpart of ``denoise4D`` over torch.distributed / NCCL -- call it on every rank of the process group.
    (Round 1's schedule, kept for 2-D ``(wx, wy)`` grids, the reference-structured two-pass iteration and as the
    cross-check of the C-ABI engine; ``denoise4D_engine`` is the default path for 1-D splits.  ``engine`` must be False.)

    ``shard``: this rank's block INCLUDING its overlap planes (``plan.read_global`` of the global
    array), a contiguous CUDA tensor.  Returns ``(recon_local, b_norm, delta_recon)`` where
    ``recon_local[plan.owned_local]`` is this rank's part of the result and the two 1-D arrays are the
    global values (owned-voxel sums, all-reduced).  ``iterations`` may be ``[n_FISTA, n_plain]``.
    ``schedule``: ``"fused"`` (one pass and one exchange per iteration, needs a second set of
    accumulator arrays), ``"two_pass"`` (the reference's structure: two sweeps, two exchanges) or ``None`` /
    ``"auto"``: fused when the second state set fits on every rank.
    Boundary: Jia-Zhao (``BC_mode=2``) or, with a ``ShardPlan(..., periodic=True)``, periodic (``BC_mode=0``).
    """
    import torch
    import torch.distributed as dist
    assert not engine, "denoise4D_sharded is the NCCL path; use denoise4D_engine for the C-ABI engine"
    unaccelerated = not FISTA
    if type(iterations) in (list, tuple):
        FISTA, unaccelerated = True, True
        nF, nU = int(iterations[0]), int(iterations[1])
    else:
        nF, nU = int(iterations * FISTA), int(iterations * (not FISTA))
    n = nF + nU
    world = plan.world
    if schedule in (None, "auto"):
        # fused needs 1 + 2*(1 + 4*(1 or 2)) arrays next to the caller's shard; two_pass 1 + 4*(1 or 2)
        per = 4 * (2 if nF > 0 else 1)
        free_b, _ = torch.cuda.mem_get_info(shard.device)
        need = (1 + 2 * (1 + per)) * shard.numel() * shard.element_size()
        fits = torch.tensor([1 if need + (1 << 30) < free_b else 0], device=shard.device)
        if world > 1:
            dist.all_reduce(fits, op=dist.ReduceOp.MIN, group=group)      # every rank must take the same schedule
        schedule = "fused" if int(fits.item()) else "two_pass"
    fused = schedule == "fused"
    sh = CudaShard(plan, shard, mu, lam, fista=nF > 0, n_iter=n, fused=fused)
    one_d = plan.grid[1] == 1
    comm_stream = torch.cuda.Stream(device=shard.device) if (overlap and one_d and world > 1) else None
    ran = np.zeros(n, dtype=bool)
    tk = 1.0
    glob = torch.zeros((max(n, 1), 3), dtype=torch.float64, device=shard.device)
    for phase, count in ((0, nF), (1, nU)):
        for j in range(count):
            it = j if phase == 0 else nF + j
            tkr = 0.0
            if phase == 0:
                tkr, tk = fista_ratio(tk)
            if fused:
                _run_iteration_fused(sh, it, tkr, phase == 0, group, comm_stream)
            elif comm_stream is not None:
                _run_iteration_overlapped(sh, it, tkr, phase == 0, group, comm_stream)
            else:
                _run_iteration_simple(sh, it, tkr, phase == 0, group)
            ran[it] = True
            if stopping_relative_change is not None:          # needs the global delta now
                s = (sh.fused_local_sums() if fused else sh.local_sums())[it].clone()
                if world > 1:
                    dist.all_reduce(s, group=group)
                glob[it] = s
                dl = float(s[1] / s[2])
                if shard.dtype == torch.float32:
                    dl = float(np.float32(dl))
                if dl < stopping_relative_change:
                    break
    if stopping_relative_change is None and n > 0:
        glob = (sh.fused_local_sums() if fused else sh.local_sums()).clone()
        if world > 1:
            dist.all_reduce(glob, group=group)
    g = glob.cpu().numpy()
    dt = np.float32 if shard.dtype == torch.float32 else np.float64
    with np.errstate(all="ignore"):
        b_norm = np.where(ran, g[:n, 0], 0.0).astype(dt)
        delta = np.where(ran, g[:n, 1] / g[:n, 2], 0.0).astype(dt)
    torch.cuda.current_stream(shard.device).synchronize()
    if return_state:
        return sh.result(), b_norm, delta, sh
    return sh.result(), b_norm, delta


# ------------------------------------------------------------------------------------------------
# all ranks of a plan on ONE device, in lockstep (no process group): used to validate the sharded
# schedule -- box order, owned-range reductions, zero-wrap, plane indices -- with the real kernels
# when fewer GPUs than ranks are available.
# ------------------------------------------------------------------------------------------------
def emulate_on_one_device(gdata, mu, world, grid=None, iterations=10, FISTA=True, split_boxes=True, lam=None,
                          schedule="two_pass", periodic=False):
    """``gdata``: the GLOBAL array as a CUDA tensor.  Returns (assembled recon, b_norm, delta)."""
    import torch
    if type(iterations) in (list, tuple):
        nF, nU = int(iterations[0]), int(iterations[1])
    else:
        nF, nU = int(iterations * FISTA), int(iterations * (not FISTA))
    n = nF + nU
    plans = [ShardPlan(gdata.shape, world, r, grid, periodic) for r in range(world)]
    fused = schedule == "fused"
    shards = [CudaShard(p, p.extract(gdata).contiguous(), mu, lam, fista=nF > 0, n_iter=n, fused=fused)
              for p in plans]
    one_d = plans[0].grid[1] == 1

    def exchange_fused():
        sent = {}
        for s in shards:
            for op in s.plan.after_fused():
                if op.kind == "send":
                    sent[(s.plan.rank, op.peer, op.axis, op.side)] = plane(s.recon2, op.axis, op.index).clone()
        for s in shards:
            for op in s.plan.after_fused():
                if op.kind == "recv":
                    other = "hi" if op.side == "lo" else "lo"
                    plane(s.recon2, op.axis, op.index).copy_(sent.pop((op.peer, s.plan.rank, op.axis, other)))
        assert not sent

    def exchange(phase):
        sent = {}
        for s in shards:
            for op in (s.plan.after_a() if phase == "a" else s.plan.after_b()):
                if op.kind == "send":
                    sent[(s.plan.rank, op.peer, op.array, op.axis, op.side)] = plane(s.arrays[op.array], op.axis, op.index).clone()
        for s in shards:
            for op in (s.plan.after_a() if phase == "a" else s.plan.after_b()):
                if op.kind == "recv":
                    other = "hi" if op.side == "lo" else "lo"
                    plane(s.arrays[op.array], op.axis, op.index).copy_(sent.pop((op.peer, s.plan.rank, op.array, op.axis, other)))
        assert not sent

    tk = 1.0
    for phase_f, cnt in ((0, nF), (1, nU)):
        for j in range(cnt):
            it = j if phase_f == 0 else nF + j
            tkr = 0.0
            if phase_f == 0:
                tkr, tk = fista_ratio(tk)
            if fused:
                slots = {}
                for s in shards:
                    first, _ = s.plan.fused_boxes() if (split_boxes and one_d) else ([], None)
                    slot = 0
                    for box in first:
                        s.fused_step(it, slot, tkr, phase_f == 0, box)
                        slot += 4
                    slots[s.plan.rank] = slot
                if split_boxes and one_d:
                    exchange_fused()       # in the real schedule this runs under the "rest" sweep
                for s in shards:
                    rest = s.plan.fused_boxes()[1] if (split_boxes and one_d) else [(0, s.plan.local_shape[0])]
                    slot = slots[s.plan.rank]
                    for box in rest:
                        s.fused_step(it, slot, tkr, phase_f == 0, box)
                        slot += 4
                if not (split_boxes and one_d):
                    exchange_fused()
                for s in shards:
                    s.fused_swap()
                continue
            for phase in ("a", "b"):
                if split_boxes and one_d:
                    slots = {}
                    for s in shards:
                        first, _ = s.plan.a_boxes() if phase == "a" else s.plan.b_boxes()
                        slot = 0 if phase == "a" else 4
                        for box in first:
                            (s.half_step_a(it, slot, tkr, phase_f == 0, box) if phase == "a" else s.half_step_b(it, slot, box))
                            slot += 1 if phase == "a" else 2
                        slots[s.plan.rank] = slot
                    exchange(phase)        # in the real schedule this runs under the "rest" sweep
                    for s in shards:
                        _, rest = s.plan.a_boxes() if phase == "a" else s.plan.b_boxes()
                        slot = slots[s.plan.rank]
                        for box in rest:
                            (s.half_step_a(it, slot, tkr, phase_f == 0, box) if phase == "a" else s.half_step_b(it, slot, box))
                            slot += 1 if phase == "a" else 2
                else:
                    for s in shards:
                        (s.half_step_a(it, 0, tkr, phase_f == 0) if phase == "a" else s.half_step_b(it, 4))
                    exchange(phase)
    out = torch.empty_like(gdata)
    tot = torch.zeros((max(n, 1), 3), dtype=torch.float64, device=gdata.device)
    for s in shards:
        out[s.plan.owned_global] = s.result()[s.plan.owned_local]
        tot += s.fused_local_sums() if fused else s.local_sums()
    g = tot.cpu().numpy()
    with np.errstate(all="ignore"):
        return out, g[:n, 0].copy(), (g[:n, 1] / g[:n, 2]).copy()


# ------------------------------------------------------------------------------------------------
# "peer" schedule: no exchange step at all.  Every rank keeps ONLY its owned planes; the fused kernel reads
# the axis-0 halo (the lower neighbour's last recon plane, the upper neighbour's first recon / b_0 / d_0 planes)
# straight from the neighbours' HBM through CUDA-IPC peer pointers over NVLink.  One launch per iteration, one
# tiny all-reduce as the inter-iteration barrier (a rank may not overwrite a state set its neighbours still read).
# ------------------------------------------------------------------------------------------------
class _DevArray:
    """Lets torch view a raw device pointer (``__cuda_array_interface__``)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class PeerShard:
    """Device state of one rank for the peer schedule: owned planes only, one cytvdn_malloc arena
    (exportable through CUDA IPC), two state sets (the fused iteration is out of place)."""

    def __init__(self, plan: ShardPlan, owned, mu, lam=None, fista=True, n_iter=1, group=None, connect=True):
        import torch
        import torch.distributed as dist
        from . import _lib
        self.torch, self._lib, self.lib, self.dist, self.group = torch, _lib, _lib.load(), dist, group
        _lib.require_gpu()
        assert plan.grid[1] == 1, "the peer schedule splits scan axis 0 only"
        assert owned.is_cuda and owned.is_contiguous() and owned.dtype in (torch.float32, torch.float64)
        n_own = plan.valid[0][1] - plan.valid[0][0]
        assert tuple(owned.shape) == (n_own,) + tuple(plan.gshape[1:]), (tuple(owned.shape), plan.gshape)
        self.plan, self.fista, self.dev = plan, fista, owned.device
        dt = np.float32 if owned.dtype == torch.float32 else np.float64
        self.tdtype, elem = owned.dtype, (4 if dt == np.float32 else 8)
        mu = np.asarray(mu, dtype=dt)
        lam = (mu * 1.0 / 32.0) if lam is None else np.asarray(lam, dtype=dt)
        self.clip = (C.c_double * 4)(*[float(v) for v in (1.0 / lam)])
        self.w = (C.c_double * 4)(*[float(v) for v in (lam / mu).astype(dt)])
        self.code = 0 if dt == np.float32 else 1
        self.bc_mode = 0 if plan.periodic else 2
        # rows padded to the 16-byte vector width (cytvdn_step_opts.row_pitch)
        self.n3 = int(owned.shape[3])
        vwf = 16 // elem
        self.n3p = (self.n3 + vwf - 1) // vwf * vwf
        self.shape_p = (n_own,) + tuple(plan.gshape[1:3]) + (self.n3p,)
        self.plane_elems = int(plan.gshape[1]) * int(plan.gshape[2]) * self.n3p
        numel = n_own * self.plane_elems
        self.pitch = (numel * elem + 255) // 256 * 256                  # bytes between arrays in the arena
        per = 1 + 4 * (2 if fista else 1)                               # recon, b x4 (, d x4)
        self.n_arrays = 1 + 2 * per
        self.arena = C.c_void_p()
        _lib.check(self.lib.cytvdn_malloc(C.byref(self.arena), self.n_arrays * self.pitch))
        self.sh = (C.c_int64 * 4)(n_own, *[int(v) for v in plan.gshape[1:]])
        typestr = "<f4" if dt == np.float32 else "<f8"
        view = lambda k: torch.as_tensor(_DevArray(self.arena.value + k * self.pitch, self.shape_p, typestr), device=self.dev)
        # arena layout (identical on every rank): 0 orig | set s: 1 + s*per + (0 recon, 1..4 b, 5..8 d)
        self.per = per
        self.orig = view(0)
        self.sets = [{"recon": view(1 + s * per), "b": [view(1 + s * per + 1 + k) for k in range(4)],
                      "d": [view(1 + s * per + 5 + k) for k in range(4)] if fista else None} for s in range(2)]
        self.orig.zero_()
        self.orig[..., :self.n3].copy_(owned)
        self.sets[0]["recon"].copy_(self.orig)
        for s in range(2):
            for t in self.sets[s]["b"] + (self.sets[s]["d"] or []):
                t.zero_()
        self.cur = 0
        self.sums = torch.zeros((max(n_iter, 1), 4), dtype=torch.float64, device=self.dev)
        self.token = torch.zeros(1, dtype=torch.float32, device=self.dev)
        self.launches = 0
        self.peer_ptr, self.peer_info, self.local = {}, {}, False
        if connect:
            self.connect_ipc()

    def connect_local(self, shards):
        """All ranks live in THIS process (validation on one GPU): neighbours are plain pointers."""
        self.local = True
        for side, has in (("lo", self.plan.has_lo[0]), ("hi", self.plan.has_hi[0])):
            if has:
                r = self.plan.peer(0, -1 if side == "lo" else +1)
                self.peer_ptr[r] = C.c_void_p(shards[r].arena.value)
                self.peer_info[r] = (None, int(shards[r].sh[0]), shards[r].pitch)

    def connect_ipc(self):
        """Exchange CUDA-IPC handles of the arenas and map the neighbours' (one process per GPU)."""
        torch, dist, plan, _lib, group = self.torch, self.dist, self.plan, self._lib, self.group
        n_own = int(self.sh[0])
        if plan.world > 1:
            h = (C.c_ubyte * 64)()
            _lib.check(self.lib.cytvdn_ipc_get_handle(self.arena, h))
            info = [None] * plan.world
            dist.all_gather_object(info, (bytes(h), n_own, self.pitch), group=group)
            self.peer_info = info
            for side, has in (("lo", plan.has_lo[0]), ("hi", plan.has_hi[0])):
                if not has:
                    continue
                r = plan.peer(0, -1 if side == "lo" else +1)
                if r not in self.peer_ptr:
                    q = C.c_void_p()
                    ph = (C.c_ubyte * 64)(*info[r][0])
                    _lib.check(self.lib.cytvdn_ipc_open(ph, C.byref(q)))
                    self.peer_ptr[r] = q
            torch.cuda.synchronize(self.dev)
            dist.barrier(group=group)          # everybody's initial state is in place before anyone reads it

    def _array_ptr(self, base, pitch, s, which, k=0):
        idx = 1 + s * self.per + {"recon": 0, "b": 1 + k, "d": 5 + k}[which]
        return base + idx * pitch

    def step(self, it, tk_ratio, fista):
        """One fused iteration over the owned planes, halo read from the neighbours, then the barrier."""
        torch, plan = self.torch, self.plan
        st = torch.cuda.current_stream(self.dev).cuda_stream
        s, elem = self.cur, (4 if self.code == 0 else 8)
        src, dst = self.sets[s], self.sets[1 - s]
        o = self._lib.StepOpts()
        o.row_pitch = self.n3p
        o.flags = plan.jz_flags
        if plan.has_lo[0]:
            r = plan.peer(0, -1)
            n_r, pitch_r = self.peer_info[r][1], self.peer_info[r][2]
            o.peer_lo_recon = self._array_ptr(self.peer_ptr[r].value, pitch_r, s, "recon") + (n_r - 1) * self.plane_elems * elem
        if plan.has_hi[0]:
            r = plan.peer(0, +1)
            pitch_r = self.peer_info[r][2]
            o.peer_hi_recon = self._array_ptr(self.peer_ptr[r].value, pitch_r, s, "recon")
            o.peer_hi_b0 = self._array_ptr(self.peer_ptr[r].value, pitch_r, s, "b", 0)
            if fista:
                o.peer_hi_d0 = self._array_ptr(self.peer_ptr[r].value, pitch_r, s, "d", 0)
        elif plan.has_lo[0]:
            o.zero_wrap_mask = 1               # global upper edge: nothing beyond the last plane (Jia-Zhao)
        ptrs = lambda ts: (C.c_void_p * 4)(*[t.data_ptr() for t in ts])
        self._lib.check(self.lib.cytvdn_fused_iteration(
            4, self.sh, self.code, self.orig.data_ptr(), src["recon"].data_ptr(), dst["recon"].data_ptr(),
            ptrs(src["b"]), ptrs(dst["b"]), ptrs(src["d"]) if fista else None, ptrs(dst["d"]) if fista else None,
            float(tk_ratio), self.clip, self.w, self.bc_mode, self.sums[it].data_ptr(), C.byref(o), st))
        self.launches += 1
        self.cur = 1 - s
        if plan.world > 1 and not self.local:  # nobody starts the next iteration before everybody finished this one
            self.dist.all_reduce(self.token, group=self.group)

    def result(self):
        r = self.sets[self.cur]["recon"]
        return r if self.n3p == self.n3 else r[..., :self.n3]

    def close(self):
        if not self.local:
            for q in self.peer_ptr.values():
                self.lib.cytvdn_ipc_close(q)
        self.peer_ptr = {}
        if self.arena:
            self.torch.cuda.synchronize(self.dev)
            if self.plan.world > 1 and not self.local:
                self.dist.barrier(group=self.group)      # neighbours have unmapped it
            self.lib.cytvdn_free(self.arena)
            self.arena = None


def denoise4D_peer(owned, mu, iterations=10, FISTA=True, stopping_relative_change=None, *, plan: ShardPlan, group=None,
                   lam=None):
    """Sharded ``denoise4D`` with the peer schedule.  ``owned``: this rank's OWNED planes only
    (``global[plan.owned_global]``, contiguous CUDA tensor; 1-D split).  Returns ``(recon_owned, b_norm, delta)``."""
    import torch
    import torch.distributed as dist
    unaccelerated = not FISTA
    if type(iterations) in (list, tuple):
        FISTA, unaccelerated = True, True
        nF, nU = int(iterations[0]), int(iterations[1])
    else:
        nF, nU = int(iterations * FISTA), int(iterations * (not FISTA))
    n, world = nF + nU, plan.world
    sh = PeerShard(plan, owned, mu, lam, fista=nF > 0, n_iter=n, group=group)
    try:
        ran = np.zeros(n, dtype=bool)
        glob = torch.zeros((max(n, 1), 4), dtype=torch.float64, device=owned.device)
        tk = 1.0
        for phase, count in ((0, nF), (1, nU)):
            for j in range(count):
                it = j if phase == 0 else nF + j
                tkr = 0.0
                if phase == 0:
                    tkr, tk = fista_ratio(tk)
                sh.step(it, tkr, phase == 0)
                ran[it] = True
                if stopping_relative_change is not None:
                    s = sh.sums[it].clone()
                    if world > 1:
                        dist.all_reduce(s, group=group)
                    glob[it] = s
                    dl = float(s[1] / s[2])
                    if owned.dtype == torch.float32:
                        dl = float(np.float32(dl))
                    if dl < stopping_relative_change:
                        break
        if stopping_relative_change is None and n > 0:
            glob = sh.sums.clone()
            if world > 1:
                dist.all_reduce(glob, group=group)
        g = glob.cpu().numpy()
        dt = np.float32 if owned.dtype == torch.float32 else np.float64
        with np.errstate(all="ignore"):
            b_norm = np.where(ran, g[:n, 0], 0.0).astype(dt)
            delta = np.where(ran, g[:n, 1] / g[:n, 2], 0.0).astype(dt)
        out = sh.result().clone()
        torch.cuda.current_stream(owned.device).synchronize()
        return out, b_norm, delta
    finally:
        sh.close()


def emulate_peer_on_one_device(gdata, mu, world, iterations=10, FISTA=True, periodic=False, lam=None):
    """All ranks of the peer schedule on ONE device (neighbours are plain pointers, launches in lockstep on one
    stream).  Returns (assembled recon, b_norm, delta)."""
    import torch
    if type(iterations) in (list, tuple):
        nF, nU = int(iterations[0]), int(iterations[1])
    else:
        nF, nU = int(iterations * FISTA), int(iterations * (not FISTA))
    n = nF + nU
    plans = [ShardPlan(gdata.shape, world, r, None, periodic) for r in range(world)]
    shards = [PeerShard(p, gdata[p.owned_global].contiguous(), mu, lam, fista=nF > 0, n_iter=n, connect=False)
              for p in plans]
    try:
        for s in shards:
            s.connect_local(shards)
        tk = 1.0
        for phase, cnt in ((0, nF), (1, nU)):
            for j in range(cnt):
                it = j if phase == 0 else nF + j
                tkr = 0.0
                if phase == 0:
                    tkr, tk = fista_ratio(tk)
                for s in shards:
                    s.step(it, tkr, phase == 0)
        out = torch.empty_like(gdata)
        tot = torch.zeros((max(n, 1), 4), dtype=torch.float64, device=gdata.device)
        for s in shards:
            out[s.plan.owned_global] = s.result()
            tot += s.sums
        g = tot.cpu().numpy()
        with np.errstate(all="ignore"):
            return out, g[:n, 0].copy(), (g[:n, 1] / g[:n, 2]).copy()
    finally:
        for s in shards:
            s.close()
