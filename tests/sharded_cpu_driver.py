"""CPU stand-in for ``cytvdn_b200.sharded.CudaShard`` used ONLY by the tests: the same ShardPlan,
HaloOps and ``halo_exchange`` (torch.distributed, gloo) as the product path, with the oracle's
kernels doing the arithmetic.  It lets the host-side sharding logic be checked without a GPU."""
from __future__ import annotations

import numpy as np

from cytvdn_b200.sharded import ShardPlan, fista_ratio, plane


class CpuShard:
    def __init__(self, plan: ShardPlan, shard: np.ndarray, mu, K, fista=True):
        assert shard.shape == tuple(plan.local_shape)
        self.plan, self.K, self.fista = plan, K, fista
        self.orig = np.ascontiguousarray(shard)
        dt = shard.dtype
        mu = np.asarray(mu, dtype=dt)
        lam = mu * 1.0 / 32.0
        self.clip = 1.0 / lam
        self.w = (lam / mu).astype(dt)
        self.recon = self.orig.copy()
        self.b = [np.zeros_like(self.orig) for _ in range(4)]
        self.d = [np.zeros_like(self.orig) for _ in range(4)] if fista else None
        self.arrays = {"b0": self.b[0], "b1": self.b[1], "recon": self.recon}

    def _own(self, x):
        return x[self.plan.owned_local]

    def half_step_a(self, tkr, fista):
        for ax in range(4):
            # periodic plan: axes that are split use Jia-Zhao inside the block, the others wrap locally
            bc = 0 if (self.plan.periodic and not (ax < 2 and self.plan.grid[ax] > 1)) else 2
            self.K.accumulator_update(self.recon, self.b[ax], self.d[ax] if fista else None, tkr, ax, self.clip[ax], bc)
        return float(sum(np.abs(self._own(x), dtype=np.float64).sum() for x in self.b))

    def half_step_b(self):
        """Reference kernel on the local block; ``zero_wrap`` emulated by appending a plane (zeros for
        the accumulator of that axis) so that the forward neighbour of the last plane is 0."""
        mask = self.plan.zero_wrap_mask
        f, u, bs = self.orig, self.recon, self.b
        pad = [(0, 1 if (mask >> k) & 1 else 0) for k in range(2)] + [(0, 0), (0, 0)]
        old = u.copy()
        if mask:
            fe = np.pad(f, pad, mode="wrap")
            ue = np.pad(u, pad, mode="wrap")
            be = []
            for k, x in enumerate(bs):
                y = np.pad(x, pad, mode="wrap")
                if k < 2 and (mask >> k) & 1:
                    if k == 0:
                        y[-1] = 0
                    else:
                        y[:, -1] = 0
                be.append(np.ascontiguousarray(y))
            ue = np.ascontiguousarray(ue)
            self.K.datacube_update(np.ascontiguousarray(fe), ue, be, self.w, 2)
            u[...] = ue[:u.shape[0], :u.shape[1]]
        else:
            self.K.datacube_update(f, u, bs, self.w, 2)
        dl = float(np.abs((self._own(u) - self._own(old)).astype(np.float64)).sum())
        on = float(np.abs(self._own(old), dtype=np.float64).sum())
        return dl, on


def apply_ops_in_process(shards, phase):
    """Sequential emulation of the exchange: copy planes between the ranks' states directly."""
    sends = {}
    for sh in shards:
        ops = sh.plan.after_a() if phase == "a" else sh.plan.after_b()
        for op in ops:
            if op.kind == "send":
                sends[(sh.plan.rank, op.peer, op.array, op.axis, op.side)] = plane(sh.arrays[op.array], op.axis, op.index).copy()
    for sh in shards:
        ops = sh.plan.after_a() if phase == "a" else sh.plan.after_b()
        for op in ops:
            if op.kind == "recv":
                other = "hi" if op.side == "lo" else "lo"
                plane(sh.arrays[op.array], op.axis, op.index)[...] = sends.pop((op.peer, sh.plan.rank, op.array, op.axis, other))
    assert not sends, "unmatched sends"


def run_in_process(gdata, mu, world, grid, n_fista, n_plain, K, periodic=False):
    """All ranks in one process.  Returns (assembled recon, b_norm, delta)."""
    plans = [ShardPlan(gdata.shape, world, r, grid, periodic) for r in range(world)]
    shards = [CpuShard(p, np.ascontiguousarray(p.extract(gdata)), mu, K, fista=n_fista > 0) for p in plans]
    bn, dl = [], []
    tk = 1.0
    for phase, cnt in ((0, n_fista), (1, n_plain)):
        for _ in range(cnt):
            tkr = 0.0
            if phase == 0:
                tkr, tk = fista_ratio(tk)
            bn.append(sum(s.half_step_a(tkr, phase == 0) for s in shards))
            apply_ops_in_process(shards, "a")
            parts = [s.half_step_b() for s in shards]
            apply_ops_in_process(shards, "b")
            dl.append(sum(p[0] for p in parts) / sum(p[1] for p in parts))
    out = np.empty_like(gdata)
    for s in shards:
        out[s.plan.owned_global] = s.recon[s.plan.owned_local]
    return out, np.array(bn), np.array(dl)


def run_distributed_rank(rank, world, port, grid, gshape, seed, n_fista, n_plain, outdir, periodic=False):
    """One gloo rank (spawned by the test): product ``halo_exchange`` + oracle kernels."""
    import os
    import torch
    import torch.distributed as dist
    from oracle import tv_oracle as O
    from cytvdn_b200.sharded import halo_exchange
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        O.set_threads(1)
        rng = np.random.default_rng(seed)
        gdata = rng.poisson(rng.uniform(20, 400, gshape)).astype(np.float32)   # every rank builds the same array
        mu = np.array([1, 1, .5, .5], dtype=np.float32)
        plan = ShardPlan(gshape, world, rank, grid, periodic)
        sh = CpuShard(plan, np.ascontiguousarray(plan.extract(gdata)), mu, O.PortKernels("D"), fista=n_fista > 0)
        tensors = {k: torch.from_numpy(v) for k, v in sh.arrays.items()}      # share memory with the numpy state
        sums = []
        tk = 1.0
        for phase, cnt in ((0, n_fista), (1, n_plain)):
            for _ in range(cnt):
                tkr = 0.0
                if phase == 0:
                    tkr, tk = fista_ratio(tk)
                a = sh.half_step_a(tkr, phase == 0)
                works, unpack = halo_exchange(plan.after_a(), tensors)
                for w in works:
                    w.wait()
                unpack()
                dl, on = sh.half_step_b()
                works, unpack = halo_exchange(plan.after_b(), tensors)
                for w in works:
                    w.wait()
                unpack()
                t = torch.tensor([a, dl, on], dtype=torch.float64)
                dist.all_reduce(t)
                sums.append(t.numpy().copy())
        np.savez(os.path.join(outdir, f"rank{rank}.npz"), recon=sh.recon[plan.owned_local],
                 lo=np.array([plan.valid[0][0], plan.valid[1][0]]), sums=np.array(sums))
    finally:
        dist.destroy_process_group()
