"""Host-side logic of the scan-axis sharding (cytvdn_b200/sharded.py) on CPU.

The partition, the halo plane indices, the zero-wrap rule and ``halo_exchange`` are the product
code; the arithmetic is done by the oracle's kernels (tests/sharded_cpu_driver.py).  The oracle for
the sharded result is the single-process ``denoise4D`` on the unsharded array (SURVEY.md section 8c:
`mpi.py`'s as-written exchange does not reproduce it and the reference has no sharded FISTA)."""
import os
import socket
import tempfile

import numpy as np
import pytest

from cytvdn_b200.sharded import ShardPlan, mpi_grid
from oracle import tv_oracle as O
from tests import sharded_cpu_driver as drv


def counts(rng, shape):
    return rng.poisson(rng.uniform(20, 400, shape)).astype(np.float32)


def test_mpi_grid_heuristic_matches_reference():
    # square scan -> 2 ranks (1,2), 4 -> (2,2), 8 -> (2,4)  (mpi.py:131-150)
    assert mpi_grid((256, 256), 2) == (1, 2)
    assert mpi_grid((256, 256), 4) == (2, 2)
    assert mpi_grid((256, 256), 8) == (2, 4)
    assert mpi_grid((1024, 64), 4) == (4, 1)
    assert mpi_grid((10, 10), 1) == (1, 1)


@pytest.mark.parametrize("gshape,world,grid", [
    ((12, 10, 3, 4), 1, None), ((12, 10, 3, 4), 2, None), ((13, 10, 3, 4), 4, None), ((12, 10, 3, 4), 4, (2, 2)),
    ((13, 11, 3, 4), 6, (3, 2)), ((16, 16, 2, 2), 8, "mpi"), ((9, 20, 2, 2), 4, (1, 4)),
])
def test_plan_tiles_the_global_array(gshape, world, grid):
    cover = np.zeros(gshape[:2], dtype=int)
    for r in range(world):
        p = ShardPlan(gshape, world, r, grid)
        cover[p.owned_global] += 1
        # stored block = owned block + one plane towards each existing neighbour (mpi.py:173-196)
        for k in range(2):
            assert p.read[k][0] == p.valid[k][0] - (1 if p.has_lo[k] else 0)
            assert p.read[k][1] == p.valid[k][1] + (1 if p.has_hi[k] else 0)
            assert p.own_hi[k] - p.own_lo[k] == p.valid[k][1] - p.valid[k][0]
        # every send has its matching receive on the peer, with the corrected plane indices
        for phase in ("after_a", "after_b"):
            for op in getattr(p, phase)():
                q = ShardPlan(gshape, world, op.peer, grid)
                mirror = [o for o in getattr(q, phase)() if o.peer == r and o.array == op.array and o.axis == op.axis
                          and o.kind != op.kind]
                assert len(mirror) == 1
                # the global index of the plane sent equals the global index of the plane received
                g_here = p.read[op.axis][0] + op.index
                g_there = q.read[op.axis][0] + mirror[0].index
                assert g_here == g_there
                if op.kind == "send":          # only owned planes are ever sent
                    assert p.valid[op.axis][0] <= g_here < p.valid[op.axis][1]
    assert np.all(cover == 1)


def test_plan_rejects_empty_tiles():
    with pytest.raises(ValueError, match="empty"):
        ShardPlan((5, 8, 2, 2), 4, 3, None)        # ceil(5/4)=2 planes per tile -> tile 3 empty


def test_boxes_partition_axis0():
    for world, rank in ((1, 0), (2, 0), (2, 1), (4, 2)):
        p = ShardPlan((16, 6, 2, 2), world, rank)
        n = p.local_shape[0]
        fa, ra = p.a_boxes()
        seen = sorted(i for lo, hi in fa + ra for i in range(lo, hi))
        assert seen == list(range(n))
        fb, rb = p.b_boxes()
        seen = sorted(i for lo, hi in fb + rb for i in range(lo, hi))
        assert seen == list(range(n - (1 if p.has_hi[0] else 0)))       # received overlap plane skipped
        if p.has_lo[0]:
            assert (1, 2) in fb and (0, 1) in fa
        if p.has_hi[0]:
            assert (n - 2, n - 1) in fa


@pytest.mark.parametrize("world,grid", [(1, None), (2, None), (2, (1, 2)), (3, None), (4, None), (4, (2, 2)),
                                        (6, (3, 2)), (6, (2, 3)), (8, "mpi")])
@pytest.mark.parametrize("fista", [True, False])
def test_in_process_emulation_equals_single_process(world, grid, fista):
    """All ranks emulated sequentially (uneven splits included): assembled recon == unsharded recon."""
    O.set_threads(O.max_threads())
    rng = np.random.default_rng(17 + world)
    gshape = (13, 10, 6, 5) if world != 8 else (13, 14, 6, 5)
    data = counts(rng, gshape)
    mu = np.array([1, 1, .5, .5], dtype=np.float32)
    K = O.PortKernels("D")
    n = 12
    ref = O.denoise4D(data, mu, n, fista, quiet=True, kernels=K, scalars="D")
    got, bn, dl = drv.run_in_process(data, mu, world, grid, n if fista else 0, 0 if fista else n, K)
    assert np.array_equal(got, ref[0]), f"max diff {np.abs(got - ref[0]).max()}"
    np.testing.assert_allclose(bn, ref[1], rtol=1e-12)
    np.testing.assert_allclose(dl, ref[2], rtol=1e-10)


def test_as_written_mpi_exchange_is_not_equivalent():
    """Documents why the plane indices differ from mpi.py:325,408 -- sending the overlap planes
    (acc[-1] / recon[0]) does NOT reproduce the single-process result."""
    O.set_threads(O.max_threads())
    rng = np.random.default_rng(3)
    data = counts(rng, (12, 10, 6, 5))
    mu = np.array([1, 1, .5, .5], dtype=np.float32)
    K = O.PortKernels("D")
    ref = O.denoise4D(data, mu, 8, False, quiet=True, kernels=K, scalars="D")[0]
    plans = [ShardPlan(data.shape, 2, r) for r in range(2)]
    sh = [drv.CpuShard(p, np.ascontiguousarray(data[p.read_global]), mu, K, fista=False) for p in plans]
    for _ in range(8):
        for s in sh:
            s.half_step_a(0.0, False)
        sh[1].b[0][0] = sh[0].b[0][-1]            # as written: mpi.py:325 sends acc0[-1]
        for s in sh:
            s.K.datacube_update(s.orig, s.recon, s.b, s.w, 2)
        sh[0].recon[-1] = sh[1].recon[0]          # as written: mpi.py:408 sends recon[0]
    got = np.concatenate([sh[0].recon[:-1], sh[1].recon[1:]])
    assert float(np.abs(got - ref).max()) > 0.1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("grid,periodic", [(None, False), ((1, 2), False), (None, True), ((1, 2), True)])
def test_gloo_world2_halo_exchange(grid, periodic):
    """Two real ranks over gloo: the product halo_exchange (axis-0 planes contiguous, axis-1 planes
    staged) and the all-reduced owned sums.  Periodic with two tiles: both neighbours of a rank are the SAME peer,
    i.e. two messages per direction between one pair in one batch -- they must not be swapped."""
    import torch.multiprocessing as mp
    gshape, seed, nF, nU = (10, 9, 5, 4), 99, 6, 3
    with tempfile.TemporaryDirectory() as tmp:
        port = _free_port()
        mp.spawn(drv.run_distributed_rank, args=(2, port, grid, gshape, seed, nF, nU, tmp, periodic), nprocs=2, join=True)
        rng = np.random.default_rng(seed)
        data = rng.poisson(rng.uniform(20, 400, gshape)).astype(np.float32)
        mu = np.array([1, 1, .5, .5], dtype=np.float32)
        O.set_threads(O.max_threads())
        ref = O.denoise4D(data, mu, [nF, nU], BC_mode=0 if periodic else 2, quiet=True, kernels=O.PortKernels("D"),
                          scalars="D")
        got = np.empty_like(data)
        for r in range(2):
            z = np.load(os.path.join(tmp, f"rank{r}.npz"))
            lo = z["lo"]
            blk = z["recon"]
            got[lo[0]:lo[0] + blk.shape[0], lo[1]:lo[1] + blk.shape[1]] = blk
            sums = z["sums"]
        assert np.array_equal(got, ref[0])
        np.testing.assert_allclose(sums[:, 0], ref[1], rtol=1e-12)
        np.testing.assert_allclose(sums[:, 1] / sums[:, 2], ref[2], rtol=1e-10)


@pytest.mark.parametrize("world,grid", [(1, None), (2, None), (3, None), (4, None), (2, (1, 2)), (4, (2, 2)), (6, (3, 2)), (6, (2, 3))])
@pytest.mark.parametrize("fista", [True, False])
def test_periodic_in_process_emulation_equals_single_process(world, grid, fista):
    """BC_mode=0 sharded: the first and last tile of a split axis are neighbours, the split axes use the
    Jia-Zhao boundary inside the block, the exchange does the wrap.  Two tiles on an axis = both neighbours are the
    same rank."""
    O.set_threads(O.max_threads())
    rng = np.random.default_rng(40 + world)
    data = counts(rng, (13, 10, 6, 5))
    mu = np.array([1, 1, .5, .5], dtype=np.float32)
    K = O.PortKernels("D")
    n = 12
    ref = O.denoise4D(data, mu, n, fista, BC_mode=0, quiet=True, kernels=K, scalars="D")
    got, bn, dl = drv.run_in_process(data, mu, world, grid, n if fista else 0, 0 if fista else n, K, periodic=True)
    assert np.array_equal(got, ref[0]), f"max diff {np.abs(got - ref[0]).max()}"
    np.testing.assert_allclose(bn, ref[1], rtol=1e-12)
    np.testing.assert_allclose(dl, ref[2], rtol=1e-10)


def test_periodic_plan_wraps():
    p0 = ShardPlan((12, 10, 3, 4), 3, 0, None, periodic=True)
    p2 = ShardPlan((12, 10, 3, 4), 3, 2, None, periodic=True)
    assert p0.has_lo[0] and p0.has_hi[0] and p0.peer(0, -1) == 2 and p2.peer(0, +1) == 0
    assert p0.read_indices(0) == [11, 0, 1, 2, 3, 4] and p2.read_indices(0) == [7, 8, 9, 10, 11, 0]
    assert p0.zero_wrap_mask == 0 and p0.jz_flags == 1 << 8 and not p0.has_lo[1]
    g = np.arange(12 * 10 * 3 * 4, dtype=np.float32).reshape(12, 10, 3, 4)
    assert np.array_equal(p0.extract(g), g[[11, 0, 1, 2, 3, 4]])
    one = ShardPlan((12, 10, 3, 4), 1, 0, None, periodic=True)
    assert not one.has_lo[0] and one.jz_flags == 0 and one.local_shape == (12, 10, 3, 4)


def _check_plan_family(gshape, world, grid, periodic):
    """Invariants of all ranks' plans together: the owned blocks tile the array, every send has exactly one
    matching receive on the peer and both name the same global plane, only owned planes are sent."""
    plans = []
    for r in range(world):
        try:
            plans.append(ShardPlan(gshape, world, r, grid, periodic))
        except ValueError as e:                     # a split that leaves a tile empty is refused for every rank alike
            assert "empty" in str(e) or "grid" in str(e) or "planes" in str(e), e
            return False
    cover = np.zeros(gshape[:2], dtype=int)
    for p in plans:
        cover[p.owned_global] += 1
        assert tuple(p.extract(np.zeros(gshape, np.float32)).shape) == tuple(p.local_shape)
    assert np.all(cover == 1)
    for phase in ("after_a", "after_b", "after_fused"):
        sends, recvs = {}, {}
        for p in plans:
            for op in getattr(p, phase)():
                gidx = p.read_indices(op.axis)[op.index]          # global index of the plane
                key = (min(p.rank, op.peer), max(p.rank, op.peer), op.array, op.axis, gidx,
                       p.rank if op.kind == "send" else op.peer)
                (sends if op.kind == "send" else recvs).setdefault(key, 0)
                (sends if op.kind == "send" else recvs)[key] += 1
                if op.kind == "send":
                    lo, hi = p.valid[op.axis]
                    assert lo <= gidx < hi, "only owned planes are sent"
        assert sends == recvs, f"{phase}: unmatched {set(sends) ^ set(recvs)}"
    return True


def test_plan_invariants_random():
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=150, deadline=None)
    @given(st.integers(2, 40), st.integers(2, 40), st.integers(1, 8), st.booleans(), st.sampled_from(["1d", "mpi", "2d"]),
           st.integers(1, 4))
    def run(n0, n1, world, periodic, kind, wx):
        gshape = (n0, n1, 2, 3)
        if kind == "1d":
            grid = None
        elif kind == "mpi":
            grid = "mpi"
        else:
            if world % wx:
                return
            grid = (wx, world // wx)
        _check_plan_family(gshape, world, grid, periodic)

    run()
