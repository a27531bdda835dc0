"""The C-ABI shard engine (csrc/cytvdn_shard.cu; cyTVDN/mpi.py:130-210, :314-438 behind `cytvdn_shard_*` and
`cytvdn_denoise_sharded`): sharded result == single-GPU result, bit for bit.

The GPU cases run in subprocesses (tests/engine_driver.py): CUDA_DEVICE_MAX_CONNECTIONS must be set before the context
exists when several ranks share one device, and a dead-locked exchange is killed by the timeout instead of hanging
the suite.  CPU part: the engine's partition equals ShardPlan's (and therefore mpi.py:161-196)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "tests", "engine_driver.py")


def _run(cmd, timeout=600):
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    lines = [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]
    assert p.returncode == 0, (p.returncode, lines, [l for l in p.stderr.splitlines() if "Error" in l or "error" in l or "assert" in l][-12:])
    assert lines and all(l["ok"] for l in lines), lines
    return lines


@pytest.mark.gpu
def test_engine_all_ranks_on_one_device():
    """Every rank of a plan in one process on cuda:0 (handles resolve to plain pointers): the engine's box order,
    flags, pushes and owned-only stores, and the C ABI's single-process loop `cytvdn_denoise_sharded`
    (`tv.denoise4D(devices=[0, 0, ...])`) -- uneven splits, one plane per rank, odd rows, periodic, hybrid counts,
    float64, early stopping."""
    lines = _run([sys.executable, DRIVER, "one_device"], timeout=300)
    assert len(lines) == 17


@pytest.mark.gpu
def test_engine_cuda_ipc_two_processes_sharing_a_gpu():
    """One process per rank, arenas mapped into the neighbour with CUDA IPC, pushes and flags across processes -- on
    ONE GPU (two contexts time-slice; gloo carries the handles), so the multi-process path is covered by the
    single-GPU tier too."""
    lines = _run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                  "--master-addr", "127.0.0.1", "--master-port", "29533", DRIVER, "ipc", "--share-gpu"], timeout=300)
    assert len(lines) >= 11


@pytest.mark.gpu
def test_engine_cuda_ipc_one_process_per_gpu():
    """The torchrun layout of bench.py: one rank per GPU, NCCL group, CUDA IPC over NVLink (needs >= 2 GPUs)."""
    import cytvdn_b200 as tv
    n = tv.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 4 if n >= 4 else 2
    _run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
          "--master-addr", "127.0.0.1", "--master-port", "29534", DRIVER, "ipc"], timeout=300)


@pytest.mark.gpu
def test_single_process_loop_over_real_devices():
    """`tv.denoise4D(devices=[0, 1, ...])`: the C ABI's loop over distinct GPUs (peer access, no IPC)."""
    import cytvdn_b200 as tv
    n = tv.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    rng = np.random.default_rng(2)
    data = rng.poisson(rng.uniform(20, 500, (16, 8, 16, 32))).astype(np.float32)
    mu = np.array([1, 1, .5, .5], dtype=np.float32)
    ref = tv.denoise4D(data, mu, 20, True, quiet=True)
    out = tv.denoise4D(data, mu, 20, True, quiet=True, devices=list(range(min(n, 4))))
    assert np.array_equal(out[0], ref[0])
    np.testing.assert_allclose(out[2], ref[2], rtol=1e-4)


def test_engine_partition_matches_shardplan():
    """plan_1d of the engine == ShardPlan == mpi.py:161-196 (checked through the Python mirror: the engine itself
    needs a device to be created)."""
    from cytvdn_b200 import sharded
    for n0 in (5, 12, 13, 64, 129):
        for world in (1, 2, 3, 4, 5, 8):
            n = -(-n0 // world)
            if (world - 1) * n >= n0:
                continue
            for periodic in (False, True):
                for rank in range(world):
                    p = sharded.ShardPlan((n0, 4, 4, 4), world, rank, None, periodic)
                    has_lo, has_hi = (rank > 0), (rank < world - 1)
                    if periodic and world > 1:
                        has_lo = has_hi = True
                    lo, hi = rank * n, min((rank + 1) * n, n0)
                    assert p.valid[0] == (lo, hi) and p.has_lo[0] == has_lo and p.has_hi[0] == has_hi
                    assert p.local_shape[0] == (hi - lo) + has_lo + has_hi
                    assert p.own_lo[0] == int(has_lo) and p.own_hi[0] == p.local_shape[0] - int(has_hi)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,dt,iters,fista,kw,ndev,planes", [
    ((40, 6, 8, 12), "float32", [9, 5], True, {}, 3, 12),            # several tiles per device, FISTA -> plain inside a pass
    ((37, 5, 6, 13), "float32", 11, True, {}, 2, 10),                # odd rows (padded), uneven tile sizes
    ((30, 4, 6, 8), "float64", 8, False, {}, 2, 8),                  # unaccelerated, float64
    ((33, 4, 6, 8), "float32", 10, True, {"isotropic_R": True, "isotropic_Q": True}, 3, 9),
    ((26, 4, 6, 8), "float32", 9, True, {"BC_mode": 3}, 2, 8),       # clamped mirror
    ((9, 4, 6, 8), "float32", 7, True, {}, 4, 8),                    # fewer tiles than devices
    ((24, 4, 6, 8), "float32", 30, True, {}, 1, 8),                  # one device through the sharded entry
])
def test_sharded_out_of_core_equals_in_core(shape, dt, iters, fista, kw, ndev, planes, monkeypatch):
    """`cytvdn_denoise_sharded_streamed` (config 5 on fewer GPUs than its state fits: README.md:104-120 of the
    reference): the out-of-core tiles of every pass dealt to several devices -- here all of them cuda:0, one host thread
    each -- with the K halo planes a device needs from the neighbouring range snapshotted before anybody writes back.
    Reconstruction bit-identical to the in-core single-GPU run; bnorm / delta to 1e-6."""
    import cytvdn_b200 as tv
    rng = np.random.default_rng(sum(shape))
    data = rng.poisson(rng.uniform(20, 500, shape)).astype(dt)
    mu = np.array([1, 1, .5, .5], dtype=dt)
    ref = tv.denoise4D(data, mu, iters, fista, quiet=True, schedule="two_pass", **kw)
    elem = 4 if dt == "float32" else 8
    n3p = -(-shape[3] // (16 // elem)) * (16 // elem)
    plane = shape[1] * shape[2] * n3p * elem
    arrays = 2 + 4 * (2 if fista else 1)
    # per-device budget that gives `planes` planes per slot (2.75 slots-worth for several devices, 2.5 for one)
    budget = (2.75 if ndev > 1 else 2.5) * arrays * plane * planes + 4096
    monkeypatch.setenv("CYTVDN_STREAM_BUDGET_MB", repr(budget / 1048576.0))
    tm = {}
    got = tv.denoise4D(data, mu, iters, fista, quiet=True, schedule="streamed", devices=[0] * ndev, timing=tm, **kw)
    assert tm["schedule"] == "streamed" and tm["devices"] == ndev
    assert np.array_equal(got[0], ref[0]), float(np.abs(got[0] - ref[0]).max())
    np.testing.assert_allclose(got[1].astype(np.float64), ref[1].astype(np.float64), rtol=1e-5)
    np.testing.assert_allclose(got[2].astype(np.float64), ref[2].astype(np.float64), rtol=1e-5)
