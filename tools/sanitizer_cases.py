"""Small-shape workout of every kernel family for compute-sanitizer (round-1 review "Hygiene"):

    compute-sanitizer --tool memcheck  python tools/sanitizer_cases.py
    compute-sanitizer --tool racecheck python tools/sanitizer_cases.py

Odd row lengths (8-byte / scalar paths and padded rows), length-1 axes, sweep boxes with owned ranges, narrow strips,
both schedules, half-isotropic pairs, the mirror boundary, reference_data (in-pass SSE), the PCIe pipeline, the
out-of-core tiles, peer pointers (all ranks on one device) and the shard engine.  Results are checked against the
two-pass schedule so that a silent corruption would also fail here.  tools/sanitize.sh runs both tools under gpurun.
"""
import os
import sys

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
os.environ.setdefault("CYTVDN_L2_BUDGET_MB", "0.05")           # narrow strips: several strips even on tiny arrays
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

import cytvdn_b200 as tv
from cytvdn_b200 import sharded


def data(shape, dt="float32", seed=0):
    rng = np.random.default_rng(seed)
    return rng.poisson(rng.uniform(20, 500, shape)).astype(dt)


def main():
    n = 0
    mu4 = lambda dt: np.array([1, 1, .5, .5], dtype=dt)
    for shape, dt in (((5, 4, 6, 16), "float32"), ((3, 5, 4, 13), "float32"), ((4, 3, 5, 6), "float32"),
                      ((1, 4, 1, 8), "float32"), ((4, 3, 5, 7), "float64"), ((2, 2, 2, 2), "float64")):
        x = data(shape, dt)
        for kw in (dict(), dict(isotropic_R=True, isotropic_Q=True), dict(BC_mode=0), dict(BC_mode=3)):
            if kw.get("BC_mode") == 3 and min(shape) < 2:
                continue
            for pad in ("1", "0"):
                os.environ["CYTVDN_PAD_ROWS"] = pad
                a = tv.denoise4D(x, mu4(dt), [4, 3], True, quiet=True, schedule="two_pass", **kw)
                vec = pad == "1" or shape[3] % (4 if dt == "float32" else 2) == 0
                auto = kw.get("BC_mode") == 3 or (kw.get("isotropic_R") and not vec)     # variants that need aligned rows
                b = tv.denoise4D(x, mu4(dt), [4, 3], True, quiet=True, schedule=None if auto else "fused", **kw)
                assert np.array_equal(a[0], b[0]), (shape, dt, kw, pad)
                n += 2
        os.environ["CYTVDN_PAD_ROWS"] = "1"
        r = tv.denoise4D(x, mu4(dt), 5, True, 1e-3, quiet=True, reference_data=x * 0.9)
        assert len(r) == 4
        n += 1
    x3 = data((6, 5, 37))
    a = tv.denoise3D(x3, np.array([1, 1, .5], np.float32), 6, FISTA=True, quiet=True, schedule="two_pass")
    b = tv.denoise3D(x3, np.array([1, 1, .5], np.float32), 6, FISTA=True, quiet=True, schedule="fused")
    assert np.array_equal(a[0], b[0])
    # step functions on caller arrays (scalar / 8-byte / 16-byte paths), boxes through the sharded emulations
    for shape in ((4, 5, 6, 9), (4, 5, 6, 10), (4, 5, 6, 12)):
        u, bb, dd = data(shape), np.zeros(shape, np.float32), np.zeros(shape, np.float32)
        tv.accumulator_update_4D_FISTA(u, bb, dd, 0.3, 1, 32.0, 0)
        tv.iso_accumulator_update_4D(u, bb, dd, 0, 3, 32.0)
        tv.datacube_update_4D(u, u.copy(), bb, dd, bb, dd, np.full(4, 1 / 32, np.float32))
        tv.sum_square_error_4D(u, bb)
        n += 4
    g = torch.from_numpy(data((11, 5, 6, 12))).cuda()
    ref = tv.denoise4D(g, mu4("float32"), 6, True, quiet=True)[0]
    for sched in ("fused", "two_pass"):
        got = sharded.emulate_on_one_device(g, mu4("float32"), 3, None, 6, True, True, schedule=sched)[0]
        assert torch.equal(got, ref)
    assert torch.equal(sharded.emulate_peer_on_one_device(g, mu4("float32"), 3, 6, True)[0], ref)
    # (the shard engine's spin-wait kernels need truly concurrent streams, which the sanitizer does not give several
    #  ranks sharing one device; its sweeps are the fused kernel with boxes and owned-only stores, exercised here:)
    e = sharded.EngineShard(g.shape, 1, 0, mu4("float32"), None, np.float32, fista=True, max_iters=6)
    e.connect_all([e.export()])
    e.load(g)
    e.iterate(6, 0)
    got = torch.empty_like(g)
    e.store(got)
    e.close()
    assert torch.equal(got, ref)
    # PCIe pipeline and out-of-core tiles
    x = data((12, 4, 6, 12))
    ref = tv.denoise4D(x, mu4("float32"), [5, 3], True, quiet=True, schedule="two_pass")[0]
    os.environ["CYTVDN_PIPELINE"] = "4"
    assert np.array_equal(tv.denoise4D(x, mu4("float32"), [5, 3], True, quiet=True)[0], ref)
    del os.environ["CYTVDN_PIPELINE"]
    os.environ["CYTVDN_STREAM_BUDGET_MB"] = str(8 * 2.5 * 10 * 4 * 6 * 12 * 4 / 1048576.0)
    assert np.array_equal(tv.denoise4D(x, mu4("float32"), [5, 3], True, quiet=True)[0], ref)
    del os.environ["CYTVDN_STREAM_BUDGET_MB"]
    torch.cuda.synchronize()
    print(f"sanitizer cases ok ({n} + sharded / pipeline / out-of-core cases)", flush=True)


if __name__ == "__main__":
    main()
