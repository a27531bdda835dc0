#!/usr/bin/env python
"""Out-of-core schedule measured: the bench workload (config 3 by default) with the device budget
forced to a fraction of what the in-core schedules take, pinned host arrays in and out, result compared bit for bit
with the in-core run.

    python tools/stream_bench.py [--shape 256 256 128 128] [--iters 100] [--budgets-gb 20 60]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", type=int, nargs=4, default=[256, 256, 128, 128])
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--budgets-gb", type=float, nargs="+", default=[20.0, 60.0])
    a = ap.parse_args()
    import torch
    import cytvdn_b200 as tv
    from cytvdn_b200 import synth
    shape = tuple(a.shape)
    vox = int(np.prod(shape))
    x = synth.stem4d_device(shape, seed=2, counts=500.0)
    host_in = tv.pinned_empty(shape, np.float32)
    host_out = tv.pinned_empty(shape, np.float32)
    torch.from_numpy(host_in).copy_(x)
    mu = np.array([1, 1, .5, .5], dtype=np.float32)
    ref = tv.denoise4D(x, mu, a.iters, True, quiet=True, schedule="fused")
    ref_sum = int(ref[0].view(torch.int32).to(torch.int64).sum().item())
    ref_bn, ref_dl = ref[1].astype(np.float64), ref[2].astype(np.float64)
    del x, ref
    torch.cuda.empty_cache()
    torch.cuda.synchronize()
    for gb in a.budgets_gb:
        os.environ["CYTVDN_STREAM_BUDGET_MB"] = repr(gb * 1024.0)
        tm = {}
        t0 = time.perf_counter()
        got = tv.denoise4D(host_in, mu, a.iters, True, quiet=True, out=host_out, timing=tm)
        dt = time.perf_counter() - t0
        chk = int(torch.from_numpy(host_out).view(torch.int32).to(torch.int64).sum().item())
        print(json.dumps({
            "shape": list(shape), "iterations": a.iters, "device_budget_GB": gb, "array_GB": vox * 4 / 1e9,
            "in_core_state_GB": {"fused": 19 * vox * 4 / 1e9, "two_pass": 10 * vox * 4 / 1e9}, "schedule": tm["schedule"], "tiles": tm["stream_tiles"],
            "wall_s": dt, "setup_ms": tm["setup_ms"], "loop_ms": tm["loop_ms"],
            "gvoxel_iter_per_s_wall": vox * a.iters / dt / 1e9,
            "gvoxel_iter_per_s_loop": vox * a.iters / (tm["loop_ms"] * 1e-3) / 1e9,
            "recon_bit_identical_to_in_core": chk == ref_sum,
            "bnorm_max_rel": float(np.max(np.abs(got[1].astype(np.float64) - ref_bn) / ref_bn)),
            "delta_max_rel": float(np.max(np.abs(got[2].astype(np.float64) - ref_dl) / ref_dl))}), flush=True)


if __name__ == "__main__":
    main()
