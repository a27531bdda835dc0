"""Throughput of the half-isotropic path (BASELINE config 4) through the public API, both schedules.

    python tools/iso_bench.py [--shape 256 256 128 128] [--iters 40] [--only fused|two_pass]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import cytvdn_b200 as tv
from cytvdn_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--shape", type=int, nargs=4, default=[256, 256, 128, 128])
ap.add_argument("--iters", type=int, default=40)
ap.add_argument("--only", default=None)
a = ap.parse_args()
x = synth.stem4d_device(tuple(a.shape), seed=2, counts=500.0)
mu = np.array([1, 1, .5, .5], np.float32)
for kw in (dict(isotropic_R=True, isotropic_Q=True), dict(isotropic_R=True), dict(isotropic_Q=True), dict()):
    for sched in ("two_pass", "fused"):
        if a.only and sched != a.only:
            continue
        tv.denoise4D(x, mu, 3, True, quiet=True, schedule=sched, **kw)
        tm = {}
        r = tv.denoise4D(x, mu, a.iters, True, quiet=True, schedule=sched, timing=tm, **kw)
        print(json.dumps({"flags": kw, "schedule": tm["schedule"], "gvox_it_s": round(x.numel() * a.iters / tm["loop_ms"] / 1e6, 2),
                          "ms_per_it": round(tm["loop_ms"] / a.iters, 3), "delta_last": float(r[2][-1])}), flush=True)
        del r
        torch.cuda.empty_cache()
