"""Loop throughput of the fused kernel's variants through the public API (device-resident tensors):

    python tools/plain_bench.py [c1|plain4d|fp64|c3] [--iters N]
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
ap.add_argument("which", nargs="?", default="all")
ap.add_argument("--iters", type=int, default=40)
a = ap.parse_args()
mu4 = np.array([1, 1, .5, .5], np.float32)
cases = {
    "c1": lambda: (tv.denoise3D, torch.from_numpy(synth.eels_cube((128, 128, 1024), seed=0)).cuda(), np.array([1, 1, .5], np.float32), dict(FISTA=False)),
    "plain4d": lambda: (tv.denoise4D, synth.stem4d_device((256, 256, 128, 128), seed=2, counts=500.0), mu4, dict(FISTA=False)),
    "c3": lambda: (tv.denoise4D, synth.stem4d_device((256, 256, 128, 128), seed=2, counts=500.0), mu4, dict(FISTA=True)),
    "fp64": lambda: (tv.denoise4D, synth.stem4d_device((128, 256, 128, 128), seed=2, counts=500.0, dtype="float64"), mu4.astype(np.float64), dict(FISTA=True)),
}
for name, make in cases.items():
    if a.which not in ("all", name):
        continue
    fn, x, mu, kw = make()
    fn(x, mu, iterations=3, quiet=True, schedule="fused", **kw)
    tm = {}
    fn(x, mu, iterations=a.iters, quiet=True, schedule="fused", timing=tm, **kw)
    print(json.dumps({"case": name, "shape": list(x.shape), "ms_per_it": round(tm["loop_ms"] / a.iters, 4),
                      "gvox_it_s": round(x.numel() * a.iters / tm["loop_ms"] / 1e6, 2)}), flush=True)
    del x
    torch.cuda.empty_cache()
