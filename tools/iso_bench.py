"""Throughput of the half-isotropic path (BASELINE config 4) through the public API."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cytvdn_b200 as tv
from cytvdn_b200 import synth
x = synth.stem4d_device((256, 256, 128, 128), seed=2, counts=500.0)
mu = np.array([1, 1, .5, .5], np.float32)
for kw in (dict(isotropic_R=True, isotropic_Q=True), dict(isotropic_R=True), dict()):
    tv.denoise4D(x, mu, 3, True, quiet=True, schedule="two_pass", **kw)
    tm = {}
    tv.denoise4D(x, mu, 40, True, quiet=True, schedule="two_pass", timing=tm, **kw)
    print(kw, round(x.numel() * 40 / tm["loop_ms"] / 1e6, 2), "Gvox*it/s", round(tm["loop_ms"] / 40, 3), "ms/it", flush=True)
