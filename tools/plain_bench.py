"""Quick throughput check of the unaccelerated (plain) fused variants."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cytvdn_b200 as tv
for shape in ((256, 256, 128, 128), (128, 128, 1024), (512, 512, 2048)):
    nd = len(shape)
    x = (torch.rand(shape, device="cuda") * 500).round()
    mu = np.array([1, 1, .5, .5][:nd] if nd == 4 else [1, 1, .5], np.float32)
    fn = tv.denoise4D if nd == 4 else tv.denoise3D
    for fista in (False, True):
        fn(x, mu, iterations=3, FISTA=fista, quiet=True, schedule="fused")
        tm = {}
        fn(x, mu, iterations=40, FISTA=fista, quiet=True, schedule="fused", timing=tm)
        print(shape, "FISTA" if fista else "plain", round(x.numel() * 40 / tm["loop_ms"] / 1e6, 2), "Gvox*it/s", flush=True)
    del x
