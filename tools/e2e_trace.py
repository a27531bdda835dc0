#!/usr/bin/env python
"""Where does the end-to-end wall time of tv.denoise4D(host arrays) go?  Runs the bench's e2e call a few times with
the PCIe pipeline on and off (CYTVDN_PIPELINE) and CYTVDN_TRACE=1 (host-clock milestones on stderr).

    python tools/e2e_trace.py [--shape 256 256 128 128] [--iters 100] [--repeat 3]
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
    ap.add_argument("--repeat", type=int, default=3)
    ap.add_argument("--modes", nargs="+", default=["16", "0"])
    a = ap.parse_args()
    import torch
    import cytvdn_b200 as tv
    from cytvdn_b200 import synth
    os.environ["CYTVDN_TRACE"] = "1"
    shape = tuple(a.shape)
    vox = int(np.prod(shape))
    x = synth.stem4d_device(shape, seed=2, counts=500.0)
    host_in = tv.pinned_empty(shape, np.float32)
    host_out = tv.pinned_empty(shape, np.float32)
    torch.from_numpy(host_in).copy_(x)
    del x
    torch.cuda.empty_cache()
    torch.cuda.synchronize()
    mu = np.array([1, 1, .5, .5], dtype=np.float32)
    for rep in range(a.repeat):
        for mode in a.modes:
            os.environ["CYTVDN_PIPELINE"] = mode
            tm = {}
            t0 = time.perf_counter()
            tv.denoise4D(host_in, mu, iterations=a.iters, FISTA=True, quiet=True, out=host_out, timing=tm, schedule="fused")
            dt = time.perf_counter() - t0
            print(json.dumps({"rep": rep, "CYTVDN_PIPELINE": mode, "wall_ms": dt * 1e3,
                              "gvoxel_iter_per_s": vox * a.iters / dt / 1e9, **tm}), flush=True)


if __name__ == "__main__":
    main()
