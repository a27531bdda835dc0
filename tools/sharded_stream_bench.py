"""Sharded + out-of-core schedule (cytvdn_denoise_sharded_streamed) on several GPUs from ONE process: a 4D-STEM array
whose state is forced out of core by a per-device budget, host arrays in and out.

    python tools/sharded_stream_bench.py --devices 2 --shape 256 1024 128 128 --iters 100 --budget-gb 60

Prints one JSON line: Gvoxel*iter/s of the loop, tiles / passes of the plan, and -- when --check -- whether the
reconstruction equals the in-core sharded run (cytvdn_denoise_sharded) bit for bit.  The schedule is PCIe bound:
per pass every device moves (arrays in + arrays out) x its share of the array over its x16 link.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import cytvdn_b200 as tv
from cytvdn_b200 import _lib, synth

ap = argparse.ArgumentParser()
ap.add_argument("--devices", type=int, default=2)
ap.add_argument("--shape", type=int, nargs=4, default=[256, 1024, 128, 128])
ap.add_argument("--iters", type=int, default=100)
ap.add_argument("--budget-gb", type=float, default=60.0)
ap.add_argument("--check", action="store_true")
a = ap.parse_args()
shape = tuple(a.shape)
mu = np.array([1, 1, .5, .5], np.float32)
t0 = time.time()
host_in = tv.pinned_empty(shape, np.float32)
host_out = tv.pinned_empty(shape, np.float32)
per = shape[0] // 8 or shape[0]
for i in range(0, shape[0], per):                      # generated on the device plane block by plane block
    blk = synth.stem4d_device(shape, offset0=i, lshape0=min(per, shape[0] - i), seed=2, counts=500.0)
    torch.from_numpy(host_in[i:i + blk.shape[0]]).copy_(blk)
    del blk
torch.cuda.synchronize()
torch.cuda.empty_cache()
t_gen = time.time() - t0
os.environ["CYTVDN_STREAM_BUDGET_MB"] = repr(a.budget_gb * 1024.0)
tm = {}
t0 = time.time()
r = tv.denoise4D(host_in, mu, a.iters, True, quiet=True, out=host_out, devices=list(range(a.devices)), schedule="streamed", timing=tm)
wall = time.time() - t0
del os.environ["CYTVDN_STREAM_BUDGET_MB"]
lib = _lib.load()
P = _lib.DenoiseParams()
P.ndim, P.dtype, P.iters_fista, P.bc_mode = 4, 0, a.iters, 2
for k in range(4):
    P.shape[k] = shape[k]
o = (C.c_int64 * 8)()
lib.cytvdn_stream_plan_sharded(C.byref(P), int(a.budget_gb * 2**30), a.devices, o)
vox = int(np.prod(shape))
line = {"devices": a.devices, "shape": list(shape), "iterations": a.iters, "budget_gb_per_device": a.budget_gb,
        "planes_per_slot": o[0], "iters_per_pass": o[1], "core_planes": o[2], "tiles": o[3], "passes": o[4],
        "host_state_gb": o[7] / 2**30, "setup_s": tm["setup_ms"] / 1e3, "loop_s": tm["loop_ms"] / 1e3, "wall_s": wall,
        "gvox_it_s_loop": vox * a.iters / tm["loop_ms"] / 1e6, "gvox_it_s_wall": vox * a.iters / wall / 1e9,
        "pcie_bytes_per_pass_per_device_gb": (10 + 9) * vox * 4 / a.devices / 2**30, "delta_last": float(r[2][-1]), "gen_s": t_gen}
if a.check:
    ref = np.empty(shape, np.float32)
    rr = tv.denoise4D(host_in, mu, a.iters, True, quiet=True, out=ref, devices=list(range(a.devices)))
    line["equals_in_core_sharded"] = bool(np.array_equal(ref, host_out))
print(json.dumps(line), flush=True)
