"""Throughput of tv.denoise4D / denoise3D over a few awkward shapes (odd inner extents -> scalar path, short rows,
fp64).  Prints Gvoxel*iter/s and the fraction of the HBM roofline for the schedule that ran."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cytvdn_b200 as tv

pk = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
cases = [((256, 256, 128, 128), "float32"), ((256, 256, 126, 127), "float32"), ((256, 256, 127, 126), "float32"),
         ((128, 128, 250, 250), "float32"), ((64, 64, 512, 512), "float32"), ((300, 300, 96, 100), "float32"),
         ((512, 512, 32, 32), "float32"), ((2048, 2048, 16, 16), "float32"), ((128, 256, 128, 128), "float64"),
         ((128, 256, 127, 127), "float64"), ((512, 512, 2048), "float32"), ((300, 301, 1999), "float32")]
for shape, dt in cases:
    n = int(np.prod(shape))
    x = (torch.rand(shape, device="cuda", dtype=torch.float32) * 500).round().to(getattr(torch, dt))
    nd = len(shape)
    mu = np.array([1, 1, .5, .5][:nd] if nd == 4 else [1, 1, .5], dtype=dt)
    fn = tv.denoise4D if nd == 4 else tv.denoise3D
    for sched in ("fused", "two_pass"):
        fn(x, mu, iterations=3, FISTA=True, quiet=True, schedule=sched)
        tm = {}
        fn(x, mu, iterations=20, FISTA=True, quiet=True, schedule=sched, timing=tm)
        gv = n * 20 / (tm["loop_ms"] * 1e-3) / 1e9
        el = 4 if dt == "float32" else 8
        bpv = ((3 + nd * 4) if sched == "fused" else (1 + nd * 4 + 2 + nd + 1)) * el
        print(f"{str(shape):24s} {dt:8s} {sched:9s} {gv:7.2f} Gvox*it/s  {gv*bpv:7.0f} GB/s  frac {gv*bpv/pk:5.3f}", flush=True)
    del x
    torch.cuda.empty_cache()
