#!/usr/bin/env python
"""Run every BASELINE.json config that fits one GPU through the PUBLIC API (tv.denoise3D / tv.denoise4D on
device-resident tensors) and print one JSON object per config: loop time from the library's CUDA events,
Gvoxel*iter/s, fraction of the HBM roofline for that variant's bytes/voxel, schedule used, and -- for the
configs the CPU reference can run in seconds -- parity against the unmodified reference (oracle/_ref).

    python tools/run_configs.py [--out profiles/configs_r1.json]

Not a bench line (bench.py is); this fills the per-config table of DESIGN.md / profiles/.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("OMP_NUM_THREADS", str(os.cpu_count() or 1))

import numpy as np


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--skip-cpu", action="store_true")
    a = ap.parse_args()
    import torch
    import cytvdn_b200 as tv
    from cytvdn_b200 import synth
    from oracle import tv_oracle as O

    pk = peak()
    res = []

    def bytes_per_voxel(ndim, fista, fused, elem=4):
        if fused:
            return (3 + ndim * (4 if fista else 2)) * elem
        return (1 + ndim * (4 if fista else 2) + 2 + ndim + 1) * elem

    def run(name, fn, data, mu, nvox, ndim, fista, iters, **kw):
        out = {"config": name, "shape": list(data.shape), "dtype": str(data.dtype).replace("torch.", "")}
        for sched in (("fused", "two_pass") if not (kw.get("isotropic_R") or kw.get("isotropic_Q")) else ("two_pass",)):
            tm = {}
            fn(data, mu, quiet=True, timing=tm, schedule=sched, **dict(kw, iterations=3))      # warm-up
            tm = {}
            r = fn(data, mu, quiet=True, timing=tm, schedule=sched, **dict(kw, iterations=iters))
            done = tm["iters_fista"] + tm["iters_plain"]
            gv = nvox * done / (tm["loop_ms"] * 1e-3) / 1e9
            bpv = bytes_per_voxel(ndim, fista, sched == "fused", 4 if data.dtype == torch.float32 else 8)
            out[sched] = {"iterations_run": done, "loop_ms": tm["loop_ms"], "ms_per_iter": tm["loop_ms"] / max(done, 1),
                          "gvoxel_iter_per_s": gv, "bytes_per_voxel": bpv, "achieved_GBps": gv * bpv,
                          "frac_of_measured_peak": gv * bpv / pk,
                          "delta_last": float(r[2][done - 1]) if done else None}
            # size-independent check: both schedules give the SAME array (compared through an order-independent
            # 64-bit checksum of the bit patterns) and the same scalars
            bits = r[0].contiguous().view(torch.int32 if r[0].dtype == torch.float32 else torch.int64)
            out[sched]["recon_checksum"] = int(bits.to(torch.int64).sum().item()) & 0xFFFFFFFFFFFFFFFF
            out[sched]["bnorm_last"] = float(r[1][done - 1]) if done else None
            del r, bits
            torch.cuda.empty_cache()
        if "fused" in out and "two_pass" in out:
            out["schedules_bit_identical"] = out["fused"]["recon_checksum"] == out["two_pass"]["recon_checksum"]
        print(json.dumps(out), flush=True)
        res.append(out)
        return out

    f32 = np.float32
    # ---- config 1: denoise3D anisotropic plain, 128x128x1024, mu=[1,1,.5], 100 iterations (+ CPU parity) ----
    cube = synth.eels_cube((128, 128, 1024), seed=0, dose=1000.0, gain=1.0)
    mu3 = np.array([1, 1, .5], dtype=f32)
    t = torch.from_numpy(cube).cuda()
    o1 = run("C1 denoise3D plain 128x128x1024", tv.denoise3D, t, mu3, cube.size, 3, False, 100, FISTA=False)
    if not a.skip_cpu:
        t0 = time.perf_counter()
        ref = O.denoise3D(cube, mu3, 100, FISTA=False, quiet=True, kernels=O.default_kernels("T"), scalars="T")
        cpu_s = time.perf_counter() - t0
        got = tv.denoise3D(cube, mu3, 100, FISTA=False, quiet=True)
        truth = O.denoise3D(cube, mu3, 100, FISTA=False, quiet=True, kernels=O.PortKernels("D"), scalars="D")
        o1["parity_vs_reference"] = {
            "recon_max_abs_diff": float(np.abs(got[0] - ref[0]).max()), "data_range": float(cube.max() - cube.min()),
            "bnorm_max_rel_vs_f64_truth": float(np.max(np.abs(got[1].astype(np.float64) - truth[1]) / truth[1])),
            "delta_max_rel_vs_f64_truth": float(np.max(np.abs(got[2].astype(np.float64) - truth[2]) / truth[2])),
            "reference_own_fp32_scalars_rel_err_vs_truth": {
                "bnorm": float(np.max(np.abs(ref[1].astype(np.float64) - truth[1]) / truth[1])),
                "delta": float(np.max(np.abs(ref[2].astype(np.float64) - truth[2]) / truth[2]))},
            "cpu_reference_s": cpu_s, "cpu_cores": int(os.environ["OMP_NUM_THREADS"]),
            "cpu_gvoxel_iter_per_s": cube.size * 100 / cpu_s / 1e9}
        print(json.dumps({"C1_parity": o1["parity_vs_reference"]}), flush=True)
    del t
    # ---- config 2: denoise3D FISTA 512x512x2048, stopping_relative_change = 0.05 and fixed 100 -------------
    big = torch.from_numpy(synth.eels_cube((64, 512, 2048), seed=1, dose=2.0, gain=8.0)).cuda()
    big = big.repeat(8, 1, 1).contiguous()                 # 512x512x2048 (host generation of the full cube is slow)
    big += torch.from_numpy(np.random.default_rng(1).poisson(2.0, (512, 1, 1)).astype(f32)).cuda() * 8.0
    run("C2 denoise3D FISTA 512x512x2048 (fixed 100 it.)", tv.denoise3D, big, mu3, big.numel(), 3, True, 100, FISTA=True)
    tm = {}
    r = tv.denoise3D(big, mu3, 100, 0.05, 2, True, quiet=True, timing=tm)
    o = {"config": "C2 denoise3D FISTA 512x512x2048, stopping_relative_change=0.05", "iterations_run": tm["iters_fista"],
         "loop_ms": tm["loop_ms"], "delta": [float(v) for v in r[2][:tm["iters_fista"]]], "schedule": tm["schedule"]}
    print(json.dumps(o), flush=True)
    res.append(o)
    del big, r
    torch.cuda.empty_cache()
    # ---- config 3 / 4: denoise4D FISTA 256x256x128x128, anisotropic and half-isotropic ----------------------
    mu4 = np.array([1, 1, .5, .5], dtype=f32)
    x = synth.stem4d_device((256, 256, 128, 128), seed=2, counts=500.0)
    run("C3 denoise4D anisotropic FISTA 256x256x128x128", tv.denoise4D, x, mu4, x.numel(), 4, True, 100, FISTA=True)
    run("C4 denoise4D half-isotropic FISTA 256x256x128x128", tv.denoise4D, x, mu4, x.numel(), 4, True, 100, FISTA=True,
        isotropic_R=True, isotropic_Q=True)
    run("4-D anisotropic plain 256x256x128x128", tv.denoise4D, x, mu4, x.numel(), 4, False, 100, FISTA=False)
    del x
    torch.cuda.empty_cache()
    # ---- fp64 ----------------------------------------------------------------------------------------------
    x64 = synth.stem4d_device((128, 256, 128, 128), seed=2, counts=500.0, dtype="float64")
    run("fp64 denoise4D anisotropic FISTA 128x256x128x128", tv.denoise4D, x64, mu4.astype(np.float64), x64.numel(), 4,
        True, 50, FISTA=True)
    if a.out:
        json.dump({"peak_GBps": pk, "results": res}, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
