#!/usr/bin/env python
"""bench.py -- headline benchmark of the TV-denoising hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A *step* is one TV iteration over the whole array: by default ONE fused kernel (half-step A for all axes +
half-step B, 76 B/voxel); ``--schedule two_pass`` runs the two half-step kernels (96 B/voxel).

N = 1  : BASELINE config 3 -- denoise4D anisotropic FISTA, fp32, 256x256x128x128, mu=[1,1,.5,.5].
N > 1  : BASELINE config 5 shape, weak scaling: every GPU owns 128x1024x128x128 of a
         (128*N)x1024x128x128 array (N = 8 is exactly config 5), axis-0 shards with a one-plane halo
         exchanged per half-step (cytvdn_b200/sharded.py); launched by torchrun, one rank per GPU.

One JSON line on rank 0.  ``value`` = voxels x K / device time of the K timed iterations (inputs
resident in HBM), ``e2e`` = the same metric through the public API ``tv.denoise4D`` with pinned HOST
buffers (H2D of the data and D2H of the result inside the timed region, 100 iterations as the config
says), ``roofline`` = the dominant kernel (the fused iteration kernel, or half-step A with --schedule two_pass)
against the measured HBM copy bandwidth,
``cpu_baseline`` = the unmodified reference kernels (oracle/_ref) timed on this box's host cores on a
bounded sample.  ``--impl reference`` times only that CPU implementation.
"""
from __future__ import annotations

import os

# All host cores for the reference's OpenMP kernels.  Must happen before libgomp is loaded, and must OVERRIDE the
# environment: torchrun exports OMP_NUM_THREADS=1 to every rank, which made round 1's N>1 reference arm single
# threaded.  Only rank 0 ever runs CPU legs.  CYTVDN_BENCH_OMP_THREADS pins another count.
if os.environ.get("RANK", "0") == "0":
    os.environ["OMP_NUM_THREADS"] = os.environ.get("CYTVDN_BENCH_OMP_THREADS", str(os.cpu_count() or 1))

import argparse
import json
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

METRIC = "4D TV-FISTA Gvoxel-iter/s"        # BASELINE.json `metric` (its throughput clause)
UNIT = "Gvoxel*iter/s"
MU = [1.0, 1.0, 0.5, 0.5]
SHAPE_1GPU = (256, 256, 128, 128)            # BASELINE config 3
SHARD_PER_GPU = (128, 1024, 128, 128)        # BASELINE config 5 split over 8 GPUs
BYTES_A, BYTES_B = 68, 28                    # algorithmic bytes / voxel, fp32 4-D FISTA
BYTES_FUSED = 76                             # fused single pass: every array crosses HBM once
CPU_SAMPLE_SHAPE = (32, 32, 128, 128)        # bounded sample of the same workload for the CPU legs


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def committed_traffic():
    """Per-launch DRAM bytes of the dominant kernel from the committed ncu --set full capture."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        return {}


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "power_w_max": float(max(pw)), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU legs (the only place bench.py executes oracle/)
# ------------------------------------------------------------------------------------------------
def mem_available_gb():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) / 1048576.0
    except Exception:
        pass
    return 0.0


def cpu_reference_run(steps, warmup, budget_s=25.0, full_size=False):
    """4-D FISTA iterations of the reference's CPU implementation (its unmodified kernels from oracle/_ref, driven
    as cyTVDN.py:154-184 drives them; the C port when oracle/_ref is absent).  ``full_size``: the config-3 array
    itself when the host has the memory for the reference's 10 arrays (43 GB), else -- and for the bounded
    cpu_baseline leg of the GPU arm -- CPU_SAMPLE_SHAPE, same generator, reduced scan size.
    Returns dict(value, ms_per_step, steps, warmup, kind, cores, sample, shape)."""
    from oracle import tv_oracle as O
    from cytvdn_b200 import synth
    kind = "reference" if O.reference_available() else "port"
    K = O.ReferenceKernels("T") if kind == "reference" else O.PortKernels("T")
    cores = int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1))
    if kind == "port":
        O.set_threads(cores)
    data = synth.stem4d_hash_numpy(CPU_SAMPLE_SHAPE, seed=2, counts=500.0)
    shape, how = CPU_SAMPLE_SHAPE, "same generator, reduced scan size"
    if full_size and mem_available_gb() >= 60.0 and os.environ.get("CYTVDN_BENCH_CPU_FULL", "1") != "0":
        # the full 256x256x128x128 array: the sample block repeated over the scan axes (generating 1.07 G voxels
        # with the NumPy mirror of the device generator would take minutes; per-voxel cost does not depend on it)
        reps = (SHAPE_1GPU[0] // shape[0], SHAPE_1GPU[1] // shape[1], 1, 1)
        data = np.tile(data, reps)
        shape, how = SHAPE_1GPU, "full config-3 size, the 32x32 scan sample tiled 8x8"
    mu = np.array(MU, dtype=np.float32)
    lam = mu / 32.0
    clip, w = 1.0 / lam, (lam / mu).astype(np.float32)
    acc = [np.zeros_like(data) for _ in range(4)]
    dd = [np.zeros_like(data) for _ in range(4)]
    recon = data.copy()
    tk = 1.0
    vox = data.size

    def one():
        nonlocal tk
        r, tk_new = O.fista_ratio(tk)
        tk = tk_new
        for ax in range(4):
            K.accumulator_update(recon, acc[ax], dd[ax], r, ax, clip[ax], 2)
        K.datacube_update(data, recon, acc, w, 2)

    # warm-up (the first iteration also first-touches the arrays and calibrates the budget)
    t0 = time.perf_counter()
    one()
    t_one = time.perf_counter() - t0
    w_eff = max(1, min(warmup, int(0.25 * budget_s / max(t_one, 1e-6))))
    for _ in range(w_eff - 1):
        one()
    k_eff = max(1, min(steps, int(0.75 * budget_s / max(t_one, 1e-6))))
    t0 = time.perf_counter()
    for _ in range(k_eff):
        one()
    dt = time.perf_counter() - t0
    return dict(value=vox * k_eff / dt / 1e9, ms_per_step=1e3 * dt / k_eff, steps=k_eff, warmup=w_eff,
                kind=kind, cores=cores, shape=list(shape),
                sample=f"4-D FISTA fp32 {'x'.join(map(str, shape))} ({how}), "
                       f"{k_eff} timed iterations after {w_eff} warm-up, OMP_NUM_THREADS={cores}")


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    r = cpu_reference_run(args.steps, args.warmup, budget_s=150.0, full_size=True)
    cfg = workload_config(args.gpus)
    cfg["cpu_sample_shape"] = r["shape"]
    same = list(r["shape"]) == list(cfg.get("shape", []))
    cfg["cpu_sample"] = ("each step is one FISTA iteration of the reference's CPU kernels over the configured array itself"
                         if same else
                         "the reference's CPU path cannot hold this array (N > 1: 10 arrays of the sharded cube; N = 1 on a "
                         "host with < 60 GB free: 43 GB); each step is one FISTA iteration over " + "x".join(map(str, r["shape"]))
                         + " voxels of the same workload, reported per voxel")
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": r["steps"], "warmup": r["warmup"], "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                         "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(n):
    if n == 1:
        return {"workload": "denoise4D anisotropic FISTA fp32 256x256x128x128 mu=[1,1,.5,.5] (BASELINE config 3)",
                "shape": list(SHAPE_1GPU), "l2": "working set 42.9 GB >> 126 MB L2 (inputs larger than L2)"}
    g = (SHARD_PER_GPU[0] * n,) + SHARD_PER_GPU[1:]
    return {"workload": f"denoise4D anisotropic FISTA fp32 {'x'.join(map(str, g))} sharded on scan axis 0 over {n} GPUs "
                        f"(BASELINE config 5 shape, {'x'.join(map(str, SHARD_PER_GPU))} per GPU; N=8 is config 5)",
            "shape": list(g), "shard": list(SHARD_PER_GPU), "halo": "fused: 1 plane of the new recon left and right per iteration; two_pass: 1 plane of b0 right / recon left per half-step (NCCL send/recv, inline after the halo planes)",
            "l2": "working set 86 GB per GPU >> 126 MB L2"}


# ------------------------------------------------------------------------------------------------
# GPU arm, N = 1
# ------------------------------------------------------------------------------------------------
def run_single(args):
    import ctypes as C
    import torch
    import cytvdn_b200 as tv
    from cytvdn_b200 import _lib, synth

    lib = _lib.load()
    _lib.require_gpu()
    torch.cuda.set_device(0)
    shape = tuple(args.shape) if args.shape else SHAPE_1GPU
    vox = int(np.prod(shape))
    fused = args.schedule != "two_pass"              # the other names only differ in how shards talk to each other
    x = synth.stem4d_device(shape, seed=2, counts=500.0)
    nset = 2 if fused else 1
    # --skew: offset the k-th state array by k * SKEW bytes (experiment: power-of-two array sizes showed no
    # set-aliasing penalty on B200, default 0)
    skew_elems = args.skew // 4
    counter = [0]

    def alloc(zero):
        counter[0] += 1
        off = counter[0] * skew_elems
        base = (torch.zeros if zero else torch.empty)(vox + off, dtype=torch.float32, device="cuda")
        return base[off:off + vox].view(shape)

    B = [[alloc(True) for _ in range(4)] for _ in range(nset)]
    D = [[alloc(True) for _ in range(4)] for _ in range(nset)]
    R = [alloc(False) for _ in range(nset)]
    sums = torch.zeros(4 * (args.steps + args.warmup + 1), dtype=torch.float64, device="cuda")
    sh = (C.c_int64 * 4)(*shape)
    mu = np.array(MU, dtype=np.float32)
    lam = mu / np.float32(32.0)
    clip = (C.c_double * 4)(*[float(v) for v in (1.0 / lam)])
    w = (C.c_double * 4)(*[float(v) for v in (lam / mu).astype(np.float32)])
    ptrs = lambda ts: (C.c_void_p * 4)(*[t.data_ptr() for t in ts])
    BP, DP = [ptrs(b) for b in B], [ptrs(d) for d in D]
    st = torch.cuda.current_stream().cuda_stream
    state = {"tk": 1.0, "it": 0, "cur": 0, "u": x}          # iteration 0 reads recon == data (cyTVDN.py:145)

    def step(ev=None):
        tk = state["tk"]
        tk_new = (1.0 + np.sqrt(1.0 + 4.0 * tk * tk)) / 2.0
        r = (tk - 1.0) / tk_new
        state["tk"] = tk_new
        s = sums.data_ptr() + 32 * state["it"]
        state["it"] += 1
        cur = state["cur"]
        u_in = state["u"]
        if ev:
            ev[0].record()
        if fused:
            u_out = R[cur]
            _lib.check(lib.cytvdn_fused_iteration(4, sh, 0, x.data_ptr(), u_in.data_ptr(), u_out.data_ptr(), BP[cur],
                                                  BP[1 - cur], DP[cur], DP[1 - cur], r, clip, w, 2, s, None, st))
            state["cur"] = 1 - cur
            if ev:
                ev[1].record()
        else:
            u_out = R[0]
            _lib.check(lib.cytvdn_accumulator_update_all(4, sh, 0, u_in.data_ptr(), BP[0], DP[0], r, clip, 0, 0, 2, s,
                                                         None, st))
            if ev:
                ev[1].record()
            _lib.check(lib.cytvdn_datacube_update(4, sh, 0, x.data_ptr(), u_in.data_ptr(), u_out.data_ptr(), BP[0], w, 2,
                                                  s + 8, None, st))
        state["u"] = u_out
        if ev:
            ev[2].record()

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    clocks = ClockSampler(0)
    clocks.start()
    time.sleep(0.3)
    l0 = tv.launch_count()
    torch.cuda.synchronize()
    for k in range(args.steps):
        step(evs[k])
    torch.cuda.synchronize()
    launches = tv.launch_count() - l0
    clk = clocks.stop()
    total_ms = evs[0][0].elapsed_time(evs[-1][2])
    a_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in evs]))
    b_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in evs]))
    ms_per_step = total_ms / args.steps
    value = vox * args.steps / (total_ms * 1e-3) / 1e9
    s_host = sums.cpu().numpy().reshape(-1, 4)
    last = s_host[state["it"] - 1]
    peak, peak_src = measured_peak()
    traffic = committed_traffic() if shape == SHAPE_1GPU else {}        # the ncu capture is of config 3
    eq96 = (BYTES_A + BYTES_B) * vox / (ms_per_step * 1e-3) / 1e9
    if fused:
        # a fused single pass is still reported against the 96 B/voxel contract figure of
        # the two-pass iteration, with the variant stated; what the kernel really moves is 76 B/voxel.
        ach = (BYTES_A + BYTES_B) * vox / (a_ms * 1e-3) / 1e9
        moved = BYTES_FUSED * vox / (a_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm",
                    "kernel": "tv_fused_kernel<float,4,FISTA,4D> (whole iteration = half-steps A+B in one pass)",
                    "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic.get("tv_fused_kernel"), "traffic_source": traffic.get("_source"),
                    "peak_source": peak_src, "ms_per_launch": a_ms,
                    "bytes_per_voxel": BYTES_A + BYTES_B,
                    "variant": "fused single pass: the contract's 96 B/voxel of work (SURVEY 8d) is done while moving "
                               "76 B/voxel (every array crosses HBM once), hence frac > 1; see `moved`",
                    "moved": {"bytes_per_voxel": BYTES_FUSED, "GB/s": moved, "frac": moved / peak,
                              "frac_of_8TBs_nominal": moved / 8000.0},
                    "frac_of_8TBs_nominal": ach / 8000.0}
    else:
        ach_a = BYTES_A * vox / (a_ms * 1e-3) / 1e9
        ach_b = BYTES_B * vox / (b_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "tv_accumulator_kernel<float,4,FISTA,ALL4> (half-step A, 68 B/voxel)",
                    "achieved": ach_a, "peak": peak, "unit": "GB/s", "frac": ach_a / peak,
                    "traffic": traffic.get("tv_accumulator_kernel"), "peak_source": peak_src, "ms_per_launch": a_ms,
                    "other_kernels": [{"kernel": "tv_datacube_kernel<float,4,true> (half-step B, 28 B/voxel)",
                                       "achieved": ach_b, "frac": ach_b / peak, "ms_per_launch": b_ms,
                                       "traffic": traffic.get("tv_datacube_kernel")}],
                    "iteration": {"bytes_per_voxel": BYTES_A + BYTES_B, "achieved": eq96, "frac": eq96 / peak,
                                  "frac_of_8TBs_nominal": eq96 / 8000.0}}
    del B, D, R, BP, DP
    state.clear()
    torch.cuda.empty_cache()

    # ---- end to end through the public API with pinned host buffers ------------------------------
    # tv.denoise4D(host array in, host array out): H2D of the data and D2H of the result inside every timed call.
    # The device working set is reserved once (tv.workspace_reserve, the C ABI's cytvdn_workspace_reserve), as a
    # caller that denoises more than one cube would do, so cudaMalloc/cudaFree of ~86 GB (60 ms .. 0.7 s, the
    # unexplained host time of round 1) are not part of a call; >= 3 timed calls, all samples reported, value = median.
    e2e = None
    if not args.no_e2e:
        iters = args.e2e_iters
        sched = "fused" if fused else "two_pass"
        host_in = tv.pinned_empty(shape, np.float32)
        host_out = tv.pinned_empty(shape, np.float32)
        torch.from_numpy(host_in).copy_(x)
        del x
        torch.cuda.empty_cache()
        torch.cuda.synchronize()
        reserved = tv.workspace_reserve((shape, np.float32), iterations=iters, FISTA=True, host_arrays=True, schedule=sched)
        # one short untimed call first (same arrays, 3 iterations): first-use costs of the copy path
        tv.denoise4D(host_in, mu, iterations=3, FISTA=True, quiet=True, out=host_out, schedule=sched)
        samples, tms = [], []
        for _ in range(max(1, args.e2e_calls)):
            tm = {}
            t0 = time.perf_counter()
            tv.denoise4D(host_in, mu, iterations=iters, FISTA=True, quiet=True, out=host_out, timing=tm, schedule=sched)
            samples.append(time.perf_counter() - t0)
            tms.append(tm)
        tv.workspace_release()
        order = np.argsort(samples)
        mid = int(order[len(order) // 2])
        dt, tm = samples[mid], tms[mid]
        nbytes = vox * 4
        e2e = {"value": vox * iters / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": nbytes / iters,
               "d2h_bytes_per_step": nbytes / iters,
               "call": f"tv.denoise4D(pinned host fp32, iterations={iters}, FISTA=True, schedule={tm.get('schedule')}), "
                       f"device working set reserved once with tv.workspace_reserve ({reserved / 2**30:.1f} GiB)",
               "wall_s": dt, "wall_s_samples": samples,
               "value_samples": [vox * iters / t / 1e9 for t in samples], "timed_calls": len(samples),
               "spread": (max(samples) - min(samples)) / dt,
               "warmup_calls": 1, "pcie_pipeline_boxes": tm.get("pipeline_boxes"), "setup_ms": tm.get("setup_ms"),
               "loop_ms": tm.get("loop_ms"), "finish_ms": tm.get("finish_ms"), "trace_ms": tm.get("trace_ms"),
               "h2d_bytes_total": nbytes, "d2h_bytes_total": nbytes}
        del host_in, host_out
    else:
        del x
    torch.cuda.empty_cache()
    configs = None
    if args.configs and not args.shape:
        try:
            from bench_configs import run_configs
            configs = run_configs(peak)
        except Exception as ex:           # a side key must never take the headline line down
            configs = {"error": repr(ex)[:300]}
    cpu = None
    if not args.no_cpu:
        r = cpu_reference_run(5, 1, budget_s=20.0)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}

    cfg = workload_config(1) if not args.shape else \
        {"workload": f"denoise4D anisotropic FISTA fp32 {'x'.join(map(str, shape))} (non-default shape)"}
    cfg["schedule"] = args.schedule
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": cfg,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk,
            "configs": configs,
            "check": {"bnorm_last": float(last[0]), "delta_last": float(last[1] / last[2]) if last[2] else None}}
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--shape", type=int, nargs=4, default=None, help="override the N=1 shape (debugging)")
    ap.add_argument("--e2e-iters", type=int, default=100)
    ap.add_argument("--e2e-calls", type=int, default=3, help="timed end-to-end calls (median reported)")
    ap.add_argument("--no-configs", dest="configs", action="store_false",
                    help="skip the `configs` extra key (C1, C2, C4, unaccelerated, fp64 through the public API)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--skew", type=int, default=0, help="byte skew between consecutive state arrays (experiment knob)")
    ap.add_argument("--schedule", default="fused", choices=["fused", "two_pass", "peer", "engine", "nccl_fused"],
                    help="fused: one pass per iteration (76 B/voxel; N > 1: the C-ABI shard engine, copy-engine halo "
                         "exchange); two_pass: half-steps A and B (96 B/voxel; N > 1: NCCL exchange); nccl_fused / peer: "
                         "round 1's torch.distributed schedules (N > 1)")
    ap.add_argument("--no-check", action="store_true", help="N > 1: skip the parity checks over the process group")
    ap.add_argument("--no-alone", action="store_true", help="N > 1: skip the single-GPU run of one shard's shape")
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the strong-scaling point of config 5 itself")
    ap.add_argument("--extras-timeout", type=float, default=120.0,
                    help="N > 1: seconds the side measurements (e2e, single shard, strong scaling, CPU baseline) may take "
                         "before the line is emitted without the unfinished ones")
    ap.add_argument("--timeline", default=None, help="N > 1: write per-phase CUDA-event timings of the timed steps here")
    ap.add_argument("--timeline-full", action="store_true", help="keep every iteration of every rank in the timeline file")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 or world > 1:
        from bench_sharded import run_sharded
        return run_sharded(args)
    return run_single(args)


if __name__ == "__main__":
    sys.exit(main())
