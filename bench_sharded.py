"""bench.py's N > 1 arm: BASELINE config 5 (4-D anisotropic FISTA, fp32, scan-axis shards with a one-plane halo),
weak scaling with 128x1024x128x128 owned per GPU (N = 8 is config 5 itself).  Launched by torchrun, one rank per GPU.

Default schedule: the C-ABI shard engine (csrc/cytvdn_shard.cu) -- fused single-pass sweeps, halo planes pushed by the
copy engines through CUDA-IPC peer pointers under the interior sweep, no NCCL in the loop.  ``--schedule two_pass`` /
``nccl_fused`` / ``peer`` run the round-1 torch.distributed schedules of cytvdn_b200/sharded.py.

Before anything is timed every run checks, over the REAL process group and with the arrays of THIS launch:
  * reduced size (61x24x32x32, uneven split): engine == NCCL fused == NCCL two-pass == single GPU, Jia-Zhao and
    periodic, bit for bit (``check.nccl_parity``);
  * full size: the first iterations of the engine against the NCCL two-pass schedule through an order-independent
    checksum of the recon bit patterns (``check.fullsize_engine_vs_two_pass``).
A mismatch makes the bench exit non-zero without a line.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np


def _checksum(t):
    """Order-independent 64-bit checksum of the bit patterns of a CUDA float32 tensor (plane chunks: no big temporaries)."""
    import torch
    tot = 0
    for i in range(0, t.shape[0], 8):
        tot += int(t[i:i + 8].contiguous().view(torch.int32).to(torch.int64).sum().item())
    return tot & 0xFFFFFFFFFFFFFFFF


def make_engine(sharded, gshape, world, rank, mu, max_iters, local):
    import torch.distributed as dist
    eng = sharded.EngineShard(gshape, world, rank, mu, None, np.float32, fista=True, max_iters=max_iters, device=local)
    handles = [None] * world
    dist.all_gather_object(handles, eng.export())
    eng.connect_all(handles)
    eng.load_synth(seed=2, counts=500.0)                      # this rank's planes straight into the arena
    dist.barrier()
    return eng


def close_engine(eng):
    eng.close_collective()                                    # synchronise, barrier, unmap neighbours, barrier, free


def all_ok(flag, dev):
    """True iff `flag` is true on EVERY rank (a rank-local failure must never leave the others inside a collective)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([int(bool(flag))], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(t.item())


def parity_check(dev, rank, world, mu):
    """Reduced-size sharded-vs-single-GPU check over the real process group.  Returns (ok, details)."""
    import torch
    import torch.distributed as dist
    import cytvdn_b200 as tv
    from cytvdn_b200 import sharded, synth
    g = (61, 24, 32, 32)                                      # 61 planes: uneven for every world size here
    data = synth.stem4d_device(g, seed=7, counts=300.0, device=dev)
    details, ok = [], True
    for periodic in (False, True):
        ref = tv.denoise4D(data, mu, [6, 3], True, BC_mode=0 if periodic else 2, quiet=True)[0]
        plan = sharded.ShardPlan(g, world, rank, None, periodic)
        block = plan.extract(data).contiguous()
        want = ref[plan.owned_global[0]]
        own = sharded.denoise4D_engine(block, mu, [6, 3], True, gshape=g, periodic=periodic)[0]
        res = {"engine": bool(torch.equal(own, want))}
        for sched in ("fused", "two_pass"):
            got = sharded.denoise4D_sharded(block, mu, [6, 3], True, plan=plan, schedule=sched, engine=False)[0]
            res["nccl_" + sched] = bool(torch.equal(got[plan.owned_local[0]], want))
        flag = torch.tensor([int(all(res.values()))], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        details.append({"periodic": periodic, **res, "all_ranks": bool(flag.item())})
        ok = ok and bool(flag.item())
    return ok, details


def run_sharded(args):
    import torch
    import torch.distributed as dist
    import cytvdn_b200 as tv
    from cytvdn_b200 import sharded, synth
    from bench import (BYTES_A, BYTES_B, METRIC, MU, SHARD_PER_GPU, UNIT, ClockSampler, cpu_reference_run, measured_peak,
                       workload_config)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # a rank that dies must not keep the others waiting for NCCL's default 10 minutes
    import datetime
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev, timeout=datetime.timedelta(seconds=150))
    rc = 0
    try:
        per = tuple(args.shape) if args.shape else SHARD_PER_GPU
        gshape = (per[0] * world,) + tuple(per[1:])
        mu = np.array(MU, dtype=np.float32)
        n_total = args.warmup + args.steps
        schedule = {"fused": "engine"}.get(args.schedule, args.schedule)       # default: the C-ABI engine
        use_engine = schedule == "engine"
        fused = schedule in ("engine", "nccl_fused", "peer")
        peer = schedule == "peer"

        # ---- parity over the real process group, before anything is timed --------------------------------
        check = {}
        if not args.no_check:
            ok, details = parity_check(dev, rank, world, mu)
            check["nccl_parity"] = "bit-exact" if ok else "MISMATCH"
            check["nccl_parity_cases"] = details
            if not ok:
                if rank == 0:
                    print(json.dumps({"error": "sharded result differs from the single-GPU result", "check": check}),
                          file=sys.stderr, flush=True)
                return 3
        torch.cuda.empty_cache()

        plan = sharded.ShardPlan(gshape, world, rank)
        n0 = plan.local_shape[0]
        eng = sh = None
        max_it = max(n_total, args.e2e_iters, 8)
        # ---- full size: the engine's first iterations against the NCCL two-pass schedule (checksums).  The two
        #      state sets do not fit side by side (19 + 10 arrays), so the engine is torn down in between. ----
        if use_engine and not args.no_check:
            eng = make_engine(sharded, gshape, world, rank, mu, max_it, local)
            eng.iterate(4, 0)
            eng.synchronize()
            mine = _checksum(eng.array("recon")[eng.own_lo:eng.own_hi])
            close_engine(eng)
            eng = None
            torch.cuda.empty_cache()
            need = 10 * int(np.prod(plan.local_shape)) * 4 + (3 << 30)           # x + 9 state arrays + slack
            if all_ok(torch.cuda.mem_get_info(dev)[0] > need, dev):
                x = synth.stem4d_device(gshape, offset0=plan.read[0][0], lshape0=n0, seed=2, counts=500.0, device=dev)
                got = sharded.denoise4D_sharded(x, mu, 4, True, plan=plan, schedule="two_pass", engine=False)[0]
                other = _checksum(got[plan.owned_local[0]])
                del got, x
                torch.cuda.empty_cache()
                same = all_ok(mine == other, dev)
                tot = torch.tensor([mine >> 8], dtype=torch.int64, device=dev)     # >> 8: the sum over ranks stays in int64
                dist.all_reduce(tot)
                check["fullsize_engine_vs_nccl_two_pass"] = {"iterations": 4, "identical_on_all_ranks": same,
                                                             "recon_checksum": int(tot.item())}
                if not same:
                    if rank == 0:
                        print(json.dumps({"error": "full-size engine vs two-pass checksum mismatch", "check": check}),
                              file=sys.stderr, flush=True)
                    return 3
            else:
                check["fullsize_engine_vs_nccl_two_pass"] = {"skipped": "not enough free device memory on some rank",
                                                             "free_gb_rank0": torch.cuda.mem_get_info(dev)[0] / 1e9}
        if use_engine:
            eng = make_engine(sharded, gshape, world, rank, mu, max_it, local)
        else:
            x = synth.stem4d_device(gshape, offset0=plan.read[0][0], lshape0=n0, seed=2, counts=500.0, device=dev)
            if peer:      # owned planes only; the halo is read from the neighbours' HBM inside the fused kernel
                own = x[plan.owned_local[0]].contiguous()
                del x
                torch.cuda.empty_cache()
                x = own
                sh = sharded.PeerShard(plan, own, mu, None, fista=True, n_iter=n_total)
            else:
                sh = sharded.CudaShard(plan, x, mu, None, fista=True, n_iter=n_total, fused=fused)
        comm_stream = torch.cuda.Stream(device=dev)           # used only with CYTVDN_SHARD_EXCHANGE=overlap
        tk = 1.0
        it = 0

        def step():
            nonlocal tk, it
            if use_engine:
                eng.iterate(1, 0)
            else:
                tkr, tk = sharded.fista_ratio(tk)
                if peer:
                    sh.step(it, tkr, True)
                elif fused:
                    sharded._run_iteration_fused(sh, it, tkr, True, None, comm_stream)
                else:
                    sharded._run_iteration_overlapped(sh, it, tkr, True, None, comm_stream)
            it += 1

        def sync():
            if use_engine:
                eng.synchronize()
            torch.cuda.synchronize()

        for _ in range(args.warmup):
            step()
        sync()
        dist.barrier()
        sync()
        clocks = ClockSampler(local) if rank == 0 else None
        if clocks:
            clocks.start()
            time.sleep(0.3)
        dist.barrier()
        sync()
        if use_engine:
            # events on the engine's compute stream (torch.cuda.Event would see torch's current stream only)
            cs = torch.cuda.ExternalStream(eng_streams(eng)[0], device=dev)
            if args.timeline:
                eng.load(None)                                # per-phase events count from a load: warm up again
                dist.barrier()
                eng.profile(True)
                for _ in range(args.warmup):
                    eng.iterate(1, 0)
                eng.synchronize()
                dist.barrier()
        else:
            cs = torch.cuda.current_stream(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = eng.refresh().launches if use_engine else sh.launches
        e0.record(cs)
        for _ in range(args.steps):
            step()
        e1.record(cs)
        sync()
        dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        total_ms = float(ms[0])
        launches = torch.tensor([(eng.refresh().launches if use_engine else sh.launches) - l0], dtype=torch.int64, device=dev)
        dist.all_reduce(launches)
        clk = clocks.stop() if clocks else None
        gvox = int(np.prod(gshape))
        value = gvox * args.steps / (total_ms * 1e-3) / 1e9
        if use_engine:
            n_run = eng.refresh().it_run
            glob = torch.from_numpy(eng.sums(n_run)).to(dev)
        else:
            glob = (sh.sums[:, :3] if peer else sh.fused_local_sums() if fused else sh.local_sums()).clone()
            n_run = it
        dist.all_reduce(glob)
        last = glob[n_run - 1].cpu().numpy()
        timeline = None
        if use_engine and args.timeline:
            tl = eng.timeline(min(args.steps + args.warmup, 256))[args.warmup:]
            allt = [None] * world
            dist.all_gather_object(allt, tl.tolist())
            eng.profile(False)
            if rank == 0:
                cols = ["start_ms", "wait_neighbours_ms", "halo_planes_ms", "interior_ms", "halo_to_push_ms", "push_ms"]
                timeline = {"columns": cols, "per_rank_mean": [dict(zip(cols[1:], np.mean(np.array(t)[:, 1:], axis=0).tolist()))
                                                                for t in allt],
                            "per_rank_iterations": allt if args.timeline_full else None}
                try:
                    with open(args.timeline, "w") as f:
                        json.dump({"n_gpus": world, "shape": list(gshape), "ms_per_step": total_ms / args.steps,
                                   "timeline": timeline}, f, indent=1)
                except OSError:
                    pass
                timeline.pop("per_rank_iterations", None)
        peak, peak_src = measured_peak()
        local_vox = int(np.prod(plan.local_shape)) if not peer else plan.owned_voxels
        ms_per_step = total_ms / args.steps
        ach = (BYTES_A + BYTES_B) * local_vox / (ms_per_step * 1e-3) / 1e9       # 96 B/voxel contract figure
        roofline = {"bound": "hbm",
                    "kernel": ("per-GPU iteration incl. halo planes and exchange: tv_fused_kernel sweeps" if fused else
                               "per-GPU iteration incl. halo planes and exchange: half-step A + half-step B sweeps"),
                    "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                    "peak_source": peak_src, "bytes_per_voxel": BYTES_A + BYTES_B}
        if fused:
            moved = 76 * local_vox / (ms_per_step * 1e-3) / 1e9
            roofline["variant"] = ("fused single pass: 96 B/voxel of contract work done while moving 76 B/voxel, "
                                   "hence frac can exceed 1; see `moved`")
            roofline["moved"] = {"bytes_per_voxel": 76, "GB/s": moved, "frac": moved / peak}

        # ---- the headline measurement is complete.  What follows (end to end, the single-shard figure, the strong
        #      scaling point, the CPU baseline) are side measurements: a safety net makes sure the line is printed even
        #      if one of them should hang (a timer shorter than NCCL's timeout emits what is finished and ends the run).
        check.update({"bnorm_last": float(last[0]), "delta_last": float(last[1] / last[2]) if last[2] else None})
        extras = {"e2e": None, "alone": None, "strong": None, "cpu": None}

        def emit():
            if rank != 0:
                return
            line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                    "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                    "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                    "config": dict(workload_config(world) if not args.shape else
                                   {"workload": f"4-D FISTA fp32 sharded, {'x'.join(map(str, per))} per GPU (non-default)"},
                                   schedule=schedule),
                    "roofline": roofline, "cpu_baseline": extras["cpu"], "e2e": extras["e2e"], "gpu_launches": int(launches[0]),
                    "clocks": clk, "check": check, "single_gpu_same_shard": extras["alone"], "timeline": timeline,
                    "strong_scaling_config5": extras["strong"]}
            print(json.dumps(line), flush=True)

        import threading

        def emergency():
            for k in extras:
                if extras[k] is None:
                    extras[k] = {"value": None, "error": f"side measurement not finished after {args.extras_timeout} s; "
                                                          "the line was emitted by the safety net"}
            try:
                emit()
            finally:
                os._exit(0)
        net = threading.Timer(args.extras_timeout, emergency)
        net.daemon = True
        net.start()

        # ---- end to end: pinned host shard in, pinned host owned planes out, through the public sharded API ----
        e2e = None
        if not args.no_e2e:
            try:
                iters = args.e2e_iters
                if use_engine:
                    host_in = host_out = None
                    try:                                      # page-locked host memory may run out on one rank only
                        host_in = tv.pinned_empty(eng.local_shape, np.float32)
                        host_out = tv.pinned_empty(eng.owned_shape, np.float32)
                    except Exception:
                        pass
                    if not all_ok(host_out is not None, dev):
                        raise MemoryError("page-locked host buffers for the end-to-end run could not be allocated on every rank")
                    torch.from_numpy(host_in).copy_(eng.array("orig"))
                    torch.cuda.synchronize()
                    call = lambda n: sharded.denoise4D_engine(host_in, mu, n, True, gshape=gshape, out=host_out, engine=eng)
                else:
                    if peer:
                        sh.close()
                    del sh
                    sh = None
                    torch.cuda.empty_cache()
                    host_in = tv.pinned_empty(tuple(x.shape), np.float32)
                    host_out = tv.pinned_empty(tuple(x.shape), np.float32)
                    torch.from_numpy(host_in).copy_(x)
                    del x
                    torch.cuda.empty_cache()
                    torch.cuda.synchronize()

                    def call(n):
                        xd = torch.empty(host_in.shape, dtype=torch.float32, device=dev)
                        xd.copy_(torch.from_numpy(host_in), non_blocking=True)
                        if peer:
                            r = sharded.denoise4D_peer(xd, mu, n, True, plan=plan)
                        else:
                            r = sharded.denoise4D_sharded(xd, mu, n, True, plan=plan, schedule="fused" if fused else "two_pass",
                                                          engine=False)
                        torch.from_numpy(host_out).copy_(r[0], non_blocking=True)
                        torch.cuda.synchronize()
                        return r
                call(3)                                       # one short untimed call: first-use costs
                samples = []
                for _ in range(max(1, args.e2e_calls)):
                    torch.cuda.synchronize()
                    dist.barrier()
                    t0 = time.perf_counter()
                    r = call(iters)
                    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
                    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
                    samples.append(float(dt[0]))
                med = float(np.median(samples))
                nb_in, nb_out = int(np.prod(host_in.shape)) * 4, int(np.prod(host_out.shape)) * 4
                e2e = {"value": gvox * iters / med / 1e9, "unit": UNIT,
                       "h2d_bytes_per_step": nb_in * world / iters, "d2h_bytes_per_step": nb_out * world / iters,
                       "call": ("sharded.denoise4D_engine(pinned host shard in, pinned host owned planes out, "
                                f"iterations={iters}, FISTA=True) on every rank; shard arena kept across calls" if use_engine else
                                f"sharded.{'denoise4D_peer' if peer else 'denoise4D_sharded'}(shard from pinned host, "
                                f"iterations={iters}, FISTA=True) on every rank"),
                       "wall_s": med, "wall_s_samples": samples, "timed_calls": len(samples),
                       "spread": (max(samples) - min(samples)) / med, "delta_last": float(r[2][-1])}
                del host_in, host_out
            except Exception as e:          # e.g. not enough pinned host memory on this box
                e2e = {"value": None, "unit": UNIT, "error": repr(e)[:300]}
        extras["e2e"] = e2e

        # ---- like-for-like single-GPU figure: one shard of the same stored shape alone on rank 0 -------------
        alone = None
        if use_engine and not args.no_alone:
            close_engine(eng)
            eng = None
            torch.cuda.empty_cache()
            if rank == 0:
                g1 = tuple(plan.local_shape)
                e1_ = sharded.EngineShard(g1, 1, 0, mu, None, np.float32, fista=True, max_iters=args.warmup + args.steps, device=local)
                e1_.connect_all([e1_.export()])
                e1_.load_synth(seed=2, counts=500.0)
                for _ in range(args.warmup):
                    e1_.iterate(1, 0)
                e1_.synchronize()
                cs1 = torch.cuda.ExternalStream(eng_streams(e1_)[0], device=dev)
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record(cs1)
                for _ in range(args.steps):
                    e1_.iterate(1, 0)
                a1.record(cs1)
                e1_.synchronize()
                ms1 = a0.elapsed_time(a1) / args.steps
                e1_.close()
                alone = {"shape": list(g1), "ms_per_step": ms1, "value": int(np.prod(g1)) * 1e-9 / (ms1 * 1e-3),
                         "note": "one shard of the per-GPU stored shape (owned + overlap planes) alone on one GPU, no "
                                 "neighbours: ms_per_step of the N-GPU run / this = per-GPU efficiency like for like"}
                alone["efficiency_like_for_like"] = ms1 / ms_per_step
            dist.barrier()
        extras["alone"] = alone

        # ---- strong scaling of BASELINE config 5 itself (the fixed 1024x1024x128x128 array, 69 GB raw) ------------
        # 8 GPUs: the main line above IS config 5 (fused, 19 arrays = 166 GB per GPU).  4 GPUs: 258 planes per GPU do not
        # fit the fused schedule's second state set, but the in-place two-pass schedule does (10 arrays = 173 GB of the
        # 192 GB): run it here over NCCL.  2 GPUs: the state (687 GB) exceeds the two GPUs' HBM (384 GB); the out-of-core
        # schedule (cytvdn_denoise_sharded_streamed) would need 687 GB of page-locked host memory for the state between
        # passes -- stated, not run, when the box does not have it.
        strong = None
        full = (1024, 1024, 128, 128)
        if not args.shape and not args.no_strong:
            if world == 8:
                strong = {"n_gpus": 8, "shape": list(full), "schedule": schedule, "value": value, "ms_per_step": ms_per_step,
                          "note": "the main line of this run"}
            elif world == 4:
                try:
                    if eng is not None:
                        close_engine(eng)
                        eng = None
                    torch.cuda.empty_cache()
                    plan5 = sharded.ShardPlan(full, world, rank)
                    x5 = sh5 = None
                    try:                                      # allocation may fail on one rank only: decide together
                        x5 = synth.stem4d_device(full, offset0=plan5.read[0][0], lshape0=plan5.local_shape[0], seed=2, counts=500.0, device=dev)
                        sh5 = sharded.CudaShard(plan5, x5, mu, None, fista=True, n_iter=args.warmup + args.steps, fused=False)
                    except Exception as ex:
                        alloc_err = repr(ex)[:200]
                    if not all_ok(sh5 is not None, dev):
                        del sh5, x5
                        torch.cuda.empty_cache()
                        raise MemoryError("the in-place two-pass state (10 arrays, 173 GB per GPU) did not fit on every rank")
                    tk5, it5 = 1.0, 0
                    for k in range(args.warmup + args.steps):
                        if k == args.warmup:
                            torch.cuda.synchronize()
                            dist.barrier()
                            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                            s0.record()
                        tkr5, tk5 = sharded.fista_ratio(tk5)
                        sharded._run_iteration_overlapped(sh5, it5, tkr5, True, None, comm_stream)
                        it5 += 1
                    s1.record()
                    torch.cuda.synchronize()
                    t5 = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
                    dist.all_reduce(t5, op=dist.ReduceOp.MAX)
                    g5 = sh5.local_sums().clone()
                    dist.all_reduce(g5)
                    l5 = g5[it5 - 1].cpu().numpy()
                    strong = {"n_gpus": 4, "shape": list(full), "schedule": "two_pass (in place, 10 arrays = 173 GB per GPU; NCCL halo exchange)",
                              "value": int(np.prod(full)) * args.steps / (float(t5[0]) * 1e-3) / 1e9,
                              "ms_per_step": float(t5[0]) / args.steps, "delta_last": float(l5[1] / l5[2])}
                    del sh5, x5
                    torch.cuda.empty_cache()
                except Exception as ex:
                    strong = {"n_gpus": 4, "shape": list(full), "error": repr(ex)[:300]}
            elif world == 2:
                from bench import mem_available_gb
                strong = {"n_gpus": 2, "shape": list(full), "feasible": False,
                          "why": "state of config 5 = 10 arrays x 68.7 GB = 687 GB; 2 GPUs hold 384 GB of HBM, so only the "
                                 "out-of-core schedule applies (cytvdn_denoise_sharded_streamed), which keeps b and d in "
                                 f"page-locked host memory between passes: 550 GB + 137 GB for data and result; this box has "
                                 f"{mem_available_gb():.0f} GB.  tools/sharded_stream_bench.py runs that schedule at the largest "
                                 "size the host holds (profiles/r2_sharded_stream_bench.jsonl)."}
        extras["strong"] = strong
        if rank == 0 and not args.no_cpu:
            r_ = cpu_reference_run(5, 1, budget_s=15.0)
            extras["cpu"] = {"value": r_["value"], "unit": UNIT, "cores": r_["cores"], "kind": r_["kind"], "sample": r_["sample"]}
        net.cancel()
        emit()
    except BaseException:
        # a failing rank leaves at once (no collective tear-down the others might not join); torchrun then stops the rest
        import traceback
        traceback.print_exc()
        sys.stderr.flush()
        os._exit(1)
    if 'eng' in locals() and eng is not None:
        close_engine(eng)
    dist.destroy_process_group()
    return rc


def eng_streams(eng):
    import ctypes as C
    a, b = C.c_void_p(), C.c_void_p()
    eng._lib.check(eng.lib.cytvdn_shard_streams(eng.h, C.byref(a), C.byref(b)))
    return a.value, b.value
