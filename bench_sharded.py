"""bench.py's N > 1 arm: BASELINE config 5 (4-D anisotropic FISTA, fp32, scan-axis shards with a one-plane
halo exchange), weak scaling with 128x1024x128x128 owned per GPU.  Launched by torchrun, one rank per GPU."""
from __future__ import annotations

import json
import os
import time

import numpy as np


def run_sharded(args):
    import torch
    import torch.distributed as dist
    import cytvdn_b200 as tv
    from cytvdn_b200 import sharded, synth
    from bench import (BYTES_A, BYTES_B, METRIC, MU, SHARD_PER_GPU, UNIT, ClockSampler, measured_peak,
                       workload_config)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        per = tuple(args.shape) if args.shape else SHARD_PER_GPU
        gshape = (per[0] * world,) + tuple(per[1:])
        plan = sharded.ShardPlan(gshape, world, rank)
        n0 = plan.local_shape[0]
        x = synth.stem4d_device(gshape, offset0=plan.read[0][0], lshape0=n0, seed=2, counts=500.0, device=dev)
        mu = np.array(MU, dtype=np.float32)
        n_total = args.warmup + args.steps
        fused = args.schedule in ("fused", "peer")
        peer = args.schedule == "peer"
        if peer:      # owned planes only; the halo is read from the neighbours' HBM inside the fused kernel
            own = x[plan.owned_local[0]].contiguous()
            del x
            torch.cuda.empty_cache()
            x = own
            sh = sharded.PeerShard(plan, own, mu, None, fista=True, n_iter=n_total)
        else:
            sh = sharded.CudaShard(plan, x, mu, None, fista=True, n_iter=n_total, fused=fused)
        comm_stream = torch.cuda.Stream(device=dev)       # used only with CYTVDN_SHARD_EXCHANGE=overlap
        tk = 1.0
        it = 0

        def step():
            nonlocal tk, it
            tkr, tk = sharded.fista_ratio(tk)
            if peer:
                sh.step(it, tkr, True)
            elif fused:
                sharded._run_iteration_fused(sh, it, tkr, True, None, comm_stream)
            else:
                sharded._run_iteration_overlapped(sh, it, tkr, True, None, comm_stream)
            it += 1

        for _ in range(args.warmup):
            step()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        clocks = ClockSampler(local) if rank == 0 else None
        if clocks:
            clocks.start()
            time.sleep(0.3)
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = sh.launches
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        total_ms = float(ms[0])
        launches = torch.tensor([sh.launches - l0], dtype=torch.int64, device=dev)
        dist.all_reduce(launches)
        clk = clocks.stop() if clocks else None
        gvox = int(np.prod(gshape))
        value = gvox * args.steps / (total_ms * 1e-3) / 1e9
        glob = (sh.sums[:, :3] if peer else sh.fused_local_sums() if fused else sh.local_sums()).clone()
        dist.all_reduce(glob)
        last = glob[it - 1].cpu().numpy()
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
        # ---- end to end through the public sharded API with pinned host buffers ----------------------
        e2e = None
        if peer:
            sh.close()
        del sh
        torch.cuda.empty_cache()
        if not args.no_e2e:
            try:
                iters = args.e2e_iters
                host_in = tv.pinned_empty(tuple(x.shape), np.float32)
                host_out = tv.pinned_empty(tuple(x.shape), np.float32)
                torch.from_numpy(host_in).copy_(x)
                del x
                torch.cuda.empty_cache()
                torch.cuda.synchronize()
                # one short untimed call first (3 iterations): first-use costs of the copy / exchange path
                xw = torch.empty(host_in.shape, dtype=torch.float32, device=dev)
                xw.copy_(torch.from_numpy(host_in), non_blocking=True)
                if peer:
                    sharded.denoise4D_peer(xw, mu, 3, True, plan=plan)
                else:
                    sharded.denoise4D_sharded(xw, mu, 3, True, plan=plan, schedule=args.schedule)
                del xw
                torch.cuda.empty_cache()
                torch.cuda.synchronize()
                dist.barrier()
                t0 = time.perf_counter()
                xd = torch.empty(host_in.shape, dtype=torch.float32, device=dev)
                xd.copy_(torch.from_numpy(host_in), non_blocking=True)
                if peer:
                    recon, bn, dl = sharded.denoise4D_peer(xd, mu, iters, True, plan=plan)
                else:
                    recon, bn, dl = sharded.denoise4D_sharded(xd, mu, iters, True, plan=plan, schedule=args.schedule)
                torch.from_numpy(host_out).copy_(recon, non_blocking=True)
                torch.cuda.synchronize()
                dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
                nbytes = int(np.prod(host_in.shape)) * 4
                e2e = {"value": gvox * iters / float(dt[0]) / 1e9, "unit": UNIT,
                       "h2d_bytes_per_step": nbytes * world / iters, "d2h_bytes_per_step": nbytes * world / iters,
                       "call": f"sharded.{'denoise4D_peer' if peer else 'denoise4D_sharded'}(shard from pinned host, "
                               f"iterations={iters}, FISTA=True) on every rank",
                       "wall_s": float(dt[0]), "delta_last": float(dl[-1])}
            except Exception as e:          # e.g. not enough pinned host memory on this box
                e2e = {"value": None, "unit": UNIT, "error": repr(e)[:200]}
        if rank == 0:
            line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                    "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                    "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                    "config": dict(workload_config(world) if not args.shape else
                                   {"workload": f"4-D FISTA fp32 sharded, {'x'.join(map(str, per))} per GPU (non-default)"},
                                   schedule=args.schedule),
                    "roofline": roofline, "cpu_baseline": None, "e2e": e2e, "gpu_launches": int(launches[0]),
                    "clocks": clk,
                    "check": {"bnorm_last": float(last[0]), "delta_last": float(last[1] / last[2]) if last[2] else None}}
            print(json.dumps(line), flush=True)
    finally:
        dist.destroy_process_group()
    return 0
