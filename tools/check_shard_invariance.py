#!/usr/bin/env python
"""Shard-count / schedule invariance at FULL size: the global array is generated
on the devices from the hash of the global index, denoised for K iterations over N ranks, and an order-independent
64-bit checksum of the bit patterns of the owned reconstruction (plus b_norm / delta) is printed.  Equal checksums
for different N (and for the two schedules) mean bit-identical global results.

    torchrun --nproc-per-node N tools/check_shard_invariance.py --shape 512 1024 128 128 --iters 5 [--schedule two_pass]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", type=int, nargs=4, default=[512, 1024, 128, 128])
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--schedule", default="fused")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from cytvdn_b200 import sharded, synth
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        plan = sharded.ShardPlan(a.shape, world, rank)
        x = synth.stem4d_device(a.shape, offset0=plan.read[0][0], lshape0=plan.local_shape[0], seed=2, counts=500.0, device=dev)
        mu = np.array([1, 1, .5, .5], np.float32)
        recon, bn, dl = sharded.denoise4D_sharded(x, mu, a.iters, True, plan=plan, schedule=a.schedule)
        own = recon[plan.owned_local].contiguous()
        cs = own.view(torch.int32).to(torch.int64).sum().reshape(1)
        if world > 1:
            dist.all_reduce(cs)
        if rank == 0:
            print(json.dumps({"shape": a.shape, "world": world, "schedule": a.schedule, "iters": a.iters,
                              "recon_checksum": int(cs.item()) & 0xFFFFFFFFFFFFFFFF,
                              "b_norm": [float(v) for v in bn], "delta": [float(v) for v in dl]}), flush=True)
    finally:
        if world > 1:
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
