#!/usr/bin/env python
"""Real multi-GPU parity of the sharded path (NCCL, one process per GPU):

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/check_sharded_nccl.py

Every rank builds the same global array (seeded / hashed generator), runs ``denoise4D_sharded`` on its block
(both schedules, 1-D split; 2-D ``grid="mpi"`` too when N allows, FISTA, hybrid and early stopping), rank 0 also
runs the single-GPU ``tv.denoise4D`` on the whole array; the gathered sharded result must equal it bit for bit and
the all-reduced ``b_norm`` / ``delta`` must agree to 1e-6.  Prints one JSON line per case on rank 0.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np


def main():
    import torch
    import torch.distributed as dist
    import cytvdn_b200 as tv
    from cytvdn_b200 import sharded, synth

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    ok_all = True
    try:
        mu = np.array([1, 1, .5, .5], dtype=np.float32)
        cases = []
        grids = [None] + ([(2, world // 2)] if world % 2 == 0 and world >= 4 else [])
        for periodic in (False, True):
            for grid in grids:
                for sched in ("fused", "two_pass") + (("peer",) if grid is None else ()):
                    for iters, stop in ((25, None), ([8, 5], None), (60, 0.002)):
                        cases.append((grid, sched, iters, stop, periodic))
        for grid, sched, iters, stop, periodic in cases:
            gshape = (8 * world + 3, 13, 32, 64)           # uneven splits
            plan = sharded.ShardPlan(gshape, world, rank, grid, periodic)
            whole = synth.stem4d_device(gshape, seed=11, counts=400.0, device=dev)
            shard = plan.extract(whole).contiguous()
            out = torch.zeros(gshape, dtype=torch.float32, device=dev)
            if sched == "peer":       # owned planes only, halo read from the neighbours' HBM through CUDA IPC
                recon, bn, dl = sharded.denoise4D_peer(whole[plan.owned_global].contiguous(), mu, iters, True, stop,
                                                       plan=plan)
                out[plan.owned_global] = recon
            else:
                recon, bn, dl = sharded.denoise4D_sharded(shard, mu, iters, True, stop, plan=plan, schedule=sched)
                out[plan.owned_global] = recon[plan.owned_local]     # gather the owned blocks on rank 0
            dist.all_reduce(out)                            # blocks are disjoint: the sum is the assembly
            if rank == 0:
                ref = tv.denoise4D(whole, mu, iters, True, stop, BC_mode=0 if periodic else 2, quiet=True,
                                   schedule="two_pass")
                same = bool(torch.equal(out, ref[0]))
                n = int(np.count_nonzero(ref[2]))
                e_bn = float(np.max(np.abs(bn[:n].astype(np.float64) - ref[1][:n]) / ref[1][:n])) if n else 0.0
                e_dl = float(np.max(np.abs(dl[:n].astype(np.float64) - ref[2][:n]) / ref[2][:n])) if n else 0.0
                ok = same and e_bn < 1e-6 and e_dl < 1e-6 and int(np.count_nonzero(dl)) == n
                ok_all &= ok
                print(json.dumps({"world": world, "grid": list(plan.grid), "periodic": periodic, "schedule": sched,
                                  "iterations": iters,
                                  "stopping": stop, "iterations_run": n, "recon_bit_identical": same,
                                  "bnorm_max_rel": e_bn, "delta_max_rel": e_dl, "ok": ok}), flush=True)
            dist.barrier()
        flag = torch.tensor([1 if ok_all else 0], device=dev)
        dist.broadcast(flag, 0)
        ok_all = bool(flag.item())
    finally:
        dist.destroy_process_group()
    sys.exit(0 if ok_all else 1)


if __name__ == "__main__":
    main()
