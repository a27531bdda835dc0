"""Driver of the shard-engine GPU tests (run as a subprocess so that CUDA_DEVICE_MAX_CONNECTIONS is set before the
CUDA context exists, and so that a dead-locked exchange can be killed by a timeout instead of hanging pytest).

    python tests/engine_driver.py one_device             # all ranks of a run in this process on cuda:0
    python -m torch.distributed.run ... tests/engine_driver.py ipc [--share-gpu]     # one process per rank, CUDA IPC

Prints one JSON line per case: {"case": ..., "ok": bool, ...}; exit code 1 if any case fails.
"""
import json
import os
import sys

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np

import signal
signal.alarm(int(os.environ.get("ENGINE_DRIVER_ALARM", "240")))      # a dead-locked exchange must not hold the GPU box


def cases_small():
    # (global shape, world, iterations, periodic, dtype)
    return [((12, 6, 8, 16), 2, 9, False, "float32"),
            ((13, 5, 6, 12), 4, [5, 4], False, "float32"),            # uneven split (4, 4, 4, 1), hybrid counts
            ((9, 4, 5, 10), 3, 7, False, "float64"),
            ((12, 6, 8, 13), 3, 8, False, "float32"),                 # odd rows: padded internally
            ((8, 4, 6, 8), 2, 8, True, "float32"),                    # periodic, both neighbours are the same rank
            ((12, 4, 6, 8), 3, [4, 3], True, "float32"),
            ((6, 4, 6, 8), 6, 6, False, "float32"),                   # one owned plane per rank
            ((10, 4, 6, 8), 1, 5, False, "float32"),
            ((10, 4, 6, 8), 2, 6, False, "plain")]                    # unaccelerated only


def make(shape, dtype, seed=3):
    rng = np.random.default_rng(seed)
    dt = np.float64 if dtype == "float64" else np.float32
    return rng.poisson(rng.uniform(20, 500, shape)).astype(dt)


def one_device():
    import torch
    import cytvdn_b200 as tv
    from cytvdn_b200 import sharded
    bad = 0
    for shape, world, iters, periodic, dtype in cases_small():
        data = make(shape, dtype)
        fista = dtype != "plain"
        mu = np.array([1, 1, .5, .5], dtype=data.dtype)
        ref = tv.denoise4D(data, mu, iters, fista, BC_mode=0 if periodic else 2, quiet=True)
        g = torch.from_numpy(data).cuda()
        got, bn, dl = sharded.emulate_engine_on_one_device(g, mu, world, iters, fista, periodic=periodic)
        ok1 = bool(np.array_equal(got.cpu().numpy(), ref[0])) and np.allclose(dl, ref[2].astype(np.float64), rtol=1e-4) \
            and np.allclose(bn, ref[1].astype(np.float64), rtol=1e-4)
        # the C ABI's single-process loop with every "device" being cuda:0
        tm = {}
        out = tv.denoise4D(data, mu, iters, fista, BC_mode=0 if periodic else 2, quiet=True, devices=[0] * world, timing=tm)
        ok2 = bool(np.array_equal(out[0], ref[0])) and np.allclose(out[2], ref[2], rtol=1e-4) and tm["devices"] == world
        print(json.dumps({"case": [list(shape), world, iters, periodic, dtype], "engine": ok1, "c_loop": ok2,
                          "ok": ok1 and ok2}), flush=True)
        bad += not (ok1 and ok2)
    # host-pipelined run (cytvdn_shard_run_host): boxes along scan axis 1, wavefront over (box, iteration), per-box
    # halo pushes and counters; every shard enqueues its whole run at once
    for shape, world, iters, nbox, mirror in (((9, 16, 6, 8), 2, [7, 5], 4, False), ((13, 24, 5, 12), 3, 11, 6, False),
                                              ((8, 12, 6, 8), 2, 3, 3, False), ((10, 16, 6, 8), 1, 9, 4, False),
                                              ((9, 16, 6, 8), 3, [40, 3], 4, False), ((9, 16, 6, 8), 2, 9, 4, True)):
        data = make(shape, "float32", 11)
        mu = np.array([1, 1, .5, .5], dtype=np.float32)
        bc = 3 if mirror else 2
        ref = tv.denoise4D(data, mu, iters, True, BC_mode=bc, quiet=True)
        nF, nU = (iters, 0) if isinstance(iters, int) else iters
        os.environ["CYTVDN_SHARD_PIPELINE"] = str(nbox)
        eng = [sharded.EngineShard(shape, world, r, mu, None, np.float32, fista=True, max_iters=nF + nU, periodic=2 if mirror else False)
               for r in range(world)]
        handles = [e.export() for e in eng]
        outs = []
        for e in eng:
            e.connect_all(handles)
        # one host thread per shard, like one process per GPU: a shard's run is thousands of launches, and a host that
        # blocks on a full launch queue must not keep the OTHER shards' pushes from being enqueued
        import threading
        outs = [np.empty(e.owned_shape, np.float32) for e in eng]
        blks = [np.ascontiguousarray(data[e.read_lo:e.read_lo + e.n_local]) for e in eng]
        errs = []

        def work(k):
            try:
                eng[k].run_host(blks[k], outs[k], nF, nU)
            except Exception as ex:
                errs.append(repr(ex))
        th = [threading.Thread(target=work, args=(k,)) for k in range(world)]
        for t_ in th:
            t_.start()
        for t_ in th:
            t_.join()
        assert not errs, errs
        tot = np.zeros((nF + nU, 3))
        for e in eng:
            tot += e.sums(nF + nU)
        got = np.concatenate(outs, axis=0)
        for e in eng:
            e.close()
        del os.environ["CYTVDN_SHARD_PIPELINE"]
        ok = bool(np.array_equal(got, ref[0])) and np.allclose(tot[:, 1] / tot[:, 2], ref[2].astype(np.float64), rtol=1e-4)
        print(json.dumps({"case": ["run_host", list(shape), world, iters, nbox, mirror], "ok": ok}), flush=True)
        bad += not ok
    # BC_mode=3 (clamped mirror) sharded: the mirror applies at the global edges only
    data = make((14, 6, 8, 12), "float32", 9)
    mu = np.array([1, 1, .5, .5], dtype=np.float32)
    ref = tv.denoise4D(data, mu, [7, 3], True, BC_mode=3, quiet=True, schedule="two_pass")
    out = tv.denoise4D(data, mu, [7, 3], True, BC_mode=3, quiet=True, devices=[0, 0, 0])
    ok = bool(np.array_equal(out[0], ref[0])) and np.allclose(out[2], ref[2], rtol=1e-4)
    print(json.dumps({"case": "BC_mode=3 sharded", "ok": ok}), flush=True)
    bad += not ok
    # early stopping through the single-process loop: stops where the single-GPU loop stops
    data = make((12, 8, 8, 16), "float32", 5)
    mu = np.array([1, 1, .5, .5], dtype=np.float32)
    full = tv.denoise4D(data, mu, 30, True, quiet=True)
    thr = float(full[2][12]) * 1.0001
    ref = tv.denoise4D(data, mu, 30, True, thr, quiet=True)
    out = tv.denoise4D(data, mu, 30, True, thr, quiet=True, devices=[0, 0, 0])
    ok = bool(np.array_equal(out[0], ref[0])) and np.array_equal(out[2] != 0, ref[2] != 0)
    print(json.dumps({"case": "early stop", "ok": ok, "iterations": int(np.count_nonzero(ref[2]))}), flush=True)
    bad += not ok
    return 1 if bad else 0


def ipc(share_gpu):
    import torch
    import torch.distributed as dist
    import cytvdn_b200 as tv
    from cytvdn_b200 import sharded
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = 0 if share_gpu else int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo" if share_gpu else "nccl", rank=rank, world_size=world,
                            **({} if share_gpu else {"device_id": torch.device("cuda", local)}))
    bad = 0
    fdev = "cpu" if share_gpu else "cuda"                # gloo reduces host tensors, NCCL device tensors
    try:
        for shape, _, iters, periodic, dtype in cases_small():
            if shape[0] < world:
                continue
            data = make(shape, dtype)
            fista = dtype != "plain"
            mu = np.array([1, 1, .5, .5], dtype=data.dtype)
            ref = tv.denoise4D(data, mu, iters, fista, BC_mode=0 if periodic else 2, quiet=True)
            plan = sharded.ShardPlan(shape, world, rank, None, periodic)
            block = torch.from_numpy(np.ascontiguousarray(plan.extract(data))).cuda()
            own, bn, dl = sharded.denoise4D_engine(block, mu, iters, fista, gshape=shape, periodic=periodic)
            ok = bool(np.array_equal(own.cpu().numpy(), ref[0][plan.owned_global[0]])) and \
                np.allclose(dl.astype(np.float64), ref[2].astype(np.float64), rtol=1e-4)
            # early stopping: every rank stops at the single-GPU iteration
            flag = torch.tensor([int(ok)], device=fdev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if rank == 0:
                print(json.dumps({"case": [list(shape), world, iters, periodic, dtype], "ok": bool(flag.item())}), flush=True)
            bad += not bool(flag.item())
        # host arrays in and out: the pipelined run (boxes along scan axis 1) over real CUDA IPC
        for shape, iters, nbox in (((11, 16, 6, 8), [7, 5], 4), ((9, 24, 5, 12), 40, 6), ((8, 12, 6, 8), 3, 3)):
            data = make(shape, "float32", 11)
            mu = np.array([1, 1, .5, .5], dtype=np.float32)
            ref = tv.denoise4D(data, mu, iters, True, quiet=True)
            plan = sharded.ShardPlan(shape, world, rank)
            os.environ["CYTVDN_SHARD_PIPELINE"] = str(nbox)
            own, bn, dl = sharded.denoise4D_engine(np.ascontiguousarray(plan.extract(data)), mu, iters, True, gshape=shape)
            del os.environ["CYTVDN_SHARD_PIPELINE"]
            ok = bool(np.array_equal(own, ref[0][plan.owned_global[0]])) and np.allclose(dl.astype(np.float64), ref[2].astype(np.float64), rtol=1e-4)
            flag = torch.tensor([int(ok)], device=fdev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if rank == 0:
                print(json.dumps({"case": ["run_host", list(shape), world, iters, nbox], "ok": bool(flag.item())}), flush=True)
            bad += not bool(flag.item())
        data = make((12, 8, 8, 16), "float32", 5)
        mu = np.array([1, 1, .5, .5], dtype=np.float32)
        full = tv.denoise4D(data, mu, 30, True, quiet=True)
        thr = float(full[2][12]) * 1.0001
        ref = tv.denoise4D(data, mu, [30, 4], True, thr, quiet=True)
        plan = sharded.ShardPlan(data.shape, world, rank)
        block = torch.from_numpy(np.ascontiguousarray(plan.extract(data))).cuda()
        own, bn, dl = sharded.denoise4D_engine(block, mu, [30, 4], True, thr, gshape=data.shape)
        ok = bool(np.array_equal(own.cpu().numpy(), ref[0][plan.owned_global[0]])) and np.array_equal(dl != 0, ref[2] != 0)
        flag = torch.tensor([int(ok)], device=fdev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(json.dumps({"case": "early stop (hybrid)", "ok": bool(flag.item())}), flush=True)
        bad += not bool(flag.item())
    finally:
        dist.destroy_process_group()
    return 1 if bad else 0


if __name__ == "__main__":
    mode = sys.argv[1]
    sys.exit(one_device() if mode == "one_device" else ipc("--share-gpu" in sys.argv))
