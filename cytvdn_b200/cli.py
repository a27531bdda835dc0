"""``cyTVMPI``-style command line for the sharded GPU path.

Same flags as the reference's console script (`cyTVDN/mpi.py:47-76`, `setup.py:89`):

    torchrun --nproc-per-node 8 -m cytvdn_b200.cli -i in.npy -o out.npy -d 4 -f 1 -n 100 -L .03 .03 .015 .015 -m 1 1 .5 .5
    python -m cytvdn_b200.cli -i in.npy -o out.npy -d 4 -f 0 -n 50 -L ... -m ...        # one GPU

One process per GPU (torchrun / WORLD_SIZE); without torchrun it runs on one GPU.  As in the reference, ``-L`` is
lambda itself (no ``mu/32`` default, `mpi.py:249-250`), the data are processed as float32 (`mpi.py:220,230`),
``BC_mode`` is 2 (`mpi.py:84`), every rank reads only its own block (owned planes + one overlap plane per
neighbour, `mpi.py:165-180`) and writes only its owned block (`mpi.py:470-498`).

File formats: ``.npy`` (memory-mapped, any size) in and out; ``-o x.emd`` / ``.h5`` writes the reference's EMD v0.7
layout (`mpi.py:440-498`) when h5py is importable.  On one GPU the arrays stay on the host: the library
overlaps the PCIe copies with the iterations and switches to its out-of-core schedule (``--schedule streamed`` forces
it) when the state does not fit in the GPU's memory.  ``.h5`` / ``.emd`` input is read through h5py when
that package is importable (dataset path ``-p``); the reference's ``.dm3/.dm4`` readers (py4DSTEM / ncempy) are out
of scope.  Unlike the reference, FISTA (`mpi.py:310-311` "haven't done FISTA yet"),
hybrid iteration counts and 3-D input (single GPU only) work, and ``--stop`` enables the relative-change stopping
criterion that its to-do list mentions (`README.md:34`).
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np


def _str2bool(v):
    if isinstance(v, bool):
        return v
    if v.lower() in ("yes", "true", "t", "y", "1"):
        return True
    if v.lower() in ("no", "false", "f", "n", "0"):
        return False
    raise argparse.ArgumentTypeError("Boolean value expected.")


def build_parser():
    p = argparse.ArgumentParser(prog="cytvdn_b200.cli", description="Launch TV denoising on B200 GPUs (cyTVMPI flags).")
    p.add_argument("-i", "--input", type=os.path.abspath, nargs=1, required=True, help="input file (.npy, .h5/.emd)")
    p.add_argument("-o", "--output", type=os.path.abspath, nargs=1, required=True, help="output file (.npy, or .emd/.h5 with h5py)")
    p.add_argument("-d", "--dimensions", type=int, nargs=1, required=True, help="Number of Dimensions (3 or 4)")
    p.add_argument("-f", "--fista", type=_str2bool, nargs=1, default=[False], help="Use acceleration? 0 or 1.")
    p.add_argument("-n", "--niterations", type=int, nargs="+", required=True,
                   help="Number of iterations (Specify 2 values for hybrid.)")
    p.add_argument("-L", "--lambda", dest="lam", type=float, nargs="+", required=True)
    p.add_argument("-m", "--mu", type=float, nargs="+", required=True)
    p.add_argument("-v", "--verbose", type=_str2bool, default=True)
    # extras (not in the reference)
    p.add_argument("-p", "--dataset", default=None, help="HDF5 dataset path (h5/emd input)")
    p.add_argument("--stop", type=float, default=None, help="stopping_relative_change")
    p.add_argument("--grid", default="1d", choices=["1d", "mpi"], help="tile layout: axis-0 split or mpi.py's (wx, wy)")
    p.add_argument("--schedule", default="auto", choices=["auto", "fused", "two_pass", "streamed"],
                   help="streamed: out of core (one process: one GPU, or --devices N)")
    p.add_argument("--devices", type=int, default=0,
                   help="without torchrun: shard over this many GPUs from one process (cytvdn_denoise_sharded)")
    return p


def open_input(path, dataset=None):
    """A sliceable array-like (nothing is read yet)."""
    ext = os.path.splitext(path)[1].lower()
    if ext == ".npy":
        return np.load(path, mmap_mode="r")
    if ext in (".h5", ".hdf5", ".emd"):
        try:
            import h5py
        except ImportError as e:
            raise SystemExit(f"{path}: HDF5 input needs h5py, which is not installed ({e})")
        f = h5py.File(path, "r")
        if dataset is None:
            found = []
            f.visititems(lambda n, o: found.append(n) if hasattr(o, "shape") and len(o.shape) in (3, 4) else None)
            if not found:
                raise SystemExit(f"{path}: no 3-D / 4-D dataset found, pass -p")
            dataset = found[0]
        return f[dataset]
    raise SystemExit(f"{path}: unsupported input format {ext!r} (supported: .npy, .h5/.emd with h5py)")


def read_block(data, slices):
    """This rank's block as a contiguous float32 array (`mpi.py:216-239`)."""
    return np.array(data[slices], dtype=np.float32, order="C")       # always a private, writable copy


def is_hdf5(path):
    return os.path.splitext(path)[1].lower() in (".h5", ".hdf5", ".emd")


def create_emd(path, shape):
    """An empty EMD v0.7 file with the group layout the reference hard-codes (`mpi.py:440-490`): top-level group
    ``4DSTEM_experiment`` (emd_group_type 2, version 0.7), the (empty) data groups, ``data/datacubes/datacube_0``
    with the float32 dataset ``data`` and the four uncalibrated dimension vectors ``dim1..dim4``.  Needs h5py
    (optional dependency; the reference needs the parallel ``mpio`` build, here the ranks write one after the other)."""
    try:
        import h5py
    except ImportError as e:
        raise SystemExit(f"{path}: HDF5 / EMD output needs h5py, which is not installed ({e}); write .npy instead")
    with h5py.File(path, "w") as f:
        top = f.create_group("4DSTEM_experiment")
        top.attrs.create("emd_group_type", 2)
        top.attrs.create("version_major", 0)
        top.attrs.create("version_minor", 7)
        top.create_group("metadata")
        data = top.create_group("data")
        cubes = data.create_group("datacubes")
        for name in ("counted_datacubes", "diffractionslices", "realslices", "pointlists", "pointlistarrays"):
            data.create_group(name)
        dc = cubes.create_group("datacube_0")
        dc.create_dataset("data", tuple(shape), dtype="float32")
        dc.attrs.create("emd_group_type", 1)
        dc.attrs.create("metadata", -1)
        for k, (n, label) in enumerate(zip(shape, ("R_x", "R_y", "Q_x", "Q_y")), start=1):
            d = dc.create_dataset(f"dim{k}", (n,))
            d[...] = np.arange(0, n)
            d.attrs.create("name", np.bytes_(label))
            d.attrs.create("units", np.bytes_("[pix]"))


def create_output(path, shape):
    """The output file: a memory-mapped .npy (returned, so that a single process can write straight into it) or an
    EMD v0.7 HDF5 file (returns None: written block by block with ``write_block``)."""
    if is_hdf5(path):
        if len(shape) != 4:
            raise SystemExit("EMD output holds 4-D datacubes (mpi.py:440-498); write 3-D results as .npy")
        create_emd(path, shape)
        return None
    return np.lib.format.open_memmap(path, mode="w+", dtype=np.float32, shape=tuple(shape))


def write_block(path, slices, block):
    if is_hdf5(path):
        import h5py
        with h5py.File(path, "r+") as f:                      # dset.write_direct(recon, ...) of mpi.py:493-497
            f["4DSTEM_experiment/data/datacubes/datacube_0/data"][slices] = block
        return
    out = np.load(path, mmap_mode="r+")
    out[slices] = block
    out.flush()
    del out


def main(argv=None):
    args = build_parser().parse_args(argv)
    ndim = args.dimensions[0]
    fista = bool(args.fista[0]) if isinstance(args.fista, list) else bool(args.fista)
    niter = args.niterations
    iterations = niter[0] if len(niter) == 1 else [niter[0], niter[1]]
    if len(niter) == 2:
        fista = True
    lam = np.array(args.lam, dtype=np.float32)
    mu = np.array(args.mu, dtype=np.float32)
    if ndim not in (3, 4) or len(lam) != ndim or len(mu) != ndim:
        raise SystemExit("-d must be 3 or 4 and -L / -m need one value per dimension")

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    head = rank == 0
    say = (lambda *a: print(*a, flush=True)) if (head and args.verbose) else (lambda *a: None)

    import torch
    import cytvdn_b200 as tv
    from cytvdn_b200 import sharded

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    data = open_input(args.input[0], args.dataset)
    if len(data.shape) != ndim:
        raise SystemExit(f"input has {len(data.shape)} dimensions, -d says {ndim}")
    say(f"Loaded memory map. Data size is: {tuple(data.shape)}")
    t0 = time.time()

    if world == 1 and args.devices and args.devices > 1:
        # several GPUs from ONE process (cytvdn_denoise_sharded): host arrays in and out, shards of scan axis 0;
        # --schedule streamed additionally runs shards that do not fit their GPU out of core
        if ndim != 4:
            raise SystemExit("sharded runs exist for 4-D data only (as in the reference, mpi.py:252-255)")
        block = read_block(data, tuple(slice(None) for _ in range(ndim)))
        out = create_output(args.output[0], data.shape)
        tm = {}
        recon, bn, dl = tv.denoise4D(block, mu, iterations=iterations, FISTA=fista, stopping_relative_change=args.stop,
                                     lam=lam, quiet=True, out=out, timing=tm, devices=list(range(args.devices)),
                                     schedule="streamed" if args.schedule == "streamed" else None)
        if out is None:
            write_block(args.output[0], tuple(slice(None) for _ in range(ndim)), recon)
        else:
            out.flush()
        say(f"{args.devices} devices from one process, schedule: {tm.get('schedule')}")
        world = args.devices
    elif world == 1:
        # host arrays in and out: the library overlaps the PCIe copies with the iterations and, when the arrays do
        # not fit in the GPU's memory, iterates them tile by tile (out-of-core schedule); the result is written
        # straight into the memory-mapped output file
        block = read_block(data, tuple(slice(None) for _ in range(ndim)))
        fn = tv.denoise4D if ndim == 4 else tv.denoise3D
        out = create_output(args.output[0], data.shape)
        tm = {}
        kw = dict(iterations=iterations, FISTA=fista, stopping_relative_change=args.stop, lam=lam, quiet=True,
                  schedule=None if args.schedule == "auto" else args.schedule, out=out, timing=tm)
        recon, bn, dl = fn(block, mu, **kw)
        if out is None:
            write_block(args.output[0], tuple(slice(None) for _ in range(ndim)), recon)
        else:
            out.flush()
        say(f"schedule: {tm.get('schedule')}" + (f", {tm['stream_tiles']} tiles" if tm.get("stream_tiles") else ""))
    else:
        if ndim != 4:
            raise SystemExit("sharded runs exist for 4-D data only (as in the reference, mpi.py:252-255); "
                             "run 3-D data on one GPU")
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        try:
            plan = sharded.ShardPlan(data.shape, world, rank, None if args.grid == "1d" else "mpi")
            say(f"Dividing work over a {plan.grid[0]} by {plan.grid[1]} grid...")
            if args.schedule == "streamed":
                raise SystemExit("--schedule streamed under torchrun: use one process with --devices N instead "
                                 "(the out-of-core shards are driven from one process)")
            if plan.grid[1] == 1 and args.schedule in ("auto", "fused"):
                # 1-D split: the C-ABI shard engine (copy-engine halo exchange over CUDA IPC); the block goes
                # host -> device inside the library, the owned planes come back as a host array
                block = read_block(data, plan.read_global)
                own, bn, dl = sharded.denoise4D_engine(block, mu, iterations, fista, args.stop, gshape=data.shape, lam=lam)
                owned = own
            else:
                block = torch.from_numpy(read_block(data, plan.read_global)).to(dev)
                recon, bn, dl = sharded.denoise4D_sharded(block, mu, iterations, fista, args.stop, plan=plan, lam=lam,
                                                          schedule=args.schedule)
                owned = recon[plan.owned_local].cpu().numpy()
            if head:
                create_output(args.output[0], data.shape)
            dist.barrier()
            if is_hdf5(args.output[0]):
                for r in range(world):                        # serial HDF5: one writer at a time (the reference uses mpio)
                    if r == rank:
                        write_block(args.output[0], plan.owned_global, owned)
                    dist.barrier()
            else:
                write_block(args.output[0], plan.owned_global, owned)
                dist.barrier()
        finally:
            dist.destroy_process_group()
    n = int(np.count_nonzero(dl)) if len(dl) else 0
    say(f"{n} iterations on {world} GPU(s) in {time.time() - t0:.2f} s (incl. file I/O); "
        f"delta[-1] = {float(dl[n - 1]) if n else float('nan'):.3e}; wrote {args.output[0]}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
