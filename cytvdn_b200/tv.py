"""Host-side mirror of the reference's Python interface for the TV-denoising hot path.

Same names, positional order, defaults, assertion messages and return values as
`cyTVDN/cyTVDN.py` (``denoise4D`` :19-31, ``denoise3D`` :250-260) and as the kernels that
`cyTVDN/__init__.py:1` re-exports (``tv.accumulator_update_4D`` ...), so callers such as
`cyTVDN/mpi.py:317-398` keep working.  Everything computes on the GPU through the C ABI of
``libcytvdn_b200.so``; arrays may be NumPy arrays (copied to the device and back, results land
in the caller's arrays exactly like the in-place Cython kernels) or CUDA ``torch`` tensors
(used in place, zero copy).  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import weakref
from contextlib import contextmanager
from typing import Optional

import numpy as np

from . import _lib
from ._lib import F32, F64, DenoiseParams, StepOpts, check

__all__ = [
    "denoise3D", "denoise4D", "check_memory", "pinned_empty", "workspace_reserve", "workspace_release",
    "accumulator_update_4D", "accumulator_update_4D_FISTA",
    "accumulator_update_3D", "accumulator_update_3D_FISTA",
    "iso_accumulator_update_4D", "iso_accumulator_update_4D_FISTA",
    "datacube_update_4D", "datacube_update_3D",
    "sum_square_error_4D", "sum_square_error_3D",
]


# ------------------------------------------------------------------------------------------------
# array plumbing
# ------------------------------------------------------------------------------------------------
def _is_torch(x) -> bool:
    return type(x).__module__.split(".")[0] == "torch"


def _np_dtype(x):
    if _is_torch(x):
        return np.dtype(str(x.dtype).replace("torch.", ""))
    return x.dtype


def _code(dt) -> int:
    return F32 if dt == np.float32 else F64


_CNAME = {np.dtype(np.float32): "float", np.dtype(np.float64): "double"}


def _check_kernel_args(ndim, first, *others):
    """Mimic the errors of the Cython fused-type dispatch."""
    dt = _np_dtype(first)
    if dt not in _CNAME or len(first.shape) != ndim:
        raise TypeError("No matching signature found")
    for x in others:
        if x is None:
            continue
        if len(x.shape) != ndim:
            raise ValueError(f"Buffer has wrong number of dimensions (expected {ndim}, got {len(x.shape)})")
        xd = _np_dtype(x)
        if xd != dt:
            raise ValueError(f"Buffer dtype mismatch, expected '{_CNAME[dt]}' but got "
                             f"'{_CNAME.get(xd, str(xd))}'")
        if tuple(x.shape) != tuple(first.shape):
            raise ValueError(f"array shapes differ: {tuple(x.shape)} vs {tuple(first.shape)}")
    return dt


def _shape_arr(shape):
    return (C.c_int64 * len(shape))(*[int(s) for s in shape])


@contextmanager
def _on_device(index: Optional[int]):
    lib = _lib.load()
    if index is None:
        yield
        return
    prev = C.c_int(0)
    check(lib.cytvdn_get_device(C.byref(prev)))
    if prev.value != index:
        check(lib.cytvdn_set_device(int(index)))
    try:
        yield
    finally:
        if prev.value != index:
            lib.cytvdn_set_device(prev.value)


class _Staged:
    """Device views of a group of arrays for one kernel call.

    CUDA torch tensors are used in place.  NumPy arrays (and CPU tensors) are uploaded to
    scratch device buffers; those listed as outputs are copied back on exit, which reproduces
    the in-place semantics of the Cython kernels.
    """

    def __init__(self, arrays, outputs):
        self.lib = _lib.load()
        _lib.require_gpu()
        self.arrays = arrays
        self.outputs = outputs
        self.ptrs = []
        self.owned = []
        self.stream = None
        self.device = None
        self.host_views = []
        cuda_tensors = [a for a in arrays if a is not None and _is_torch(a) and a.is_cuda]
        if cuda_tensors:
            import torch
            dev = cuda_tensors[0].device
            for a in arrays:
                if a is not None and not (_is_torch(a) and a.is_cuda and a.device == dev):
                    raise ValueError("all arrays of one call must live on the same CUDA device")
            self.device = dev.index if dev.index is not None else torch.cuda.current_device()
            self.stream = torch.cuda.current_stream(dev).cuda_stream

    def __enter__(self):
        lib = self.lib
        self._ctx = _on_device(self.device)
        self._ctx.__enter__()
        # ONE device allocation for everything this call stages (host arrays + the reduction slots)
        hosts, total = [], 256
        for a in self.arrays:
            if a is None or (_is_torch(a) and a.is_cuda):
                hosts.append(None)
                continue
            h = a.numpy() if _is_torch(a) else a
            if not h.flags["C_CONTIGUOUS"]:
                raise ValueError("ndarray is not C-contiguous")
            hosts.append((h, total))
            total += (h.nbytes + 255) // 256 * 256
        base = C.c_void_p()
        check(lib.cytvdn_malloc(C.byref(base), total))
        self.owned.append(base)
        self.sums = C.c_void_p(base.value)                      # first 256 bytes: up to 8 doubles of sums
        for a, hv in zip(self.arrays, hosts):
            if a is None:
                self.ptrs.append(None)
                self.host_views.append(None)
            elif hv is None:
                if not a.is_contiguous():
                    raise ValueError("ndarray is not C-contiguous")
                self.ptrs.append(a.data_ptr())
                self.host_views.append(None)
            else:
                h, off = hv
                check(lib.cytvdn_memcpy(C.c_void_p(base.value + off), h.ctypes.data, h.nbytes, self.stream))
                self.ptrs.append(base.value + off)
                self.host_views.append(h)
        return self

    def read_sums(self, n):
        out = (C.c_double * n)()
        check(self.lib.cytvdn_memcpy(out, self.sums, 8 * n, self.stream))
        check(self.lib.cytvdn_stream_synchronize(self.stream))
        return [float(v) for v in out]

    def __exit__(self, et, ev, tb):
        lib = self.lib
        try:
            if et is None:
                for idx in self.outputs:
                    h = self.host_views[idx]
                    if h is not None:
                        check(lib.cytvdn_memcpy(h.ctypes.data, self.ptrs[idx], h.nbytes, self.stream))
                check(lib.cytvdn_stream_synchronize(self.stream))
        finally:
            for p in self.owned:
                lib.cytvdn_free(p)
            self._ctx.__exit__(et, ev, tb)
        return False


def _vp(p):
    return C.c_void_p(p) if p is not None else None


# ------------------------------------------------------------------------------------------------
# step functions (names of cyTVDN/__init__.py:1)
# ------------------------------------------------------------------------------------------------
def _acc(ndim, a, b, d, tk, ax, clip, BC_mode):
    dt = _check_kernel_args(ndim, a, b, d)
    with _Staged([a, b, d], outputs=[1, 2]) as s:
        check(s.lib.cytvdn_accumulator_update(ndim, _shape_arr(a.shape), _code(dt), _vp(s.ptrs[0]), _vp(s.ptrs[1]),
                                              _vp(s.ptrs[2]), float(tk), int(ax), float(clip), int(BC_mode),
                                              s.sums, None, s.stream))
        return s.read_sums(1)[0]


def accumulator_update_4D(a, b, ax, clip, BC_mode=2):
    """b = clip(a - shift(a, ax) + b) in place; returns sum|b|.  Replaces anisotropic.pyx:17-84."""
    return _acc(4, a, b, None, 0.0, ax, clip, BC_mode)


def accumulator_update_4D_FISTA(a, b, d, tk, ax, clip, BC_mode=2):
    """As above with b = v + tk (v - d), d = v.  Replaces anisotropic.pyx:89-164."""
    if d is None:
        raise TypeError("No matching signature found")
    return _acc(4, a, b, d, tk, ax, clip, BC_mode)


def accumulator_update_3D(a, b, ax, clip, BC_mode=2):
    """Replaces anisotropic.pyx:169-237."""
    return _acc(3, a, b, None, 0.0, ax, clip, BC_mode)


def accumulator_update_3D_FISTA(a, b, d, tk, ax, clip, BC_mode=2):
    """Replaces anisotropic.pyx:243-317."""
    if d is None:
        raise TypeError("No matching signature found")
    return _acc(3, a, b, d, tk, ax, clip, BC_mode)


def _iso(a, b1, b2, d1, d2, tk, ax1, ax2, clip):
    dt = _check_kernel_args(4, a, b1, b2, d1, d2)
    with _Staged([a, b1, b2, d1, d2], outputs=[1, 2, 3, 4]) as s:
        check(s.lib.cytvdn_iso_accumulator_update(_shape_arr(a.shape), _code(dt), _vp(s.ptrs[0]), _vp(s.ptrs[1]),
                                                  _vp(s.ptrs[2]), _vp(s.ptrs[3]), _vp(s.ptrs[4]), float(tk),
                                                  int(ax1), int(ax2), float(clip), s.sums, None, s.stream))
        return s.read_sums(1)[0]


def iso_accumulator_update_4D(a, b1, b2, ax1, ax2, clip):
    """Joint 2-norm shrink of the (ax1, ax2) accumulators.  Replaces halfisotropic.pyx:17-97."""
    return _iso(a, b1, b2, None, None, 0.0, ax1, ax2, clip)


def iso_accumulator_update_4D_FISTA(a, b1, b2, d1, d2, tk, ax1, ax2, clip):
    """Replaces halfisotropic.pyx:102-188."""
    if d1 is None or d2 is None:
        raise TypeError("No matching signature found")
    return _iso(a, b1, b2, d1, d2, tk, ax1, ax2, clip)


def _dcu(ndim, orig, recon, bs, lambda_mu, BC_mode):
    dt = _check_kernel_args(ndim, orig, recon, *bs)
    w = lambda_mu.detach().cpu().numpy() if _is_torch(lambda_mu) else np.asarray(lambda_mu)
    if w.ndim != 1:
        raise ValueError(f"Buffer has wrong number of dimensions (expected 1, got {w.ndim})")
    if w.dtype != dt:
        raise ValueError(f"Buffer dtype mismatch, expected '{_CNAME[dt]}' but got '{_CNAME.get(w.dtype, str(w.dtype))}'")
    if w.shape[0] < ndim:
        raise ValueError("lambda_mu must have one entry per axis")
    wd = (C.c_double * ndim)(*[float(x) for x in w[:ndim]])
    with _Staged([orig, recon] + list(bs), outputs=[1]) as s:
        bp = (C.c_void_p * ndim)(*[s.ptrs[2 + k] for k in range(ndim)])
        check(s.lib.cytvdn_datacube_update(ndim, _shape_arr(orig.shape), _code(dt), _vp(s.ptrs[0]), _vp(s.ptrs[1]),
                                           _vp(s.ptrs[1]), bp, wd, int(BC_mode), s.sums, None, s.stream))
        num, den = s.read_sums(2)
    with np.errstate(all="ignore"):
        return float(np.float64(num) / np.float64(den))      # C division: 0/0 -> nan (utils.pyx:125)


def datacube_update_4D(orig, recon, b1, b2, b3, b4, lambda_mu, BC_mode=2):
    """recon = orig - sum_k lambda_mu[k] (b_k - shift(b_k,-1,k)) in place; returns
    sum|recon_new - recon_old| / sum|recon_old|.  Replaces utils.pyx:54-125."""
    return _dcu(4, orig, recon, (b1, b2, b3, b4), lambda_mu, BC_mode)


def datacube_update_3D(orig, recon, b1, b2, b3, lambda_mu, BC_mode=2):
    """Replaces utils.pyx:131-199."""
    return _dcu(3, orig, recon, (b1, b2, b3), lambda_mu, BC_mode)


def _sse(ndim, a, b):
    dt = _check_kernel_args(ndim, a, b)
    n = int(np.prod(a.shape))
    with _Staged([a, b], outputs=[]) as s:
        check(s.lib.cytvdn_sum_square_error(n, _code(dt), _vp(s.ptrs[0]), _vp(s.ptrs[1]), s.sums, s.stream))
        return s.read_sums(1)[0]


def sum_square_error_4D(a, b):
    """sum (a-b)^2 (not a mean).  Replaces utils.pyx:14-30."""
    return _sse(4, a, b)


def sum_square_error_3D(a, b):
    """Replaces utils.pyx:35-49."""
    return _sse(3, a, b)


# ------------------------------------------------------------------------------------------------
# pinned host memory (fast host<->device copies for the NumPy path)
# ------------------------------------------------------------------------------------------------
def pinned_empty(shape, dtype=np.float32) -> np.ndarray:
    """A page-locked NumPy array (freed when the array is garbage collected)."""
    lib = _lib.load()
    _lib.require_gpu()
    dtype = np.dtype(dtype)
    n = int(np.prod(shape))
    p = C.c_void_p()
    check(lib.cytvdn_host_alloc(C.byref(p), max(n, 1) * dtype.itemsize))
    buf = (C.c_char * (max(n, 1) * dtype.itemsize)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)
    weakref.finalize(buf, lib.cytvdn_host_free, p)
    return arr


# ------------------------------------------------------------------------------------------------
# optional reservation of the device working set (repeated calls skip cudaMalloc / cudaFree)
# ------------------------------------------------------------------------------------------------
def workspace_reserve(like=None, *, nbytes=None, iterations=10, FISTA=True, host_arrays=True, schedule=None,
                      device=None) -> int:
    """Make the library hold the device working set of later ``denoise3D`` / ``denoise4D`` calls.

    Either ``nbytes`` or ``like`` (an array / tensor, or a ``(shape, dtype)`` pair) with the call's ``iterations``
    / ``FISTA`` / ``schedule``; ``host_arrays``: the calls will pass NumPy arrays (input and result copies are part of
    the working set).  Calls whose state fits carve the block up instead of allocating (~86 GB and 60 ms .. 0.7 s
    per call for a 256x256x128x128 float32 cube); others allocate as before.  Returns the size in bytes.
    ``workspace_release()`` gives the memory back.  Not part of the reference API (its state is NumPy arrays)."""
    lib = _lib.load()
    _lib.require_gpu()
    if nbytes is None:
        if like is None:
            raise ValueError("workspace_reserve needs `like` or `nbytes`")
        shape, dt = (like[0], np.dtype(like[1])) if isinstance(like, tuple) else (tuple(like.shape), _np_dtype(like))
        if type(iterations) in (list, tuple):
            nF, nU = int(iterations[0]), int(iterations[1])
        else:
            nF, nU = int(iterations * bool(FISTA)), int(iterations * (not FISTA))
        P = DenoiseParams()
        P.ndim, P.dtype = len(shape), _code(dt)
        for k in range(len(shape)):
            P.shape[k] = int(shape[k])
        P.iters_fista, P.iters_plain, P.bc_mode = nF, nU, 2
        P.schedule = {None: 0, "auto": 0, "two_pass": 1, "fused": 2}[schedule]
        need = C.c_int64(0)
        check(lib.cytvdn_denoise_workspace_bytes(C.byref(P), int(not host_arrays), int(not host_arrays), C.byref(need)))
        nbytes = need.value
    with _on_device(device):
        check(lib.cytvdn_workspace_reserve(int(nbytes)))
    return int(nbytes)


def workspace_release(device=None) -> None:
    """Free the block ``workspace_reserve`` holds on the (current) device."""
    lib = _lib.load()
    with _on_device(device):
        check(lib.cytvdn_workspace_release())


# ------------------------------------------------------------------------------------------------
# drivers
# ------------------------------------------------------------------------------------------------
def _fmt_bytes(n) -> str:
    for unit in ("B", "KB", "MB", "GB", "TB"):
        if n < 1024 or unit == "TB":
            return f"{n:.0f} {unit}" if unit == "B" else f"{n:.1f} {unit}"
        n /= 1024.0


def _as_host_vector(x, name):
    if _is_torch(x):
        x = x.detach().cpu().numpy()
    return x


def _denoise(ndim, datacube, mu, iterations, FISTA, stopping_relative_change, isotropic_R, isotropic_Q,
             reference_data, BC_mode, lam, quiet, out, timing, schedule=None, devices=None):
    lib = _lib.load()
    torch_in = _is_torch(datacube)
    dt = _np_dtype(datacube)
    assert dt in (np.float32, np.float64), "datacube must be floating point datatype."

    mu = _as_host_vector(mu, "mu")
    if lam is None:
        lam = mu * 1.0 / 32.0 if ndim == 4 else mu / 16.0            # cyTVDN.py:67-68 / :294-295
    lam = _as_host_vector(lam, "lam")
    assert lam.dtype == dt, "Lambda must have same dtype as datacube."
    if ndim == 4:
        assert mu.dtype == dt, "Mu must have same dtype as datacube."   # 4-D only (cyTVDN.py:71)
    contiguous = datacube.is_contiguous() if torch_in else datacube.flags["C_CONTIGUOUS"]
    assert contiguous, ("datacube must be C-contiguous. Try np.ascontiguousarray(datacube) on the array"
                        + ("." if ndim == 4 else ""))
    if len(datacube.shape) != ndim:
        raise TypeError("No matching signature found")

    lambdaInv = 1.0 / lam                                            # cyTVDN.py:77 / :303
    lam_mu = (lam / mu).astype(dt)                                   # cyTVDN.py:78 / :304
    if ndim == 3:
        assert np.all(lam_mu <= (1.0 / 16.0)) & np.all(lam_mu > 0), "Parameters must satisfy 0 < λ/μ <= 1/8"
    if not quiet:
        try:
            print("λ/μ ≈ [" + ", ".join(f"1/{mu[k] / lam[k]:.0f}" for k in range(ndim)) + "]")
        except Exception:
            print("I tried to print with pretty characters but your system doesn't like Unicode...")
    if ndim == 4 and (np.any(lam_mu > (1.0 / 32.0)) or np.any(lam_mu <= 0)) and not quiet:
        print("WARNING: Parameters must satisfy 0 < λ/μ <= 1/32 or result may diverge!")

    unaccelerated = not FISTA                                        # cyTVDN.py:98-108 / :324-334
    if type(iterations) in (list, tuple):
        FISTA = True
        unaccelerated = True
        nF, nU = int(iterations[0]), int(iterations[1])
    else:
        nF, nU = int(iterations * FISTA), int(iterations * (not FISTA))

    if BC_mode == 1:
        raise NotImplementedError("BC_mode=1 (mirror) is undefined behaviour in the reference's datacube_update "
                                  "(utils.pyx:117-120) and is not implemented; BC_mode=3 is the well-defined "
                                  "(clamped-index) mirror")
    _lib.require_gpu()

    P = DenoiseParams()
    P.ndim, P.dtype = ndim, _code(dt)
    for k in range(ndim):
        P.shape[k] = int(datacube.shape[k])
        P.clip[k] = float(lambdaInv[k])
        P.lambda_mu[k] = float(lam_mu[k])
    P.iters_fista, P.iters_plain = nF, nU
    P.isotropic_R, P.isotropic_Q = int(bool(isotropic_R)), int(bool(isotropic_Q))
    P.bc_mode = int(BC_mode)
    P.use_stopping = int(stopping_relative_change is not None)
    P.stopping_relative_change = float(stopping_relative_change) if stopping_relative_change is not None else 0.0
    P.device = -1
    P.schedule = {None: 0, "auto": 0, "two_pass": 1, "fused": 2, "streamed": 3}[schedule]
    P.stream = None

    device = None
    if torch_in:
        import torch
        if not datacube.is_cuda:
            raise ValueError("torch input must be a CUDA tensor (pass a NumPy array for host data)")
        device = datacube.device.index if datacube.device.index is not None else torch.cuda.current_device()
        P.stream = torch.cuda.current_stream(datacube.device).cuda_stream
        if out is not None:
            # the library writes prod(shape) elements of the input's dtype through this pointer
            if not (_is_torch(out) and out.is_cuda):
                raise ValueError("out must be a CUDA tensor when datacube is one")
            if out.device != datacube.device:
                raise ValueError(f"out lives on {out.device}, datacube on {datacube.device}")
            if out.dtype != datacube.dtype:
                raise ValueError(f"out has dtype {out.dtype}, datacube {datacube.dtype}")
            if tuple(out.shape) != tuple(datacube.shape):
                raise ValueError(f"out has shape {tuple(out.shape)}, datacube {tuple(datacube.shape)}")
            if not out.is_contiguous():
                raise ValueError("out must be contiguous")
            lo, hi = out.data_ptr(), out.data_ptr() + out.numel() * out.element_size()
            dlo, dhi = datacube.data_ptr(), datacube.data_ptr() + datacube.numel() * datacube.element_size()
            if lo < dhi and dlo < hi:
                raise ValueError("out must not overlap datacube (the input is read in every iteration)")
        recon = out if out is not None else torch.empty_like(datacube)
        data_p, recon_p = datacube.data_ptr(), recon.data_ptr()
        keep = None
        if reference_data is not None:
            keep = reference_data if (_is_torch(reference_data) and reference_data.is_cuda) else \
                torch.as_tensor(np.asarray(reference_data), device=datacube.device)
            keep = keep.to(device=datacube.device, dtype=datacube.dtype).contiguous()
            if tuple(keep.shape) != tuple(datacube.shape):
                raise ValueError("reference_data must have the shape of datacube")
            ref_p = keep.data_ptr()
        else:
            ref_p = None
    else:
        recon = out if out is not None else np.empty_like(datacube)
        if not (isinstance(recon, np.ndarray) and recon.flags["C_CONTIGUOUS"] and recon.dtype == dt
                and recon.shape == datacube.shape and recon.flags["WRITEABLE"]):
            raise ValueError("out must be a writeable C-contiguous NumPy array with the dtype and shape of datacube")
        if np.shares_memory(recon, datacube):
            raise ValueError("out must not overlap datacube (the input is read in every iteration)")
        data_p, recon_p = datacube.ctypes.data, recon.ctypes.data
        if reference_data is not None:
            keep = np.ascontiguousarray(reference_data, dtype=dt)
            if keep.shape != datacube.shape:
                raise ValueError("reference_data must have the shape of datacube")
            ref_p = keep.ctypes.data
        else:
            keep, ref_p = None, None

    if not quiet:
        need, free_b, tot_b = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        with _on_device(device):
            check(lib.cytvdn_denoise_workspace_bytes(C.byref(P), int(torch_in), int(torch_in), C.byref(need)))
            check(lib.cytvdn_mem_info(C.byref(free_b), C.byref(tot_b)))
        print(f"Available GPU memory: {_fmt_bytes(free_b.value)}", flush=True)
        kind = "FISTA Accelerated" if FISTA else "Unaccelerated"
        print(f"{kind} TV denoising will require {_fmt_bytes(need.value)} of GPU memory...", flush=True)

    n = nF + nU
    bn = (C.c_double * max(n, 1))()
    dl = (C.c_double * max(n, 1))()
    ms = (C.c_double * (n + 1))()
    done = (C.c_int32 * 3)()
    tm = (C.c_double * 3)()
    if devices is not None:
        # several GPUs from this one process: scan-axis shards, halo planes pushed by the copy engines (cytvdn_shard.cu)
        devs = [int(d) for d in devices]
        if torch_in or reference_data is not None:
            raise ValueError("devices=[...] takes a NumPy datacube (host arrays in and out) and no reference_data")
        check(lib.cytvdn_denoise_sharded(C.byref(P), len(devs), (C.c_int * len(devs))(*devs), _vp(data_p), _vp(recon_p),
                                         bn, dl, done, tm))
    else:
        with _on_device(device):
            check(lib.cytvdn_denoise(C.byref(P), _vp(data_p), _vp(recon_p), _vp(ref_p), bn, dl,
                                     ms if reference_data is not None else None, done, tm))
    if timing is not None:
        cnt = C.c_int(0)
        tms, tws = (C.c_double * 12)(), (C.c_char_p * 12)()
        check(lib.cytvdn_last_trace(tms, tws, 12, C.byref(cnt)))
        timing["trace_ms"] = ({tws[k].decode(): round(tms[k], 3) for k in range(min(cnt.value, 12))}
                              if devices is None else {})
        timing["devices"] = (int(done[2]) >> 8) if devices is not None else 1
        timing.update(setup_ms=tm[0], loop_ms=tm[1], finish_ms=tm[2], iters_fista=int(done[0]),
                      iters_plain=int(done[1]),
                      schedule={1: "two_pass", 2: "fused", 3: "streamed"}.get(int(done[2]) & 0xff, "none"),
                      pipeline_boxes=(int(done[2]) >> 8) if ((int(done[2]) & 0xff) != 3 and devices is None) else 0,
                      stream_tiles=(int(done[2]) >> 8) if ((int(done[2]) & 0xff) == 3 and devices is None) else 0)
    with np.errstate(all="ignore"):
        b_norm = np.array(bn[:n], dtype=np.float64).astype(dt)
        delta_recon = np.array(dl[:n], dtype=np.float64).astype(dt)
    if (not quiet and unaccelerated and stopping_relative_change is not None and done[1] < nU):
        print(f"Stopping condition reached after {nF + done[1] - 1} iterations, stopping.")
    if reference_data is not None:
        with np.errstate(all="ignore"):
            MSE = np.array(ms[:n + 1], dtype=np.float64).astype(dt)
        return recon, b_norm, delta_recon, MSE
    return recon, b_norm, delta_recon


def denoise4D(datacube, mu, iterations=10, FISTA=True, stopping_relative_change=None, isotropic_R=False,
              isotropic_Q=False, reference_data=None, BC_mode=2, lam=None, quiet=False, *, out=None, timing=None,
              schedule=None, devices=None):
    """Proximal (an)isotropic TV denoising of a 4-D datacube on the GPU.

    Drop-in for ``cyTVDN.denoise4D`` (`cyTVDN/cyTVDN.py:19-247`): same arguments in the same order,
    same assertions, returns ``(recon, b_norm, delta_recon[, MSE])`` with ``b_norm``/``delta_recon`` of
    length ``iterations`` (trailing zeros after an early stop).  ``iterations`` may be ``[n_FISTA,
    n_unaccelerated]``.  ``BC_mode=1`` raises (undefined behaviour in the reference); ``BC_mode=3`` (not in the
    reference) is the well-defined mirror: half-step A as the reference's mirror (`anisotropic.pyx:69-70`), half-step
    B with the forward index of `utils.pyx:117-120` clamped to ``min(i+1, N-1)``; anisotropic only.

    Extras (keyword only): ``out`` -- array/tensor that receives ``recon`` (e.g. ``pinned_empty``);
    ``timing`` -- dict filled with setup/loop/finish milliseconds measured with CUDA events;
    ``schedule`` -- ``"fused"`` (one pass per iteration, 76 B/voxel, needs a second set of accumulator
    arrays), ``"two_pass"`` (96 B/voxel, in place) or ``None``: fused when it applies and fits in memory.
    ``"streamed"``: out of core -- host arrays larger than the GPU's memory are iterated tile by tile (temporal
    blocking over PCIe, `DESIGN.md` section 4); chosen automatically when nothing else fits.
    All schedules give bit-identical results.  ``devices=[0, 1, ...]``: shard scan axis 0 of a NumPy datacube over
    several GPUs from this process (the C ABI's ``cytvdn_denoise_sharded``, the counterpart of `cyTVDN/mpi.py`:
    anisotropic, ``BC_mode`` 2 or 0; ``schedule="streamed"`` runs shards larger than the GPUs out of core); the
    result equals the single-GPU one bit for bit.

    Early stopping (``stopping_relative_change``): ``delta_recon[i]`` is compared with the threshold exactly as in
    `cyTVDN.py:189-194`, but it is computed from float64 sums.  The reference accumulates ``delta`` in the array
    dtype, sequentially per OpenMP thread; in float32 that is off by 0.1 % .. 90 % for >= 4 M voxels and depends on
    the thread count, so on large float32 inputs the reference may stop at another iteration
    than this function (compare reconstructions at equal iteration counts).  Small arrays and float64 stop alike.
    """
    return _denoise(4, datacube, mu, iterations, FISTA, stopping_relative_change, isotropic_R, isotropic_Q,
                    reference_data, BC_mode, lam, quiet, out, timing, schedule, devices)


def denoise3D(datacube, mu, iterations=7_500, stopping_relative_change=None, BC_mode=2, FISTA=False,
              reference_data=None, lam=None, quiet=False, *, out=None, timing=None, schedule=None):
    """Drop-in for ``cyTVDN.denoise3D`` (`cyTVDN/cyTVDN.py:250-435`).  Note the positional order
    differs from ``denoise4D`` exactly as in the reference.  Early stopping compares a float64-accumulated ``delta``
    with the threshold (see ``denoise4D``): on large float32 cubes the reference's own float32 sums can make it stop
    at another iteration."""
    return _denoise(3, datacube, mu, iterations, FISTA, stopping_relative_change, False, False,
                    reference_data, BC_mode, lam, quiet, out, timing, schedule)


def check_memory(datacube):
    """GPU analogue of `cyTVDN/cyTVDN.py:438-467`: device memory needed by each algorithm."""
    lib = _lib.load()
    _lib.require_gpu()
    free_b, tot_b = C.c_int64(0), C.c_int64(0)
    check(lib.cytvdn_mem_info(C.byref(free_b), C.byref(tot_b)))
    nbytes = int(np.prod(datacube.shape)) * _np_dtype(datacube).itemsize
    nd = len(datacube.shape)
    rows = [("Anisotropic Unaccelerated", nbytes * (nd + 2)), ("Anisotropic FISTA", nbytes * (2 * nd + 2))]
    if nd == 4:
        rows += [("(Half-)Isotropic Unaccelerated", nbytes * (nd + 2)), ("(Half-)Isotropic FISTA", nbytes * (2 * nd + 2))]
    print(f"Datacube size is {_fmt_bytes(nbytes)} with dtype {_np_dtype(datacube)}; "
          f"GPU memory free {_fmt_bytes(free_b.value)} of {_fmt_bytes(tot_b.value)}")
    for name, need in rows:
        # what does not fit in HBM still runs from host arrays on the out-of-core schedule (tiles over PCIe)
        print(f"{name:34s} {_fmt_bytes(need):>10s}  {'OK' if need < free_b.value else 'OUT OF CORE (host arrays, schedule=streamed)'}")
    return {name: need for name, need in rows}
