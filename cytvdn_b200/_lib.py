"""ctypes binding of ``libcytvdn_b200.so`` (C ABI declared in ``include/cytvdn_b200.h``).

There is no CPU fallback anywhere in this package: if the shared library has not been built
(``make -C cytvdn_b200/csrc`` or ``python -c "import __graft_entry__ as g; g.build()"``) loading
fails with an explicit error, and every compute entry point fails if no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CYTVDN_LIB") or os.path.join(HERE, "libcytvdn_b200.so")     # CYTVDN_LIB: experimental builds

F32, F64 = 0, 1
OK = 0


class CytvdnError(RuntimeError):
    """A non-zero status from the C ABI (CUDA error, bad argument, ...)."""


class StepOpts(C.Structure):
    """``cytvdn_step_opts`` (include/cytvdn_b200.h)."""
    _fields_ = [
        ("box_lo", C.c_int64 * 2),
        ("box_hi", C.c_int64 * 2),
        ("own_lo", C.c_int64 * 2),
        ("own_hi", C.c_int64 * 2),
        ("zero_wrap_mask", C.c_int32),
        ("flags", C.c_int32),
        ("l2_budget_bytes", C.c_int64),
        ("row_pitch", C.c_int64),
        ("peer_lo_recon", C.c_void_p),
        ("peer_hi_recon", C.c_void_p),
        ("peer_hi_b0", C.c_void_p),
        ("peer_hi_d0", C.c_void_p),
        ("sse_reference", C.c_void_p),
    ]


class ShardParams(C.Structure):
    """``cytvdn_shard_params`` (include/cytvdn_b200.h)."""
    _fields_ = [
        ("dtype", C.c_int32),
        ("world", C.c_int32),
        ("rank", C.c_int32),
        ("periodic", C.c_int32),
        ("fista", C.c_int32),
        ("max_iters", C.c_int32),
        ("device", C.c_int32),
        ("reserved", C.c_int32),
        ("gshape", C.c_int64 * 4),
        ("clip", C.c_double * 4),
        ("lambda_mu", C.c_double * 4),
    ]


class DenoiseParams(C.Structure):
    """``cytvdn_denoise_params`` (include/cytvdn_b200.h)."""
    _fields_ = [
        ("ndim", C.c_int32),
        ("dtype", C.c_int32),
        ("shape", C.c_int64 * 4),
        ("iters_fista", C.c_int32),
        ("iters_plain", C.c_int32),
        ("isotropic_R", C.c_int32),
        ("isotropic_Q", C.c_int32),
        ("bc_mode", C.c_int32),
        ("use_stopping", C.c_int32),
        ("stopping_relative_change", C.c_double),
        ("clip", C.c_double * 4),
        ("lambda_mu", C.c_double * 4),
        ("device", C.c_int32),
        ("schedule", C.c_int32),
        ("stream", C.c_void_p),
    ]


_vp, _i64p, _dp = C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_double)
_vpp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); every symbol include/cytvdn_b200.h declares
PROTOTYPES = {
    "cytvdn_version": (C.c_int, []),
    "cytvdn_last_error": (C.c_char_p, []),
    "cytvdn_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "cytvdn_accumulator_update": (C.c_int, [C.c_int, _i64p, C.c_int, _vp, _vp, _vp, C.c_double, C.c_int,
                                            C.c_double, C.c_int, _vp, C.POINTER(StepOpts), _vp]),
    "cytvdn_iso_accumulator_update": (C.c_int, [_i64p, C.c_int, _vp, _vp, _vp, _vp, _vp, C.c_double, C.c_int,
                                                C.c_int, C.c_double, _vp, C.POINTER(StepOpts), _vp]),
    "cytvdn_accumulator_update_all": (C.c_int, [C.c_int, _i64p, C.c_int, _vp, _vpp, _vpp, C.c_double, _dp,
                                                C.c_int, C.c_int, C.c_int, _vp, C.POINTER(StepOpts), _vp]),
    "cytvdn_datacube_update": (C.c_int, [C.c_int, _i64p, C.c_int, _vp, _vp, _vp, _vpp, _dp, C.c_int, _vp,
                                         C.POINTER(StepOpts), _vp]),
    "cytvdn_fused_iteration": (C.c_int, [C.c_int, _i64p, C.c_int, _vp, _vp, _vp, _vpp, _vpp, _vpp, _vpp, C.c_double,
                                         _dp, _dp, C.c_int, _vp, C.POINTER(StepOpts), _vp]),
    "cytvdn_sum_square_error": (C.c_int, [C.c_int64, C.c_int, _vp, _vp, _vp, _vp]),
    "cytvdn_denoise": (C.c_int, [C.POINTER(DenoiseParams), _vp, _vp, _vp, _dp, _dp, _dp,
                                 C.POINTER(C.c_int32), _dp]),
    "cytvdn_denoise_workspace_bytes": (C.c_int, [C.POINTER(DenoiseParams), C.c_int, C.c_int, _i64p]),
    "cytvdn_workspace_reserve": (C.c_int, [C.c_int64]),
    "cytvdn_workspace_release": (C.c_int, []),
    "cytvdn_workspace_info": (C.c_int, [_i64p, _i64p]),
    "cytvdn_last_trace": (C.c_int, [_dp, C.POINTER(C.c_char_p), C.c_int, C.POINTER(C.c_int)]),
    "cytvdn_shard_create": (C.c_int, [C.POINTER(ShardParams), _vpp]),
    "cytvdn_shard_destroy": (C.c_int, [_vp]),
    "cytvdn_shard_disconnect": (C.c_int, [_vp]),
    "cytvdn_shard_info": (C.c_int, [_vp, _i64p]),
    "cytvdn_shard_export": (C.c_int, [_vp, C.POINTER(C.c_ubyte)]),
    "cytvdn_shard_connect": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_ubyte)]),
    "cytvdn_shard_load": (C.c_int, [_vp, _vp]),
    "cytvdn_shard_iterate": (C.c_int, [_vp, C.c_int, C.c_int]),
    "cytvdn_shard_run_host": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int]),
    "cytvdn_shard_synchronize": (C.c_int, [_vp]),
    "cytvdn_shard_sums": (C.c_int, [_vp, _dp, C.c_int]),
    "cytvdn_shard_store": (C.c_int, [_vp, _vp]),
    "cytvdn_shard_array": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vpp]),
    "cytvdn_shard_streams": (C.c_int, [_vp, _vpp, _vpp]),
    "cytvdn_shard_profile": (C.c_int, [_vp, C.c_int]),
    "cytvdn_shard_timeline": (C.c_int, [_vp, _dp, C.c_int]),
    "cytvdn_denoise_sharded": (C.c_int, [C.POINTER(DenoiseParams), C.c_int, C.POINTER(C.c_int), _vp, _vp, _dp, _dp,
                                         C.POINTER(C.c_int32), _dp]),
    "cytvdn_denoise_sharded_streamed": (C.c_int, [C.POINTER(DenoiseParams), C.c_int, C.POINTER(C.c_int), _vp, _vp, _dp,
                                                  _dp, C.POINTER(C.c_int32), _dp]),
    "cytvdn_pipeline_schedule": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int64, _i64p]),
    "cytvdn_stream_plan": (C.c_int, [C.POINTER(DenoiseParams), C.c_int64, _i64p]),
    "cytvdn_stream_plan_sharded": (C.c_int, [C.POINTER(DenoiseParams), C.c_int64, C.c_int, _i64p]),
    "cytvdn_synth_counts": (C.c_int, [_i64p, C.c_int64, C.c_int64, C.c_int, _vp, _vp, C.c_double, C.c_uint64,
                                      _vp, _vp]),
    "cytvdn_malloc": (C.c_int, [_vpp, C.c_int64]),
    "cytvdn_free": (C.c_int, [_vp]),
    "cytvdn_host_alloc": (C.c_int, [_vpp, C.c_int64]),
    "cytvdn_host_free": (C.c_int, [_vp]),
    "cytvdn_memcpy": (C.c_int, [_vp, _vp, C.c_int64, _vp]),
    "cytvdn_memset": (C.c_int, [_vp, C.c_int, C.c_int64, _vp]),
    "cytvdn_stream_synchronize": (C.c_int, [_vp]),
    "cytvdn_set_device": (C.c_int, [C.c_int]),
    "cytvdn_get_device": (C.c_int, [C.POINTER(C.c_int)]),
    "cytvdn_mem_info": (C.c_int, [_i64p, _i64p]),
    "cytvdn_ipc_get_handle": (C.c_int, [_vp, C.POINTER(C.c_ubyte)]),
    "cytvdn_ipc_open": (C.c_int, [C.POINTER(C.c_ubyte), _vpp]),
    "cytvdn_ipc_close": (C.c_int, [_vp]),
    "cytvdn_launch_count": (C.c_int64, []),
}

_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CytvdnError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built. "
            "Run `make -C cytvdn_b200/csrc` (needs nvcc) -- there is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != OK:
        msg = load().cytvdn_last_error().decode("utf-8", "replace")
        if rc == 4:
            raise NotImplementedError(msg)
        if rc == 1:
            raise ValueError(msg)
        if rc == 3:
            raise MemoryError(msg)
        raise CytvdnError(f"libcytvdn_b200 error {rc}: {msg}")


def device_count() -> int:
    n = C.c_int(0)
    check(load().cytvdn_device_count(C.byref(n)))
    return n.value


def require_gpu() -> None:
    if device_count() < 1:
        raise CytvdnError("cytvdn_b200 needs an NVIDIA GPU (built for B200 / sm_100a); none is visible "
                          "and there is no CPU fallback.")


def launch_count() -> int:
    return int(load().cytvdn_launch_count())
