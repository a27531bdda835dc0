"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/cytvdn_b200.h declares, the ctypes structures match the C layout, the Python mirror keeps the
reference's signatures, and the product path fails loudly without a GPU (no CPU fallback)."""
import ctypes as C
import inspect
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cytvdn_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cytvdn_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from cytvdn_b200 import _lib
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/cytvdn_b200.h but not exported"
    assert sorted(_lib.PROTOTYPES) == names, "ctypes prototypes and header disagree"
    assert lib.cytvdn_version() == 100


def test_ctypes_structs_match_c_layout():
    from cytvdn_b200 import _lib
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "cytvdn_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu\n", sizeof(cytvdn_step_opts), offsetof(cytvdn_step_opts, own_lo),
         offsetof(cytvdn_step_opts, zero_wrap_mask), offsetof(cytvdn_step_opts, l2_budget_bytes), (size_t)0);
  printf("%zu %zu %zu %zu %zu %zu\n", sizeof(cytvdn_denoise_params), offsetof(cytvdn_denoise_params, shape),
         offsetof(cytvdn_denoise_params, stopping_relative_change), offsetof(cytvdn_denoise_params, clip),
         offsetof(cytvdn_denoise_params, device), offsetof(cytvdn_denoise_params, stream));
  return 0; }'''
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "t.c")
        open(src, "w").write(prog)
        exe = os.path.join(tmp, "t")
        gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        subprocess.run([gcc, "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split("\n")
    a = [int(v) for v in out[0].split()]
    S = _lib.StepOpts
    assert a[:4] == [C.sizeof(S), S.own_lo.offset, S.zero_wrap_mask.offset, S.l2_budget_bytes.offset]
    b = [int(v) for v in out[1].split()]
    P = _lib.DenoiseParams
    assert b == [C.sizeof(P), P.shape.offset, P.stopping_relative_change.offset, P.clip.offset, P.device.offset,
                 P.stream.offset]


def test_python_mirror_keeps_reference_signatures():
    """Names, order and defaults of cyTVDN.py:19-31 and :250-260 (extras are keyword-only)."""
    import cytvdn_b200 as tv
    p4 = inspect.signature(tv.denoise4D).parameters
    pos4 = [(k, v.default) for k, v in p4.items() if v.kind == v.POSITIONAL_OR_KEYWORD]
    assert pos4 == [("datacube", inspect._empty), ("mu", inspect._empty), ("iterations", 10), ("FISTA", True),
                    ("stopping_relative_change", None), ("isotropic_R", False), ("isotropic_Q", False),
                    ("reference_data", None), ("BC_mode", 2), ("lam", None), ("quiet", False)]
    p3 = inspect.signature(tv.denoise3D).parameters
    pos3 = [(k, v.default) for k, v in p3.items() if v.kind == v.POSITIONAL_OR_KEYWORD]
    assert pos3 == [("datacube", inspect._empty), ("mu", inspect._empty), ("iterations", 7500),
                    ("stopping_relative_change", None), ("BC_mode", 2), ("FISTA", False), ("reference_data", None),
                    ("lam", None), ("quiet", False)]
    for name in ("accumulator_update_4D", "accumulator_update_4D_FISTA", "accumulator_update_3D",
                 "accumulator_update_3D_FISTA", "iso_accumulator_update_4D", "iso_accumulator_update_4D_FISTA",
                 "datacube_update_4D", "datacube_update_3D", "sum_square_error_4D", "sum_square_error_3D",
                 "check_memory"):
        assert callable(getattr(tv, name))
    assert list(inspect.signature(tv.accumulator_update_4D_FISTA).parameters) == ["a", "b", "d", "tk", "ax", "clip", "BC_mode"]
    assert list(inspect.signature(tv.iso_accumulator_update_4D_FISTA).parameters) == \
        ["a", "b1", "b2", "d1", "d2", "tk", "ax1", "ax2", "clip"]
    assert list(inspect.signature(tv.datacube_update_4D).parameters) == \
        ["orig", "recon", "b1", "b2", "b3", "b4", "lambda_mu", "BC_mode"]


def test_no_cpu_fallback_and_host_side_checks():
    import cytvdn_b200 as tv
    data = np.ones((2, 3, 4, 4), np.float32)
    mu = np.array([1, 1, .5, .5], np.float32)
    # the reference's assertions fire before anything touches the device
    with pytest.raises(AssertionError, match="datacube must be floating point datatype."):
        tv.denoise4D(data.astype(np.int16), mu, 2, quiet=True)
    with pytest.raises(AssertionError, match="Mu must have same dtype as datacube."):
        tv.denoise4D(data, mu.astype(np.float64), 2, lam=mu / 32, quiet=True)
    with pytest.raises(AssertionError, match="Parameters must satisfy"):
        tv.denoise3D(data[0], np.ones(3, np.float32), 2, lam=np.ones(3, np.float32), quiet=True)
    if tv.device_count() == 0:
        with pytest.raises(tv.CytvdnError, match="no CPU fallback"):
            tv.denoise4D(data, mu, 2, quiet=True)
        with pytest.raises(tv.CytvdnError, match="no CPU fallback"):
            tv.accumulator_update_4D(data, np.zeros_like(data), 0, 1.0)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cytvdn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt, f


def test_c_abi_argument_validation_without_gpu():
    """Bad arguments are rejected with CYTVDN_E_INVALID / _UNSUPPORTED and a message before any CUDA call."""
    from cytvdn_b200 import _lib
    lib = _lib.load()
    sh4 = (C.c_int64 * 4)(4, 4, 4, 4)
    one = C.c_void_p(16)            # never dereferenced: validation fails first
    err = lambda: lib.cytvdn_last_error().decode()
    assert lib.cytvdn_accumulator_update(5, sh4, 0, one, one, None, 0.0, 0, 1.0, 2, one, None, None) == 1
    assert "ndim must be 3 or 4" in err()
    assert lib.cytvdn_accumulator_update(4, sh4, 7, one, one, None, 0.0, 0, 1.0, 2, one, None, None) == 1
    assert "dtype" in err()
    assert lib.cytvdn_accumulator_update(4, sh4, 0, None, one, None, 0.0, 0, 1.0, 2, one, None, None) == 1
    assert "NULL" in err()
    assert lib.cytvdn_accumulator_update(4, sh4, 0, one, one, None, 0.0, 4, 1.0, 2, one, None, None) == 1
    assert "ax=4" in err()
    assert lib.cytvdn_accumulator_update(3, sh4, 0, one, one, None, 0.0, 3, 1.0, 2, one, None, None) == 1
    assert lib.cytvdn_accumulator_update(4, sh4, 0, one, one, None, 0.0, 0, 1.0, 4, one, None, None) == 1
    assert "BC_mode" in err()
    bad = (C.c_int64 * 4)(4, 0, 4, 4)
    assert lib.cytvdn_accumulator_update(4, bad, 0, one, one, None, 0.0, 0, 1.0, 2, one, None, None) == 1
    assert "shape[1]" in err()
    m1 = (C.c_int64 * 4)(1, 4, 4, 4)
    assert lib.cytvdn_accumulator_update(4, m1, 0, one, one, None, 0.0, 0, 1.0, 1, one, None, None) == 1
    assert "mirror" in err()
    assert lib.cytvdn_iso_accumulator_update(sh4, 0, one, one, one, None, None, 0.0, 1, 1, 1.0, one, None, None) == 1
    assert "different axes" in err()
    assert lib.cytvdn_iso_accumulator_update(sh4, 0, one, one, one, one, None, 0.0, 0, 1, 1.0, one, None, None) == 1
    bp = (C.c_void_p * 4)(16, 16, 16, 16)
    w = (C.c_double * 4)(.1, .1, .1, .1)
    assert lib.cytvdn_datacube_update(4, sh4, 0, one, one, one, bp, w, 1, one, None, None) == 4     # mirror: unsupported
    assert "undefined behaviour" in err()
    assert lib.cytvdn_fused_iteration(4, sh4, 0, one, one, one, bp, bp, None, bp, 0.0, w, w, 2, one, None, None) == 1
    assert "d_in and d_out" in err()
    assert lib.cytvdn_fused_iteration(4, sh4, 0, one, one, one, bp, bp, None, None, 0.0, w, w, 1, one, None, None) == 4
    P = _lib.DenoiseParams()
    P.ndim, P.dtype = 4, 0
    for k in range(4):
        P.shape[k] = 4
    P.iters_fista = -1
    assert lib.cytvdn_denoise(C.byref(P), one, one, None, None, None, None, None, None) == 1
    assert "negative iteration count" in err()
    P.iters_fista = 2
    assert lib.cytvdn_denoise(C.byref(P), one, one, None, None, None, None, None, None) == 1
    assert "alias" in err()
    two = C.c_void_p(32)
    P.bc_mode = 1
    assert lib.cytvdn_denoise(C.byref(P), one, two, None, None, None, None, None, None) == 4
    assert "BC_mode=3" in err()                        # the message points at the well-defined mirror
    P.bc_mode, P.isotropic_Q = 3, 1
    assert lib.cytvdn_denoise(C.byref(P), one, two, None, None, None, None, None, None) == 1
    assert "anisotropic update only" in err()
    P.isotropic_Q, P.schedule = 0, 2
    assert lib.cytvdn_fused_iteration(4, sh4, 0, one, one, one, bp, bp, None, None, 0.0, w, w, 3, one, None, None) == 1
    assert "out of place" in err()                      # BC_mode 3 is accepted by the fused iteration (round 2)
    assert lib.cytvdn_fused_iteration(4, sh4, 0, one, one, one, bp, bp, None, None, 0.0, w, w, 4, one, None, None) == 1
    assert "BC_mode must be 0, 2 or 3" in err()
    P.schedule = 0
    P.shape[2] = 1
    assert lib.cytvdn_denoise(C.byref(P), one, two, None, None, None, None, None, None) == 1
    assert "extent >= 2 on axis 2" in err()
    P.shape[2] = 4
    P.bc_mode, P.ndim, P.isotropic_R = 2, 3, 1
    assert lib.cytvdn_denoise(C.byref(P), one, two, None, None, None, None, None, None) == 1
    assert "4-D only" in err()
    n = C.c_int64(0)
    P.ndim, P.isotropic_R, P.iters_fista, P.iters_plain = 4, 0, 10, 0
    assert lib.cytvdn_denoise_workspace_bytes(C.byref(P), 1, 1, C.byref(n)) == 0
    scratch = ((11 * 4 * 16 * 8 + 255) // 256) * 256 + 8192    # reduction slots of 10 iterations (+ MSE[0])
    assert n.value == (8 * 2 + 1) * 4 * 256 + scratch   # fused: two b/d sets + the second recon buffer
    P.schedule = 1
    assert lib.cytvdn_denoise_workspace_bytes(C.byref(P), 0, 0, C.byref(n)) == 0
    assert n.value == (8 + 2) * 4 * 256 + scratch       # two-pass from host data: b, d, orig, recon
    # the optional reservation needs a device: without one it reports the CUDA error instead of crashing
    cnt = C.c_int(0)
    lib.cytvdn_denoise(C.byref(P), None, None, None, None, None, None, None, None)
    assert lib.cytvdn_last_trace(None, None, 0, C.byref(cnt)) == 0 and cnt.value == 0
    assert lib.cytvdn_launch_count() >= 0


def _build_c_example(tmp_path):
    exe = str(tmp_path / "denoise_c_abi")
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    lib = os.path.join(ROOT, "cytvdn_b200")
    subprocess.run([gcc, os.path.join(ROOT, "examples", "denoise_c_abi.c"), "-I", os.path.join(ROOT, "include"), "-L", lib,
                    "-lcytvdn_b200", f"-Wl,-rpath,{lib}", "-lm", "-o", exe], check=True)
    return exe


def test_c_example_links_against_the_c_abi(tmp_path):
    """A plain C host compiles against include/cytvdn_b200.h and links the shared library."""
    out = subprocess.run([_build_c_example(tmp_path)], check=True, capture_output=True, text=True).stdout
    assert "cytvdn ABI version 100" in out


@pytest.mark.gpu
def test_c_example_runs_on_gpu(tmp_path):
    # (with one GPU the example puts both shards on it: give their streams separate hardware queues)
    env = dict(os.environ, CUDA_DEVICE_MAX_CONNECTIONS="32")
    out = subprocess.run([_build_c_example(tmp_path)], check=True, capture_output=True, text=True, env=env, timeout=300).stdout
    assert "20 FISTA iterations" in out and "schedule 2" in out
    assert "on 2 shards, 0 voxels differ" in out          # cytvdn_denoise_sharded from plain C
