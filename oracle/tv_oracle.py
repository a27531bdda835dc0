"""CPU oracle for the cyTVDN hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import this module.  Nothing under ``cytvdn_b200/`` does.

Two kernel providers with one interface:

* ``PortKernels``       -- ``oracle/tv_oracle.c`` (plain C restatement of SURVEY.md §3.4) via ctypes;
* ``ReferenceKernels``  -- the UNMODIFIED reference kernels compiled into ``oracle/_ref`` by
                           ``oracle/build_ref.py`` (present in the authoring container and, as
                           prebuilt binaries, on the GPU box).

and a restatement of the reference's Python host loops (`cyTVDN/cyTVDN.py:19-247` ``denoise4D``,
`:250-435` ``denoise3D``) that can drive either provider.

Parity pin (``tests/test_oracle_pin.py``): PortKernels == ReferenceKernels bit-for-bit on arrays,
and on the returned array-dtype scalars at ``OMP_NUM_THREADS=1``; host loop == the reference's own
``tv.denoise3D/4D`` through the golden vectors of ``tests/golden``.

Scalars come in two flavours:
  ``scalars="T"``  array-dtype accumulation, exactly what the reference returns;
  ``scalars="D"``  float64 accumulation over the same arrays: the truth for ``bnorm``/``delta``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
import sysconfig

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(HERE, "_build", "libtv_oracle.so")
_REF_DIR = os.path.join(HERE, "_ref")


def build_port(force: bool = False) -> str:
    """Compile oracle/tv_oracle.c (gcc -O2 -fopenmp -ffp-contract=off)."""
    src = os.path.join(HERE, "tv_oracle.c")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src)):
        return _LIB_PATH
    subprocess.run(["make", "-C", HERE, "-s", "-B", "all"], check=True)
    return _LIB_PATH


_lib = None


def _load():
    global _lib
    if _lib is None:
        build_port()
        _lib = C.CDLL(_LIB_PATH)
        pf, pd, pi = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int64)
        for sfx, p, t in (("f32", pf, C.c_float), ("f64", pd, C.c_double)):
            for acc in ("T", "D"):
                f = getattr(_lib, f"tvo_acc_{sfx}_{acc}")
                f.restype = C.c_double
                f.argtypes = [p, p, p, pi, C.c_int, t, t, C.c_int]
                f = getattr(_lib, f"tvo_iso_{sfx}_{acc}")
                f.restype = C.c_double
                f.argtypes = [p, p, p, p, p, pi, C.c_int, C.c_int, t, t]
                f = getattr(_lib, f"tvo_dcu_{sfx}_{acc}")
                f.restype = None
                f.argtypes = [p, p, p, p, p, p, p, pi, C.c_int, C.c_int, pd]
                f = getattr(_lib, f"tvo_sse_{sfx}_{acc}")
                f.restype = C.c_double
                f.argtypes = [p, p, C.c_int64]
        _lib.tvo_max_threads.restype = C.c_int
    return _lib


def max_threads() -> int:
    return int(_load().tvo_max_threads())


def set_threads(n: int) -> None:
    """OpenMP thread count of the C port (the reference kernels obey OMP_NUM_THREADS instead)."""
    _load().tvo_set_threads(int(n))


def _sfx(a: np.ndarray) -> str:
    if a.dtype == np.float32:
        return "f32"
    if a.dtype == np.float64:
        return "f64"
    raise TypeError("oracle: float32/float64 only")


def _ptr(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    ct = C.c_float if a.dtype == np.float32 else C.c_double
    return a.ctypes.data_as(C.POINTER(ct))


def _shape4(a: np.ndarray):
    s = list(a.shape) + [1] * (4 - a.ndim)
    return (C.c_int64 * 4)(*s)


class PortKernels:
    """oracle/tv_oracle.c behind the reference's kernel names.  ``acc``: "T" or "D"."""

    name = "port"

    def __init__(self, acc: str = "D"):
        assert acc in ("T", "D")
        self.acc = acc
        self.lib = _load()

    # anisotropic.pyx:17 / :89 / :169 / :243
    def accumulator_update(self, a, b, d, tk, ax, clip, BC_mode=2):
        if BC_mode == 3:          # this repo's well-defined mirror: half-step A is the reference's BC_mode=1
            BC_mode = 1
        if BC_mode == 1 and a.shape[ax] < 2:
            raise ValueError("mirror boundary needs extent >= 2")
        f = getattr(self.lib, f"tvo_acc_{_sfx(a)}_{self.acc}")
        for x in (b, d):
            assert x is None or (x.dtype == a.dtype and x.shape == a.shape)
        return float(f(_ptr(a), _ptr(b), _ptr(d), _shape4(a), int(ax), float(tk), float(clip), int(BC_mode)))

    # halfisotropic.pyx:17 / :102
    def iso_accumulator_update(self, a, b1, b2, d1, d2, tk, ax1, ax2, clip):
        assert a.ndim == 4
        f = getattr(self.lib, f"tvo_iso_{_sfx(a)}_{self.acc}")
        return float(f(_ptr(a), _ptr(b1), _ptr(b2), _ptr(d1), _ptr(d2), _shape4(a), int(ax1), int(ax2),
                       float(tk), float(clip)))

    # utils.pyx:54 / :131 ; returns (sum|delta|, sum|old|)
    def datacube_update_sums(self, orig, recon, bs, lambda_mu, BC_mode=2):
        if BC_mode == 1:
            raise ValueError("BC_mode=1 is undefined behaviour in the reference (utils.pyx:117-120)")
        # BC_mode 3 (NOT in the reference): the forward index of utils.pyx:117-120 clamped, min(i+1, N-1)
        f = getattr(self.lib, f"tvo_dcu_{_sfx(orig)}_{self.acc}")
        w = np.ascontiguousarray(lambda_mu, dtype=orig.dtype)
        bs = list(bs) + [bs[0]] * (4 - len(bs))
        sums = (C.c_double * 2)()
        f(_ptr(orig), _ptr(recon), _ptr(bs[0]), _ptr(bs[1]), _ptr(bs[2]), _ptr(bs[3]), _ptr(w),
          _shape4(orig), orig.ndim, 1 if BC_mode == 3 else 0, sums)
        return float(sums[0]), float(sums[1])

    def datacube_update(self, orig, recon, bs, lambda_mu, BC_mode=2):
        s = self.datacube_update_sums(orig, recon, bs, lambda_mu, BC_mode)
        if self.acc == "T":  # the reference divides in the array dtype (utils.pyx:125)
            t = orig.dtype.type
            with np.errstate(all="ignore"):
                return float(t(s[0]) / t(s[1]))
        with np.errstate(all="ignore"):
            return float(np.float64(s[0]) / np.float64(s[1]))

    # utils.pyx:14 / :35
    def sum_square_error(self, a, b):
        f = getattr(self.lib, f"tvo_sse_{_sfx(a)}_{self.acc}")
        return float(f(_ptr(a), _ptr(b), a.size))


def reference_available() -> bool:
    suf = sysconfig.get_config_var("EXT_SUFFIX")
    return all(os.path.exists(os.path.join(_REF_DIR, "cyTVDN", m + suf))
               for m in ("anisotropic", "halfisotropic", "utils"))


class ReferenceKernels:
    """The compiled, unmodified reference kernels (oracle/_ref).  ``acc="T"`` returns what the
    reference returns; ``acc="D"`` re-sums the reference's arrays in float64 with numpy."""

    name = "reference"

    def __init__(self, acc: str = "T"):
        assert acc in ("T", "D")
        if not reference_available():
            raise RuntimeError("oracle/_ref not built (run python oracle/build_ref.py)")
        self.acc = acc
        if _REF_DIR not in sys.path:
            sys.path.insert(0, _REF_DIR)
        from cyTVDN import anisotropic, halfisotropic, utils  # type: ignore
        self.an, self.hi, self.ut = anisotropic, halfisotropic, utils

    @staticmethod
    def _abs_sum(*arrs):
        return float(sum(np.abs(x, dtype=np.float64).sum() for x in arrs))

    def accumulator_update(self, a, b, d, tk, ax, clip, BC_mode=2):
        t = a.dtype.type
        if BC_mode == 3:          # this repo's mirror: half-step A is the reference's (well-defined) BC_mode=1
            BC_mode = 1
        if a.ndim == 4:
            r = (self.an.accumulator_update_4D(a, b, ax, t(clip), BC_mode) if d is None else
                 self.an.accumulator_update_4D_FISTA(a, b, d, t(tk), ax, t(clip), BC_mode))
        else:
            r = (self.an.accumulator_update_3D(a, b, ax, t(clip), BC_mode) if d is None else
                 self.an.accumulator_update_3D_FISTA(a, b, d, t(tk), ax, t(clip), BC_mode))
        return float(r) if self.acc == "T" else self._abs_sum(b)

    def iso_accumulator_update(self, a, b1, b2, d1, d2, tk, ax1, ax2, clip):
        t = a.dtype.type
        r = (self.hi.iso_accumulator_update_4D(a, b1, b2, ax1, ax2, t(clip)) if d1 is None else
             self.hi.iso_accumulator_update_4D_FISTA(a, b1, b2, d1, d2, t(tk), ax1, ax2, t(clip)))
        return float(r) if self.acc == "T" else self._abs_sum(b1, b2)

    def datacube_update_sums(self, orig, recon, bs, lambda_mu, BC_mode=2):
        old = recon.copy()
        self._dcu(orig, recon, bs, lambda_mu, BC_mode)
        # |new - old| with the difference rounded in the array dtype (utils.pyx:103), summed in float64
        return float(np.abs((recon - old).astype(np.float64)).sum()), self._abs_sum(old)

    def _dcu(self, orig, recon, bs, lambda_mu, BC_mode):
        if BC_mode not in (0, 2):     # 1: out-of-bounds reads in the reference; 3 does not exist there
            raise ValueError(f"the reference's datacube_update is defined for BC_mode 0 and 2 only (got {BC_mode})")
        w = np.ascontiguousarray(lambda_mu, dtype=orig.dtype)
        if orig.ndim == 4:
            return self.ut.datacube_update_4D(orig, recon, bs[0], bs[1], bs[2], bs[3], w, BC_mode)
        return self.ut.datacube_update_3D(orig, recon, bs[0], bs[1], bs[2], w, BC_mode)

    def datacube_update(self, orig, recon, bs, lambda_mu, BC_mode=2):
        if self.acc == "T":
            return float(self._dcu(orig, recon, bs, lambda_mu, BC_mode))
        s = self.datacube_update_sums(orig, recon, bs, lambda_mu, BC_mode)
        with np.errstate(all="ignore"):
            return float(np.float64(s[0]) / np.float64(s[1]))

    def sum_square_error(self, a, b):
        if self.acc == "T":
            return float(self.ut.sum_square_error_4D(a, b) if a.ndim == 4 else self.ut.sum_square_error_3D(a, b))
        t = (a - b)
        return float((t * t).astype(np.float64).sum())


def default_kernels(acc: str = "D"):
    """Reference kernels when built, else the C port."""
    return ReferenceKernels(acc) if reference_available() else PortKernels(acc)


# ---------------------------------------------------------------------------------------------
# Host loops (restatement of cyTVDN/cyTVDN.py).  One routine for 3-D and 4-D.
# ---------------------------------------------------------------------------------------------

def fista_ratio(tk: float):
    """cyTVDN.py:154-156 / :373-375.  Returns (tk_ratio, tk_next); float64 on the host."""
    tk_new = (1 + np.sqrt(1 + 4 * tk ** 2)) / 2
    return (tk - 1.0) / tk_new, tk_new


def _denoise(datacube, mu, iterations, FISTA, stopping_relative_change, isotropic_R, isotropic_Q,
             reference_data, BC_mode, lam, kernels, scalars, state_out=None):
    ndim = datacube.ndim
    K = kernels if kernels is not None else default_kernels("T" if scalars == "T" else "D")
    assert datacube.dtype in (np.float32, np.float64), "datacube must be floating point datatype."
    if lam is None:
        lam = mu * 1.0 / 32.0 if ndim == 4 else mu / 16.0          # cyTVDN.py:67-68 / :294-295
    assert lam.dtype == datacube.dtype, "Lambda must have same dtype as datacube."
    if ndim == 4:
        assert mu.dtype == datacube.dtype, "Mu must have same dtype as datacube."   # :71 (4-D only)
    assert datacube.flags["C_CONTIGUOUS"]
    lambdaInv = 1.0 / lam                                                            # :77 / :303
    lam_mu = (lam / mu).astype(datacube.dtype)                                       # :78 / :304
    if ndim == 3:
        assert np.all(lam_mu <= (1.0 / 16.0)) & np.all(lam_mu > 0), "Parameters must satisfy 0 < λ/μ <= 1/8"

    unaccelerated = not FISTA                                                        # :98-108 / :324-334
    if type(iterations) in (list, tuple):
        FISTA = True
        unaccelerated = True
        nF, nU = iterations[0], iterations[1]
    else:
        nF, nU = iterations * FISTA, iterations * (not FISTA)

    sdt = datacube.dtype if scalars == "T" else np.float64
    calc = reference_data is not None
    if calc:
        MSE = np.zeros((nF + nU + 1,), dtype=sdt)
        MSE[0] = K.sum_square_error(datacube, reference_data)                        # :122-125
    b_norm = np.zeros((nF + nU), dtype=sdt)                                          # :127-128
    delta_recon = np.zeros_like(b_norm)
    acc = [np.zeros_like(datacube) for _ in range(ndim)]                             # :131-134
    dd = [np.zeros_like(datacube) for _ in range(ndim)] if FISTA else None           # :137-141
    tk = 1.0
    recon = datacube.copy()                                                          # :145

    def half_step_a(i, tkr, fista):
        dsel = (lambda k: dd[k]) if fista else (lambda k: None)
        if ndim == 4 and isotropic_R:                                                # :159-162
            b_norm[i] += K.iso_accumulator_update(recon, acc[0], acc[1], dsel(0), dsel(1), tkr, 0, 1, lambdaInv[0])
        else:
            b_norm[i] += K.accumulator_update(recon, acc[0], dsel(0), tkr, 0, lambdaInv[0], BC_mode)
            b_norm[i] += K.accumulator_update(recon, acc[1], dsel(1), tkr, 1, lambdaInv[1], BC_mode)
        if ndim == 4 and isotropic_Q:                                                # :170-173
            b_norm[i] += K.iso_accumulator_update(recon, acc[2], acc[3], dsel(2), dsel(3), tkr, 2, 3, lambdaInv[2])
        else:
            for ax in range(2, ndim):
                b_norm[i] += K.accumulator_update(recon, acc[ax], dsel(ax), tkr, ax, lambdaInv[ax], BC_mode)

    if FISTA:
        for i in range(int(nF)):                                                     # :148-194
            tk_ratio, tk = fista_ratio(tk)
            half_step_a(i, tk_ratio, True)
            delta_recon[i] = K.datacube_update(datacube, recon, acc, lam_mu, BC_mode)   # :182-184
            if calc:
                MSE[i + 1] = K.sum_square_error(reference_data, recon)
            if stopping_relative_change is not None and delta_recon[i] < stopping_relative_change:
                break
    if unaccelerated:
        for j in range(int(nU)):                                                     # :195-242
            i = j + nF
            half_step_a(i, 0.0, False)
            delta_recon[i] = K.datacube_update(datacube, recon, acc, lam_mu, BC_mode)
            if calc:
                MSE[i + 1] = K.sum_square_error(reference_data, recon)
            if stopping_relative_change is not None and delta_recon[i] < stopping_relative_change:
                break
    if state_out is not None:
        state_out.update(acc=acc, d=dd)
    if calc:
        return recon, b_norm, delta_recon, MSE
    return recon, b_norm, delta_recon


def denoise4D(datacube, mu, iterations=10, FISTA=True, stopping_relative_change=None, isotropic_R=False,
              isotropic_Q=False, reference_data=None, BC_mode=2, lam=None, quiet=True, *,
              kernels=None, scalars="T", state_out=None):
    """Restatement of cyTVDN.py:19-247 (same positional order)."""
    assert datacube.ndim == 4
    return _denoise(datacube, mu, iterations, FISTA, stopping_relative_change, isotropic_R, isotropic_Q,
                    reference_data, BC_mode, lam, kernels, scalars, state_out)


def denoise3D(datacube, mu, iterations=7_500, stopping_relative_change=None, BC_mode=2, FISTA=False,
              reference_data=None, lam=None, quiet=True, *, kernels=None, scalars="T", state_out=None):
    """Restatement of cyTVDN.py:250-435 (same positional order -- differs from denoise4D)."""
    assert datacube.ndim == 3
    return _denoise(datacube, mu, iterations, FISTA, stopping_relative_change, False, False,
                    reference_data, BC_mode, lam, kernels, scalars, state_out)
