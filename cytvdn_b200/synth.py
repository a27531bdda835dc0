"""Deterministic synthetic inputs for benchmarks and parity runs.

Count-like data on purpose: with unit-variance noise the clip (32/mu, 16/mu) is never reached and
the shrink path would go untested.

* ``eels_cube``      -- 3-D EELS-like cube (power-law background + two edges), Poisson noise, host.
* ``stem4d_poisson`` -- 4D-STEM-like datacube (central + four Bragg disks), Poisson noise, host.
* ``stem4d_device``  -- the same scene generated ON THE DEVICE from a hash of the global linear
                        index, so any axis-0 sharding yields the identical global array and arrays
                        larger than host RAM never exist on the host (BASELINE config 5).
* ``stem4d_hash_numpy`` -- bit-identical NumPy mirror of ``stem4d_device`` for CPU-side checks.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

_IH_SCALE = np.float32(2.6429137e-05)      # 1 / sqrt(4 * (65536^2 - 1) / 12): Irwin-Hall(4) of 16-bit fields


def eels_clean(shape=(128, 128, 1024), dose=1000.0, dtype=np.float32):
    """Noise-free expectation of ``eels_cube`` (float32 by default; e.g. to draw the Poisson noise on the device)."""
    X, Y, E = shape
    e = np.arange(E, dtype=np.float64)
    bg = (e + 0.05 * E) ** -1.5
    bg /= bg.max()
    edges = []
    for ek in (0.35 * E, 0.65 * E):
        ed = np.zeros(E)
        m = e > ek
        ed[m] = (e[m] - ek + 1.0) ** -0.6
        edges.append(ed)
    x = np.arange(X, dtype=np.float64)[:, None]
    y = np.arange(Y, dtype=np.float64)[None, :]
    w1 = (0.5 * (1 + np.tanh((x - X / 2) / max(X / 16, 1e-9))) * np.ones((1, Y))).astype(dtype)
    w2 = (0.5 * (1 + np.sin(2 * np.pi * y / max(Y / 2, 1e-9))) * np.ones((X, 1))).astype(dtype)
    out = np.empty(shape, dtype=dtype)
    bg_d, e1, e2 = (dose * bg).astype(dtype), (dose * 0.25 * edges[0]).astype(dtype), (dose * 0.15 * edges[1]).astype(dtype)
    for i in range(X):                      # plane by plane: no float64 temporaries of the full cube
        out[i] = bg_d[None, :] + w1[i][:, None] * e1[None, :] + w2[i][:, None] * e2[None, :]
    return out


def eels_cube(shape=(128, 128, 1024), seed=0, dose=1000.0, gain=1.0, dtype=np.float32):
    X, Y, E = shape
    rng = np.random.default_rng(seed)
    e = np.arange(E, dtype=np.float64)
    bg = (e + 0.05 * E) ** -1.5
    bg /= bg.max()
    edges = []
    for ek in (0.35 * E, 0.65 * E):
        ed = np.zeros(E)
        m = e > ek
        ed[m] = (e[m] - ek + 1.0) ** -0.6
        edges.append(ed)
    x = np.arange(X, dtype=np.float64)[:, None]
    y = np.arange(Y, dtype=np.float64)[None, :]
    w1 = 0.5 * (1 + np.tanh((x - X / 2) / max(X / 16, 1e-9))) * np.ones((1, Y))
    w2 = 0.5 * (1 + np.sin(2 * np.pi * y / max(Y / 2, 1e-9))) * np.ones((X, 1))
    clean = dose * (bg[None, None, :] + 0.25 * w1[..., None] * edges[0] + 0.15 * w2[..., None] * edges[1])
    return (gain * rng.poisson(clean)).astype(dtype)


def _stem_tables(gshape):
    """Scan modulation [N0*N1] and diffraction template [N2*N3] as float32 tables (host, float64 math)."""
    n0, n1, q2, q3 = gshape
    x = np.arange(n0, dtype=np.float64)[:, None]
    y = np.arange(n1, dtype=np.float64)[None, :]
    mod = (1.0 + 0.5 * np.sin(x / 3.0) * np.cos(y / 4.0)).astype(np.float32)
    k = np.arange(q2, dtype=np.float64)[:, None] - (q2 - 1) / 2.0
    l = np.arange(q3, dtype=np.float64)[None, :] - (q3 - 1) / 2.0
    rad = max(min(q2, q3) / 10.0, 0.5)
    templ = ((k ** 2 + l ** 2) <= rad ** 2).astype(np.float64)
    for dk, dl in ((q2 / 4.0, 0), (-q2 / 4.0, 0), (0, q3 / 4.0), (0, -q3 / 4.0)):
        templ += 0.3 * (((k - dk) ** 2 + (l - dl) ** 2) <= rad ** 2)
    return np.ascontiguousarray(mod.reshape(-1)), np.ascontiguousarray(templ.astype(np.float32).reshape(-1))


def stem4d_poisson(shape=(64, 64, 64, 64), seed=2, counts=500.0, dtype=np.float32):
    mod, templ = _stem_tables(shape)
    clean = (np.float32(counts) * mod[:, None] * templ[None, :] + np.float32(0.02 * counts)).reshape(shape)
    rng = np.random.default_rng(seed)
    return rng.poisson(clean.astype(np.float64)).astype(dtype)


def _mix64(z):
    z = (z + np.uint64(0x9E3779B97F4A7C15))
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def stem4d_hash_numpy(gshape, offset0=0, lshape0=None, seed=2, counts=500.0, dtype=np.float32):
    """NumPy mirror of the device generator (same float32 operation order, no FMA)."""
    n0, n1, q2, q3 = gshape
    lshape0 = n0 - offset0 if lshape0 is None else lshape0
    mod, templ = _stem_tables(gshape)
    m = q2 * q3
    st0 = n1 * m
    with np.errstate(over="ignore"):
        g = np.arange(offset0 * st0, (offset0 + lshape0) * st0, dtype=np.uint64)
        h = _mix64(np.uint64(seed) ^ _mix64(g))
    ij = (g // np.uint64(m)).astype(np.int64)
    kl = (g % np.uint64(m)).astype(np.int64)
    cts = np.float32(counts)
    c = (cts * mod[ij]) * templ[kl] + np.float32(0.02) * cts
    s = ((h & np.uint64(0xFFFF)) + ((h >> np.uint64(16)) & np.uint64(0xFFFF)) +
         ((h >> np.uint64(32)) & np.uint64(0xFFFF)) + ((h >> np.uint64(48)) & np.uint64(0xFFFF))).astype(np.int64)
    z = (s - 131070).astype(np.float32) * _IH_SCALE
    val = np.rint(c + np.sqrt(c) * z)
    val = np.where(val < 0, np.float32(0), val).astype(np.float32)
    return val.astype(dtype).reshape((lshape0, n1, q2, q3))


def stem4d_device(gshape, offset0=0, lshape0=None, seed=2, counts=500.0, dtype="float32", device=None):
    """Local block [offset0, offset0+lshape0) of the global synthetic array, as a CUDA torch tensor."""
    import torch
    lib = _lib.load()
    _lib.require_gpu()
    n0, n1, q2, q3 = [int(v) for v in gshape]
    lshape0 = n0 - offset0 if lshape0 is None else int(lshape0)
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    tdt = torch.float32 if np.dtype(dtype) == np.float32 else torch.float64
    mod, templ = _stem_tables((n0, n1, q2, q3))
    with torch.cuda.device(dev):
        tmod = torch.from_numpy(mod).to(dev)
        ttem = torch.from_numpy(templ).to(dev)
        out = torch.empty((lshape0, n1, q2, q3), dtype=tdt, device=dev)
        gs = (C.c_int64 * 4)(n0, n1, q2, q3)
        _lib.check(lib.cytvdn_synth_counts(gs, int(offset0), lshape0, 0 if tdt == torch.float32 else 1,
                                           tmod.data_ptr(), ttem.data_ptr(), float(counts), int(seed),
                                           out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
        torch.cuda.current_stream(dev).synchronize()
    return out
