"""BC_mode 3 -- the well-defined mirror boundary.  NOT reference behaviour: the reference's
mirror is defined in half-step A only (`anisotropic.pyx:69-70`, backward neighbour of index 0 is index 1); in half-step
B it takes the forward index as ``max(i+1, N-1)`` (`utils.pyx:117-120`), which reads out of bounds.  The spec here is
the evident intent, ``min(i+1, N-1)``.  CPU part: the oracle's restatement of that spec against an independent NumPy
formula, and its half-step A against the compiled reference's own BC_mode=1.  GPU part: the CUDA path against the
oracle, bit for bit."""
import os

import numpy as np
import pytest

from oracle import tv_oracle as O


def counts(rng, shape, dtype):
    mean = rng.uniform(20, 400, size=shape)
    for ax in range(len(shape)):
        mean = 0.5 * (mean + np.roll(mean, 1, axis=ax))
    return rng.poisson(mean).astype(dtype)


def numpy_half_step_b(f, bs, w):
    """recon = f - (((w0 (b0 - b0[clamped i+1]) + w1 (...)) + w2 (...)) + w3 (...)), every operation rounded in the
    array dtype, same association as utils.pyx:111-116."""
    s = None
    for ax, b in enumerate(bs):
        idx = np.minimum(np.arange(b.shape[ax]) + 1, b.shape[ax] - 1)
        term = w[ax] * (b - np.take(b, idx, axis=ax))
        s = term if s is None else s + term
    return f - s


@pytest.mark.parametrize("dt", ["float32", "float64"])
@pytest.mark.parametrize("shape", [(5, 4, 6, 7), (2, 2, 2, 2), (3, 9, 11), (2, 5, 2)])
def test_oracle_mirror_half_step_b_is_the_clamped_index_formula(dt, shape):
    rng = np.random.default_rng(len(shape) * 100 + shape[0])
    f = counts(rng, shape, dt)
    recon = counts(rng, shape, dt)
    bs = [rng.normal(0, 30, shape).astype(dt) for _ in shape]
    w = np.array([1 / 32, 1 / 32, 1 / 16, 1 / 16][:len(shape)], dtype=dt)
    K = O.PortKernels("D")
    want = numpy_half_step_b(f, bs, w)
    old = recon.copy()
    num, den = K.datacube_update_sums(f, recon, bs, w, 3)
    assert np.array_equal(recon, want)
    assert num == pytest.approx(float(np.abs((want - old).astype(np.float64)).sum()), rel=1e-12)
    assert den == pytest.approx(float(np.abs(old.astype(np.float64)).sum()), rel=1e-12)
    # the term of an axis vanishes at its last index: with b constant along every axis nothing changes
    const_b = [np.full(shape, 3.5, dtype=dt) for _ in shape]
    r2 = recon.copy()
    K.datacube_update_sums(f, r2, const_b, w, 3)
    assert np.array_equal(r2, f)


@pytest.mark.parametrize("dt", ["float32", "float64"])
@pytest.mark.parametrize("fista", [False, True])
def test_mirror_half_step_a_is_the_references_bc_mode_1(dt, fista):
    """Half-step A of BC_mode 3 == the reference's BC_mode=1, which is well defined (anisotropic.pyx:69-70)."""
    rng = np.random.default_rng(5)
    for shape in [(4, 5, 6, 7), (6, 5, 9)]:
        a = counts(rng, shape, dt)
        kernels = [O.PortKernels("D")]
        if O.reference_available():
            kernels.append(O.ReferenceKernels("D"))
        for ax in range(len(shape)):
            outs = []
            for K in kernels:
                for mode in (1, 3):
                    b = np.zeros(shape, dt) + a[::-1].reshape(shape) * 0.25
                    d = (b * 0.5).astype(dt) if fista else None
                    n = K.accumulator_update(a, b, d, 0.4, ax, 20.0, mode)
                    outs.append((b, d, n))
            for b, d, n in outs[1:]:
                assert np.array_equal(b, outs[0][0])
                assert d is None or np.array_equal(d, outs[0][1])
                assert n == pytest.approx(outs[0][2], rel=1e-12)
            # index 0 differs from its neighbour at index 1, not from itself
            b = np.zeros(shape, dt)
            O.PortKernels("D").accumulator_update(a, b, None, 0.0, ax, 1e9, 3)
            lo = [slice(None)] * len(shape)
            nx = list(lo)
            lo[ax], nx[ax] = 0, 1
            assert np.array_equal(b[tuple(lo)], a[tuple(lo)] - a[tuple(nx)])


def test_mirror_loop_on_the_oracle():
    rng = np.random.default_rng(8)
    x = counts(rng, (6, 5, 8, 7), "float32")
    mu = np.array([1, 1, .5, .5], dtype=np.float32)
    K = O.PortKernels("D")
    r3 = O.denoise4D(x, mu, 30, True, BC_mode=3, quiet=True, kernels=K, scalars="D")
    r2 = O.denoise4D(x, mu, 30, True, BC_mode=2, quiet=True, kernels=K, scalars="D")
    assert np.all(np.isfinite(r3[0])) and np.all(np.isfinite(r3[2]))
    assert not np.array_equal(r3[0], r2[0])
    assert r3[2][-1] < r3[2][0]                                   # the relative change decays
    flat = np.full((4, 4, 4, 4), 7.0, dtype=np.float32)
    rf = O.denoise4D(flat, mu, 5, True, BC_mode=3, quiet=True, kernels=K, scalars="D")
    assert np.array_equal(rf[0], flat) and np.all(rf[2] == 0)
    with pytest.raises(ValueError, match="undefined behaviour"):
        O.denoise4D(x, mu, 2, True, BC_mode=1, quiet=True, kernels=K, scalars="D")
    if O.reference_available():
        with pytest.raises(ValueError, match="BC_mode 0 and 2 only"):
            O.denoise4D(x, mu, 2, True, BC_mode=3, quiet=True, kernels=O.ReferenceKernels("D"), scalars="D")


# ------------------------------------------------------------------------------------------------
# GPU
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def tv():
    import cytvdn_b200 as tv
    if tv.device_count() < 1:
        pytest.fail("no CUDA device visible: the gpu-marked tests must run on a GPU box")
    return tv


@pytest.mark.gpu
@pytest.mark.parametrize("shape,dt,iters,fista", [
    ((7, 6, 8, 16), "float32", 25, True),
    ((5, 4, 6, 13), "float32", [6, 5], True),       # odd rows (padded internally)
    ((4, 3, 5, 6), "float64", 20, False),
    ((9, 7, 24), "float32", 30, True),              # 3-D
    ((2, 2, 2, 2), "float64", 10, True),            # smallest legal extents
    ((3, 40, 5, 37), "float32", 12, True),
])
@pytest.mark.parametrize("pad", ["1", "0"])
def test_gpu_mirror_denoise_equals_oracle(tv, shape, dt, iters, fista, pad, monkeypatch):
    monkeypatch.setenv("CYTVDN_PAD_ROWS", pad)      # 0: odd rows run on the 8-byte / scalar kernels
    O.set_threads(O.max_threads())
    rng = np.random.default_rng(sum(shape))
    x = counts(rng, shape, dt)
    mu = np.array([1, 1, .5, .5] if len(shape) == 4 else [1, 1, .5], dtype=dt)
    fn, ofn = (tv.denoise4D, O.denoise4D) if len(shape) == 4 else (tv.denoise3D, O.denoise3D)
    want = ofn(x, mu, iters, FISTA=fista, BC_mode=3, quiet=True, kernels=O.PortKernels("D"), scalars="D")
    vec_rows = pad == "1" or shape[-1] % (4 if dt == "float32" else 2) == 0
    for sched in (None, "two_pass"):
        tm = {}
        got = fn(x, mu, iters, FISTA=fista, BC_mode=3, quiet=True, timing=tm, schedule=sched)
        # auto: the fused kernel's mirror variant wherever rows are 16-byte aligned (padded internally by default)
        assert tm["schedule"] == ("fused" if (sched is None and vec_rows) else "two_pass"), (tm, sched)
        assert np.array_equal(got[0], want[0]), (sched, float(np.abs(got[0] - want[0]).max()))
        np.testing.assert_allclose(got[1].astype(np.float64), want[1], rtol=1e-6)
        np.testing.assert_allclose(got[2].astype(np.float64), want[2], rtol=1e-6)


@pytest.mark.gpu
def test_gpu_mirror_step_function_and_errors(tv):
    rng = np.random.default_rng(3)
    shape = (4, 5, 6, 9)
    f = counts(rng, shape, "float32")
    recon = counts(rng, shape, "float32")
    bs = [rng.normal(0, 30, shape).astype(np.float32) for _ in range(4)]
    w = np.array([1 / 32, 1 / 32, 1 / 16, 1 / 16], dtype=np.float32)
    want = numpy_half_step_b(f, bs, w)
    old = recon.copy()
    r = tv.datacube_update_4D(f, recon, *bs, w, 3)
    assert np.array_equal(recon, want)
    assert r == pytest.approx(float(np.abs((want - old).astype(np.float64)).sum() / np.abs(old.astype(np.float64)).sum()),
                              rel=1e-9)
    mu = np.array([1, 1, .5, .5], dtype=np.float32)
    a = tv.denoise4D(f, mu, 3, True, BC_mode=3, quiet=True, schedule="fused")      # round 2: the fused kernel knows the mirror
    b = tv.denoise4D(f, mu, 3, True, BC_mode=3, quiet=True, schedule="two_pass")
    assert np.array_equal(a[0], b[0])
    with pytest.raises(Exception, match="anisotropic update only"):
        tv.denoise4D(f, mu, 3, True, isotropic_Q=True, BC_mode=3, quiet=True)
    with pytest.raises(NotImplementedError, match="BC_mode=3"):
        tv.denoise4D(f, mu, 3, True, BC_mode=1, quiet=True)


@pytest.mark.gpu
def test_gpu_mirror_pipelined_and_streamed(tv, monkeypatch):
    """BC_mode=3 outside the plain loop (round-1 review): PCIe pipeline and out-of-core tiles give the in-core
    single-GPU result bit for bit."""
    rng = np.random.default_rng(8)
    x = counts(rng, (24, 6, 8, 16), "float32")
    mu = np.array([1, 1, .5, .5], dtype=np.float32)
    ref = tv.denoise4D(x, mu, [9, 4], True, BC_mode=3, quiet=True, schedule="two_pass")
    monkeypatch.setenv("CYTVDN_PIPELINE", "5")
    tm = {}
    got = tv.denoise4D(x, mu, [9, 4], True, BC_mode=3, quiet=True, timing=tm)
    assert tm["schedule"] == "fused" and tm["pipeline_boxes"] == 5
    assert np.array_equal(got[0], ref[0])
    monkeypatch.delenv("CYTVDN_PIPELINE")
    plane = 6 * 8 * 16 * 4
    monkeypatch.setenv("CYTVDN_STREAM_BUDGET_MB", str(10 * 2.5 * 12 * plane / 1048576.0))
    tm = {}
    got = tv.denoise4D(x, mu, [9, 4], True, BC_mode=3, quiet=True, timing=tm)
    assert tm["schedule"] == "streamed" and tm["stream_tiles"] > 1
    assert np.array_equal(got[0], ref[0])
    # (axis-0 shards with BC_mode=3: tests/engine_driver.py, run by tests/test_shard_engine.py in its own process)
