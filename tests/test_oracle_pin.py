"""Pin the CPU oracle (oracle/tv_oracle.c + oracle/tv_oracle.py) before anything trusts it.

1. against the golden vectors produced by the UNMODIFIED reference (kernels + its own Python
   driver) -- tests/golden/*.npz, generator tests/golden/make_golden.py;
2. against the compiled reference kernels of oracle/_ref, when they are present (authoring
   container and GPU box), on fresh random inputs.

Bit-exact for arrays.  The reference's array-dtype scalars are reproduced bit-exactly by the
"T" flavour of the port at one thread (the goldens were made with OMP_NUM_THREADS=1).
"""
import json
import os

import numpy as np
import pytest

from oracle import tv_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
INDEX = json.load(open(os.path.join(GOLDEN, "index.json")))
DENOISE = sorted(k for k, v in INDEX.items() if v["kind"] == "denoise")


@pytest.fixture(scope="module")
def steps():
    return np.load(os.path.join(GOLDEN, "steps.npz"))


@pytest.fixture(autouse=True)
def one_thread():
    O.set_threads(1)
    yield
    O.set_threads(O.max_threads())


def _kwargs(z, meta):
    kw = dict(meta["kwargs"])
    if kw.get("reference_data"):
        kw["reference_data"] = z["reference_data"]
    if "lam" in kw:
        kw["lam"] = z["lam"]
    if isinstance(kw.get("iterations"), list):
        kw["iterations"] = list(kw["iterations"])
    return kw


@pytest.mark.parametrize("name", DENOISE)
def test_port_host_loop_matches_reference_driver(name):
    """oracle host loop + C port == tv.denoise3D/4D of the reference, bit for bit."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = INDEX[name]
    fn = O.denoise4D if meta["ndim"] == 4 else O.denoise3D
    data = z["data"].copy()
    out = fn(data, z["mu"], quiet=True, kernels=O.PortKernels("T"), scalars="T", **_kwargs(z, meta))
    assert np.array_equal(data, z["data"])
    assert out[0].dtype == z["recon"].dtype
    assert np.array_equal(out[0], z["recon"]), f"max diff {np.abs(out[0]-z['recon']).max()}"
    assert np.array_equal(out[1], z["b_norm"])
    assert np.array_equal(out[2], z["delta_recon"], equal_nan=True)
    if "MSE" in z.files:
        assert len(out) == 4 and np.array_equal(out[3], z["MSE"])
    else:
        assert len(out) == 3


@pytest.mark.parametrize("name", DENOISE)
def test_truth_scalars_close_to_reference_scalars(name):
    """float64-accumulated scalars agree with the reference's array-dtype scalars at these tiny
    sizes (<= 4k voxels, where the fp32 accumulation error is still small)."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = INDEX[name]
    fn = O.denoise4D if meta["ndim"] == 4 else O.denoise3D
    kw = _kwargs(z, meta)
    kw.pop("stopping_relative_change", None)  # compare at equal iteration counts
    n = meta["iters_nonzero"]
    out = fn(z["data"].copy(), z["mu"], quiet=True, kernels=O.PortKernels("D"), scalars="D", **kw)
    tol = 2e-5 if z["data"].dtype == np.float32 else 1e-12
    np.testing.assert_allclose(out[1][:n], z["b_norm"][:n].astype(np.float64), rtol=tol)
    np.testing.assert_allclose(out[2][:n], z["delta_recon"][:n].astype(np.float64), rtol=20 * tol, atol=1e-12)


def test_port_step_kernels_match_reference_vectors(steps):
    K = O.PortKernels("T")
    n_checked = 0
    for key in steps.files:
        if key.startswith("acc_") and key.endswith("_in"):
            base = key[:-3]
            a, b, d = steps[key]
            ax = int(base.split("_ax")[1][0])
            bc = int(base.split("_bc")[1][0])
            n1, n2, clip, tk = steps[base + "_norms"]
            b1 = b.copy()
            r1 = K.accumulator_update(a, b1, None, 0.0, ax, clip, bc)
            assert np.array_equal(b1, steps[base + "_plain"]), base
            assert r1 == n1, base
            b2, d2 = b.copy(), d.copy()
            r2 = K.accumulator_update(a, b2, d2, tk, ax, clip, bc)
            assert np.array_equal(np.stack([b2, d2]), steps[base + "_fista"]), base
            assert r2 == n2, base
            n_checked += 1
        elif key.startswith("dcu_") and key.endswith("_in"):
            base = key[:-3]
            arrs = steps[key]
            f, u, bs = arrs[0], arrs[1].copy(), list(arrs[2:])
            bc = int(base.split("_bc")[1][0])
            r = K.datacube_update(f, u, bs, steps[base + "_w"], bc)
            assert np.array_equal(u, steps[base + "_out"]), base
            assert r == steps[base + "_ratio"][0], base
            n_checked += 1
        elif key.startswith("iso_") and key.endswith("_in"):
            base = key[:-3]
            a, b1, b2, d1, d2 = steps[key]
            p = int(base.split("_p")[1][0])
            q = int(base.split("q")[-1])
            n1, n2, clip, tk = steps[base + "_norms"]
            x1, x2 = b1.copy(), b2.copy()
            r1 = K.iso_accumulator_update(a, x1, x2, None, None, 0.0, p, q, clip)
            assert np.array_equal(np.stack([x1, x2]), steps[base + "_plain"]), base
            assert r1 == n1, base
            y1, y2, e1, e2 = b1.copy(), b2.copy(), d1.copy(), d2.copy()
            r2 = K.iso_accumulator_update(a, y1, y2, e1, e2, tk, p, q, clip)
            assert np.array_equal(np.stack([y1, y2, e1, e2]), steps[base + "_fista"]), base
            assert r2 == n2, base
            n_checked += 1
        elif key.startswith("sse_") and not key.endswith("_in"):
            a, b = steps[key + "_in"]
            assert K.sum_square_error(a, b) == steps[key][0], key
            n_checked += 1
    assert n_checked > 100


needs_ref = pytest.mark.skipif(not O.reference_available(), reason="oracle/_ref not built")


@needs_ref
@pytest.mark.parametrize("dt", ["float32", "float64"])
@pytest.mark.parametrize("shape", [(9, 7, 16, 20), (3, 5, 7, 11), (12, 9, 40), (2, 2, 2, 2), (1, 3, 1, 5)])
def test_port_equals_compiled_reference_random(dt, shape):
    """Fresh random state, all threads: arrays bit-equal (scalars differ only by summation order)."""
    O.set_threads(O.max_threads())
    rng = np.random.default_rng(hash((dt,) + shape) % 2**32)
    P, R = O.PortKernels("D"), O.ReferenceKernels("D")
    nd = len(shape)
    a = rng.poisson(rng.uniform(5, 300, shape)).astype(dt)
    for ax in range(nd):
        for bc in (0, 2):
            b = rng.normal(0, 20, shape).astype(dt)
            d = rng.normal(0, 20, shape).astype(dt)
            for fista in (False, True):
                bp, dp, br, dr = b.copy(), d.copy(), b.copy(), d.copy()
                sp = P.accumulator_update(a, bp, dp if fista else None, 0.41, ax, 24.0, bc)
                sr = R.accumulator_update(a, br, dr if fista else None, 0.41, ax, 24.0, bc)
                assert np.array_equal(bp, br) and np.array_equal(dp, dr)
                assert sp == pytest.approx(sr, rel=1e-12)
    bs = [rng.normal(0, 30, shape).astype(dt) for _ in range(nd)]
    w = np.array([1 / 32., 1 / 40., 1 / 64., 1 / 50.][:nd], dtype=dt)
    up = a + rng.normal(0, 3, shape).astype(dt)
    ur = up.copy()
    sp = P.datacube_update_sums(a, up, bs, w)
    sr = R.datacube_update_sums(a, ur, bs, w)
    assert np.array_equal(up, ur)
    assert sp == pytest.approx(sr, rel=1e-12)


@needs_ref
def test_full_loop_port_equals_reference_kernels_4d_fista():
    """30 FISTA iterations on count-like data where the clip is active: same arrays either way."""
    O.set_threads(O.max_threads())
    rng = np.random.default_rng(5)
    data = rng.poisson(rng.uniform(10, 500, (10, 9, 16, 16))).astype(np.float32)
    mu = np.array([1, 1, .5, .5], dtype=np.float32)
    sp, sr = {}, {}
    rp = O.denoise4D(data, mu, 30, True, kernels=O.PortKernels("D"), scalars="D", state_out=sp)
    rr = O.denoise4D(data, mu, 30, True, kernels=O.ReferenceKernels("D"), scalars="D", state_out=sr)
    assert np.array_equal(rp[0], rr[0])
    for k in range(4):
        assert np.array_equal(sp["acc"][k], sr["acc"][k])
        assert float(np.abs(sp["acc"][k]).max()) > 31.9 * (1 + (k >= 2))  # the clip was reached
    np.testing.assert_allclose(rp[1], rr[1], rtol=1e-12)
    np.testing.assert_allclose(rp[2], rr[2], rtol=1e-10)
