#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (kernels AND its own Python
driver `cyTVDN/cyTVDN.py`) in the authoring container.

The reference tree (/root/reference) is read-only and does not travel to the GPU box, so the
vectors it produces are committed here together with this script.  How the reference is run:

* kernels: compiled from the reference's .pyx by ``oracle/build_ref.py`` (SURVEY.md §8c recipe);
* driver:  a scratch package is assembled under a temp dir (never inside this repo) holding the
  reference's ``__init__.py`` / ``cyTVDN.py`` next to the compiled kernels, plus a 2-symbol
  ``hurry.filesize`` stub (`cyTVDN.py:13` imports it unconditionally, it is not installed);
* OMP_NUM_THREADS=1 -- the half-isotropic kernels race with more threads
  (`halfisotropic.pyx:45-48,70-82`) and one thread makes the array-dtype scalars deterministic.

Run:  python tests/golden/make_golden.py        (needs /root/reference; ~10 s)
"""
import os
os.environ["OMP_NUM_THREADS"] = "1"

import json
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFERENCE = "/root/reference"


def load_reference():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import build_ref
    print("oracle/_ref:", build_ref.build(REFERENCE))
    tmp = tempfile.mkdtemp(prefix="cytvdn_ref_")
    pkg = os.path.join(tmp, "cyTVDN")
    os.makedirs(pkg)
    for f in os.listdir(build_ref.PKG):
        if f.endswith(".so"):
            shutil.copy(os.path.join(build_ref.PKG, f), pkg)
    for f in ("__init__.py", "cyTVDN.py"):
        shutil.copy(os.path.join(REFERENCE, "cyTVDN", f), pkg)
    os.makedirs(os.path.join(tmp, "hurry"))
    open(os.path.join(tmp, "hurry", "__init__.py"), "w").close()
    with open(os.path.join(tmp, "hurry", "filesize.py"), "w") as f:
        f.write("alternative = None\n\ndef size(n, system=None):\n    return '%d B' % n\n")
    sys.path.insert(0, tmp)
    for m in [m for m in sys.modules if m == "cyTVDN" or m.startswith("cyTVDN.")]:
        del sys.modules[m]
    import cyTVDN as tv
    return tv, tmp


def counts(rng, shape, dtype, lo=20.0, hi=400.0):
    """Count-like data (Poisson around a smooth-ish random mean) so the clip is active."""
    mean = rng.uniform(lo, hi, size=shape)
    for ax in range(len(shape)):   # cheap smoothing so that TV has structure to keep
        mean = 0.5 * (mean + np.roll(mean, 1, axis=ax))
    return rng.poisson(mean).astype(dtype)


DENOISE_CASES = [
    # name, ndim, shape, dtype, mu, kwargs
    ("d4_f32_fista", 4, (6, 5, 8, 12), "float32", [1, 1, .5, .5], dict(iterations=12, FISTA=True)),
    ("d4_f32_plain_odd", 4, (5, 6, 7, 9), "float32", [1, 1, .5, .5], dict(iterations=10, FISTA=False)),
    ("d4_f64_fista", 4, (4, 5, 6, 8), "float64", [1, 1, .5, .5], dict(iterations=15, FISTA=True)),
    ("d4_f32_fista_100", 4, (8, 8, 8, 8), "float32", [1, 1, .5, .5], dict(iterations=100, FISTA=True)),
    ("d4_f64_fista_100", 4, (6, 6, 8, 8), "float64", [1, 1, .5, .5], dict(iterations=100, FISTA=True)),
    ("d4_f32_iso_rq_fista", 4, (6, 5, 8, 8), "float32", [1, 1, .5, .5],
     dict(iterations=10, FISTA=True, isotropic_R=True, isotropic_Q=True)),
    ("d4_f32_iso_r_plain", 4, (5, 7, 6, 4), "float32", [1, 1, .5, .5],
     dict(iterations=8, FISTA=False, isotropic_R=True)),
    ("d4_f64_iso_q_fista", 4, (4, 4, 7, 6), "float64", [2, 1, .5, .25],
     dict(iterations=9, FISTA=True, isotropic_Q=True)),
    ("d4_f32_hybrid_mse", 4, (5, 4, 6, 8), "float32", [1, 1, .5, .5],
     dict(iterations=[4, 3], reference_data=True)),
    ("d4_f32_bc0_fista", 4, (4, 5, 6, 8), "float32", [1, 1, .5, .5], dict(iterations=8, FISTA=True, BC_mode=0)),
    ("d4_f32_stop", 4, (6, 6, 8, 8), "float32", [1, 1, .5, .5],
     dict(iterations=40, FISTA=True, stopping_relative_change=0.02)),
    ("d4_f32_lam", 4, (4, 4, 4, 8), "float32", [1, 2, .5, .5], dict(iterations=6, FISTA=True, lam=[.02, .03, .01, .015])),
    ("d4_f32_degenerate", 4, (3, 1, 1, 4), "float32", [1, 1, .5, .5], dict(iterations=5, FISTA=True)),
    ("d4_f32_unit", 4, (1, 1, 1, 1), "float32", [1, 1, .5, .5], dict(iterations=3, FISTA=False)),
    ("d3_f32_plain", 3, (7, 6, 32), "float32", [1, 1, .5], dict(iterations=20, FISTA=False)),
    ("d3_f32_fista_stop", 3, (8, 8, 64), "float32", [1, 1, .5],
     dict(iterations=30, FISTA=True, stopping_relative_change=0.03)),
    ("d3_f64_fista_bc0", 3, (5, 6, 17), "float64", [1, 1, .5], dict(iterations=12, FISTA=True, BC_mode=0)),
    ("d3_f32_hybrid_mse", 3, (6, 5, 20), "float32", [1, .5, .5], dict(iterations=(3, 4), reference_data=True)),
    ("d3_f32_degenerate", 3, (5, 1, 7), "float32", [1, 1, .5], dict(iterations=6, FISTA=True)),
    ("d3_f32_fista_100", 3, (8, 8, 32), "float32", [1, 1, .5], dict(iterations=100, FISTA=True)),
]


def main():
    tv, tmp = load_reference()
    rng = np.random.default_rng(20261018)
    index = {}

    # ---- driver-level vectors: tv.denoise3D / tv.denoise4D of the reference itself -------------
    for name, ndim, shape, dt, mu, kw in DENOISE_CASES:
        dt = np.dtype(dt)
        data = counts(rng, shape, dt)
        mu_a = np.array(mu, dtype=dt)
        kwargs = dict(kw)
        save = dict(data=data, mu=mu_a)
        if kwargs.get("reference_data") is True:
            ref_data = counts(rng, shape, dt)
            kwargs["reference_data"] = ref_data
            save["reference_data"] = ref_data
        if "lam" in kwargs:
            kwargs["lam"] = np.array(kwargs["lam"], dtype=dt)
            save["lam"] = kwargs["lam"]
        keep = data.copy()
        fn = tv.denoise4D if ndim == 4 else tv.denoise3D
        out = fn(data, mu_a, quiet=True, **kwargs)
        assert np.array_equal(keep, data), "reference modified its input"
        save.update(recon=out[0], b_norm=out[1], delta_recon=out[2])
        if len(out) == 4:
            save["MSE"] = out[3]
        jkw = {k: (v if not isinstance(v, np.ndarray) else "<array>") for k, v in kw.items()}
        save["kwargs_json"] = np.array(json.dumps(jkw))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **save)
        index[name] = dict(kind="denoise", ndim=ndim, shape=list(shape), dtype=str(dt), kwargs=jkw,
                           iters_nonzero=int(np.count_nonzero(out[2])))
        print(f"{name:24s} range={float(data.max()-data.min()):7.1f} max|recon-data|="
              f"{float(np.abs(out[0]-data).max()):8.3f} delta[-1]={float(out[2][-1]):.3e}")

    # ---- kernel-level vectors: every step function on arbitrary (non-invariant) state ----------
    steps = {}
    for dt in ("float32", "float64"):
        t = np.dtype(dt).type
        for shape in ((4, 3, 5, 8), (3, 4, 6, 7), (2, 1, 3, 5), (5, 6, 12), (4, 3, 9), (1, 4, 6)):
            nd = len(shape)
            a = counts(rng, shape, dt)
            key0 = f"{dt}_{'x'.join(map(str, shape))}"
            for ax in range(nd):
                for bc in (0, 1, 2):
                    if bc == 1 and shape[ax] < 2:
                        continue
                    b = rng.normal(0, 20, shape).astype(dt)
                    d = rng.normal(0, 20, shape).astype(dt)
                    clip = t(32.0 / (1 + ax % 2))
                    tk = t(0.37)
                    k = f"acc_{key0}_ax{ax}_bc{bc}"
                    b1 = b.copy()
                    f_plain = tv.accumulator_update_4D if nd == 4 else tv.accumulator_update_3D
                    n1 = f_plain(a, b1, ax, clip, bc)
                    b2, d2 = b.copy(), d.copy()
                    f_fista = tv.accumulator_update_4D_FISTA if nd == 4 else tv.accumulator_update_3D_FISTA
                    n2 = f_fista(a, b2, d2, tk, ax, clip, bc)
                    steps[k + "_in"] = np.stack([a, b, d])
                    steps[k + "_plain"] = b1
                    steps[k + "_fista"] = np.stack([b2, d2])
                    steps[k + "_norms"] = np.array([n1, n2, clip, tk], dtype=np.float64)
            # reconstruction update on arbitrary b (plane 0 of b non-zero: the wrap term is read)
            bs = [rng.normal(0, 30, shape).astype(dt) for _ in range(nd)]
            w = np.array([1 / 32., 1 / 32., 1 / 64., 1 / 48.][:nd], dtype=dt)
            for bc in (0, 2):
                u = a.copy() + rng.normal(0, 5, shape).astype(dt)
                u_in = u.copy()
                f_dcu = tv.datacube_update_4D if nd == 4 else tv.datacube_update_3D
                r = f_dcu(a, u, *bs, w, bc)
                k = f"dcu_{key0}_bc{bc}"
                steps[k + "_in"] = np.stack([a, u_in] + bs)
                steps[k + "_w"] = w
                steps[k + "_out"] = u
                steps[k + "_ratio"] = np.array([r], dtype=np.float64)
            # sum of squared error
            f_sse = tv.sum_square_error_4D if nd == 4 else tv.sum_square_error_3D
            steps[f"sse_{key0}"] = np.array([f_sse(a, bs[0])], dtype=np.float64)
            steps[f"sse_{key0}_in"] = np.stack([a, bs[0]])
            # half-isotropic (4-D only), a few axis pairs incl. non-adjacent
            if nd == 4:
                for (p, q) in ((0, 1), (2, 3), (1, 3), (3, 0)):
                    b1 = rng.normal(0, 25, shape).astype(dt); b2 = rng.normal(0, 25, shape).astype(dt)
                    d1 = rng.normal(0, 25, shape).astype(dt); d2 = rng.normal(0, 25, shape).astype(dt)
                    clip, tk = t(32.0), t(0.6)
                    k = f"iso_{key0}_p{p}q{q}"
                    steps[k + "_in"] = np.stack([a, b1, b2, d1, d2])
                    x1, x2 = b1.copy(), b2.copy()
                    n1 = tv.iso_accumulator_update_4D(a, x1, x2, p, q, clip)
                    steps[k + "_plain"] = np.stack([x1, x2])
                    y1, y2, e1, e2 = b1.copy(), b2.copy(), d1.copy(), d2.copy()
                    n2 = tv.iso_accumulator_update_4D_FISTA(a, y1, y2, e1, e2, tk, p, q, clip)
                    steps[k + "_fista"] = np.stack([y1, y2, e1, e2])
                    steps[k + "_norms"] = np.array([n1, n2, clip, tk], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "steps.npz"), **steps)
    index["steps"] = dict(kind="steps", n_arrays=len(steps))

    with open(os.path.join(HERE, "index.json"), "w") as f:
        json.dump(index, f, indent=1, sort_keys=True)
    shutil.rmtree(tmp, ignore_errors=True)
    tot = sum(os.path.getsize(os.path.join(HERE, f)) for f in os.listdir(HERE) if f.endswith(".npz"))
    print(f"wrote {len(index)} golden files, {tot/1024:.0f} KiB")


if __name__ == "__main__":
    main()
