"""Run the reference's half-isotropic driver in a fresh process with ONE OpenMP thread and save the result.

The reference's iso kernels share scratch arrays between OpenMP threads (`halfisotropic.pyx:45-48,70-82`) and are
only deterministic single threaded; libgomp reads OMP_NUM_THREADS once at load time, so the
test that compares the GPU with the COMPILED reference (oracle/_ref) spawns this script with OMP_NUM_THREADS=1.

    python tests/ref_one_thread.py in.npz out.npz
in.npz: data, mu, iterations (1 or 2 ints), fista, iso_r, iso_q.   out.npz: recon, bnorm, delta, kernels (name).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np


def main():
    assert os.environ.get("OMP_NUM_THREADS") == "1", "spawn me with OMP_NUM_THREADS=1"
    from oracle import tv_oracle as O
    z = np.load(sys.argv[1])
    it = [int(v) for v in z["iterations"]]
    iters = it[0] if len(it) == 1 else it
    K = O.default_kernels("D")
    r = O.denoise4D(np.ascontiguousarray(z["data"]), z["mu"], iters, bool(z["fista"]), None, bool(z["iso_r"]),
                    bool(z["iso_q"]), quiet=True, kernels=K, scalars="D")
    np.savez(sys.argv[2], recon=r[0], bnorm=r[1], delta=r[2], kernels=K.name)


if __name__ == "__main__":
    main()
