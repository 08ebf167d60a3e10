"""Regenerates tests/golden/mpm88_ref.npz by RUNNING the reference's 88-line program
(/root/reference/mls-mpm88.cpp:16-69, compiled where it lies by `make -C oracle ref` against the
stand-in header oracle/taichi_stub/taichi.h): seeded particle states in, states after 1 / 20 steps out,
elastic and plastic.  Run in the build container:

    python tests/golden/make_mpm88_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyoracle as O  # noqa: E402


STRIDE = 10  # the golden file keeps every 10th particle (and the whole grid): small fixture, full-size run


def golden_inputs(per=1000, seed=7):
    """Three squares of 1000 particles as the program seeds them (mls-mpm88.cpp:70-77; fewer particles
    per cell than that is not a stable discretisation), with a stirred state so that stress, polar
    decomposition, SVD and the snow clamp are all active from the first step."""
    rng = np.random.default_rng(seed)
    n = 3 * per
    x = np.concatenate([(rng.random((per, 2)) * 2 - 1) * 0.08 + np.array(c) for c in ((0.55, 0.45), (0.45, 0.65), (0.55, 0.85))])
    v = rng.normal(size=(n, 2)) * 0.5
    F = np.tile(np.array([1.0, 0, 0, 1.0]), (n, 1)) + rng.normal(size=(n, 4)) * 0.03
    C = rng.normal(size=(n, 4)) * 2.0
    Jp = 1 + rng.normal(size=n) * 0.03
    return tuple(a.astype(np.float32) for a in (x, v, F, C, Jp))


def main():
    x, v, F, C, Jp = golden_inputs()
    out = {}
    for plastic in (0, 1):
        for steps in (1, 20):
            r = O.ref88_run(x, v, F, C, Jp, steps, plastic)
            for name, a in zip(("x", "v", "F", "C", "Jp", "grid"), r):
                out["p%d_s%d_%s" % (plastic, steps, name)] = a if name == "grid" else a[::STRIDE]
    path = os.path.join(HERE, "mpm88_ref.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
