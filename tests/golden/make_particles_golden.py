"""Regenerates tests/golden/particles_ref.npz by RUNNING the reference's constitutive models
(/root/reference/src/particles.cpp: <Type>Particle<3>::plasticity and ::calculate_force, and
friction_project of src/mpm_fwd.h:25-57), compiled where they lie by `make -C oracle ref` against the
stand-in core headers oracle/taichi_stub/taichi/*.h.  Run in the build container:

    python tests/golden/make_particles_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyoracle as O  # noqa: E402

# (kind, |F - I| scale, |cdg - I| scale): the regimes the parity tests of the kernels use
CASES = [(0, 0.05, 0.01), (1, 0.1, 0.01), (2, 0.03, 0.01), (3, 0.0, 0.01), (4, 2e-3, 1e-3), (4, 0.05, 0.02),
         (5, 0.05, 0.01), (6, 0.02, 0.01), (6, 1e-3, 1e-3), (7, 0.05, 0.01)]
VOL = 1e-6


def golden_states(kind, strain, rate, count=40, seed=0):
    """States with det F > 0 (inverted elements depend on the SVD's sign convention)."""
    rng = np.random.default_rng(1000 * kind + seed + int(strain * 1e4))
    out = []
    while len(out) < count:
        F = (np.eye(3) + rng.normal(size=(3, 3)) * strain).astype(np.float32)
        cdg = (np.eye(3) + rng.normal(size=(3, 3)) * rate).astype(np.float32)
        if np.linalg.det(F.astype(np.float64)) <= 0.2 or np.linalg.det(cdg.astype(np.float64)) <= 0.2:
            continue
        ps = {2: 1 + rng.normal() * 0.05, 3: 1 + rng.normal() * 0.02, 4: abs(rng.normal()) * 2e-3 * (rng.random() < 0.5),
              7: 1000.0 * (0.5 + rng.random())}.get(kind, 0.0)   # 7: visco_tau around its default
        out.append((F, cdg, np.float32(ps)))
    return out


def friction_cases(seed=9, count=60):
    rng = np.random.default_rng(seed)
    v = rng.normal(size=(count, 3)).astype(np.float32)
    base = (rng.normal(size=(count, 3)) * 0.2).astype(np.float32)
    n = rng.normal(size=(count, 3))
    n = (n / np.linalg.norm(n, axis=1, keepdims=True)).astype(np.float32)
    fr = np.array([-1.0, -2.0, -2.3, 0.0, 0.4, 5.0], np.float32)[rng.integers(0, 6, size=count)]
    return v, base, n, fr


def main():
    out = {}
    for ci, (kind, strain, rate) in enumerate(CASES):
        prm = O.ref_default_params(kind)
        out["c%d_params" % ci] = prm
        Fs, pss, fs = [], [], []
        for F, cdg, ps in golden_states(kind, strain, rate):
            F1, ps1, force = O.ref_particle_step(kind, prm, cdg, F, ps, VOL)
            Fs.append(F1); pss.append(ps1); fs.append(force)
        out["c%d_F" % ci], out["c%d_ps" % ci], out["c%d_force" % ci] = np.array(Fs, np.float32), np.array(pss, np.float32), np.array(fs, np.float32)
    v, base, n, fr = friction_cases()
    out["friction_out"] = np.array([O.ref_friction_project(v[i], base[i], n[i], fr[i]) for i in range(len(v))], np.float32)
    path = os.path.join(HERE, "particles_ref.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
