"""Regenerates tests/golden/transfer_ref.npz by RUNNING the reference's solver core
(/root/reference/src/transfer.cpp: rasterize_optimized / resample_optimized — the SSE fast path the
hot path uses — and the scalar rasterize / resample; /root/reference/src/mpm.cpp: grid normalisation,
level-set boundary condition, ordering, boundary deletion, and MPM<3>::substep() itself), compiled where
they lie by `make -C oracle ref` (oracle/transfer_ref.cpp, stand-in core headers
oracle/taichi_stub/taichi/*.h, vendored SPGrid).
Run in the build container:

    python tests/golden/make_transfer_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyoracle as O  # noqa: E402
from tests import common as T  # noqa: E402

RES, CELLS = 24, 3   # 216 particles per case
KINDS = [0, 1, 2, 3, 4, 5, 6, 7]


SUBSTEPS = 10


def golden_scene(kind):
    """The stirred block the kernel parity tests use (random affine velocity field, apic_b, F), with the floor."""
    seed = {0: 20, 1: 41, 2: 22, 3: 23, 4: 24, 5: 20, 6: 41, 7: 22}[kind]   # seeds whose velocity field drives nodes into the floor: the boundary condition acts
    return T.perturbed_scene(kind, res=RES, cells=CELLS, seed=seed, **T.KIND_KW.get(kind, {}))


def substep_scene(kind):
    """The same block with two particles moved into the 7-cell deletion band (src/mpm.h:269-276)."""
    scene, st = golden_scene(kind)
    st["x"][0] = [6.5 / RES, 0.5, 0.5]
    st["x"][1] = [0.5, 0.5, (RES - 6.5) / RES]
    return scene, st


def oracle_grid_vel(scene, st):
    """Node velocities after the grid update as the oracle computes them: the INPUT handed to the reference's G2P
    (the grid update itself lives in src/mpm.cpp, which is not part of the reference build here)."""
    _, _, grid_vel = O.substep(scene, st, np.float64)
    return np.ascontiguousarray(grid_vel, np.float32)


def sparse(grid):
    """(flat node indices, values) of the nodes that hold anything: keeps the fixture small."""
    flat = grid.reshape(-1, 4)
    idx = np.nonzero(np.abs(flat).max(1) > 0)[0].astype(np.int32)
    return idx, flat[idx]


def dense(idx, val):
    g = np.zeros(((RES + 1) ** 3, 4), np.float32)
    g[idx] = val
    return g.reshape(RES + 1, RES + 1, RES + 1, 4)


def main():
    out = {}
    for kind in KINDS:
        scene, st = golden_scene(kind)
        gv = oracle_grid_vel(scene, st)
        out["k%d_grid_vel_in_idx" % kind], out["k%d_grid_vel_in_val" % kind] = sparse(gv)
        for tag, opt in (("opt", True), ("scalar", False)):
            grid, p = O.ref_transfer_substep(scene, st, gv, optimized=opt)
            out["k%d_%s_grid_idx" % (kind, tag)], out["k%d_%s_grid_val" % (kind, tag)] = sparse(grid)
            for name, a in p.items():
                out["k%d_%s_%s" % (kind, tag, name)] = a
        # the grid update by the reference (normalize + boundary condition) after its own P2G
        s = O.RefSolver(scene, st)
        s.p2g(True)
        s.grid_update()
        out["k%d_gridupd_idx" % kind], out["k%d_gridupd_val" % kind] = sparse(s.get_grid())
        s.close()
        # whole substeps by MPM<3>::substep()
        scene2, st2 = substep_scene(kind)
        s = O.RefSolver(scene2, st2)
        out["k%d_sub_alive" % kind] = np.int32(s.substep(SUBSTEPS))
        p = s.particles()
        s.close()
        for name, a in p.items():
            out["k%d_sub_%s" % (kind, name)] = a
    path = os.path.join(HERE, "transfer_ref.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
