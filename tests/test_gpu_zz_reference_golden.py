"""The CUDA engine against the REFERENCE ITSELF: tests/golden/transfer_ref.npz holds what the reference's own
MPM<3>::substep() (src/mpm.cpp + src/transfer.cpp + src/particles.cpp, compiled in place for the golden run,
oracle/transfer_ref.cpp) produced after 10 substeps of the stirred block with a friction floor and two
particles in the deletion band.  No oracle in between: same inputs through the C-ABI, same survivors, same state."""
import importlib.util
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_transfer_golden", os.path.join(HERE, "golden", "make_transfer_golden.py"))
G = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(G)


@pytest.mark.parametrize("kind", G.KINDS)
def test_engine_matches_reference_substeps(kind):
    from tests import common as T
    from taichi_mpm_b200 import scenes
    z = np.load(os.path.join(HERE, "golden", "transfer_ref.npz"))
    scene, st = G.substep_scene(kind)
    e = T.make_engine(scene, st)
    e.substep(G.SUBSTEPS)
    got = e.download()
    e.close()
    ids = got["id"].astype(np.int64)
    assert np.array_equal(ids, z["k%d_sub_alive_ids" % kind])                     # the same two particles were deleted
    assert len(ids) == int(z["k%d_sub_alive" % kind]) == len(st["x"]) - 2
    ref = {k: z["k%d_sub_%s" % (kind, k)][ids] for k in ("x", "v", "F", "b", "ps")}
    # fp32 on both sides, different summation orders, 10 substeps: ten times the single-substep tolerances
    assert np.abs(got["x"] - ref["x"]).max() <= 1e-5
    assert np.abs(got["v"] - ref["v"]).max() <= 1e-3 * np.abs(ref["v"]).max()
    assert np.abs(got["b"] - ref["b"]).max() <= 2e-3 * np.abs(ref["b"]).max()
    if kind != scenes.MAT_WATER:
        assert np.abs(got["F"] - ref["F"]).max() <= 2e-4
    assert np.abs(got["ps"] - ref["ps"]).max() <= 1e-4
