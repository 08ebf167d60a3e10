"""The drop-in boundary, executed: the reference's OWN MPM<3> object
(src/mpm.cpp, transfer.cpp, particles.cpp compiled in place, oracle/transfer_ref.cpp) steps a scene twice —
once by itself (MPM<3>::substep) and once by handing its AoS particle pool to libmpmb through the C-ABI: mpmb_create
from the solver's fields, mpmb_set_material from the particle's parameters, mpmb_upload_aos with the slot layout taken by
offsetof on the reference's own classes, mpmb_substep, mpmb_download_aos back into the pool and the `particles` index
vector.  Afterwards both objects go on with the reference's own substep() and write_partio.  Here the library is the
SIMT-emulator build of the CUDA source (tests/simt); tests/test_gpu_zz_reference_golden.py does the same on the B200."""
import numpy as np
import pytest

from oracle import pyoracle as O
from taichi_mpm_b200 import scenes
from tests import common as T

pytestmark = pytest.mark.skipif(not O.ref_transfer_available(), reason="reference build (oracle/_ref) not available")


def run_dropin(lib_path, kind, tmp_path, nsub=12):
    from tests import common as T
    scene, st = T.perturbed_scene(kind, res=32, cells=5, seed=9 + kind)
    st["x"][0] = [6.5 / 32, 0.5, 0.5]                       # one particle in the deletion band
    a, b = O.RefSolver(scene, st), O.RefSolver(scene, st)
    na = a.substep(nsub)                                    # the reference steps itself
    nb = b.substep_via_mpmb(lib_path, nsub)                 # the reference steps through the C-ABI
    pa, pb = a.particles(), b.particles()
    ids = pa["alive_ids"]
    assert na == nb == len(st["x"]) - 1 and np.array_equal(ids, pb["alive_ids"])
    assert np.abs(pa["x"][ids] - pb["x"][ids]).max() <= 2e-6
    assert np.abs(pa["v"][ids] - pb["v"][ids]).max() <= 2e-4 * np.abs(pa["v"][ids]).max()
    assert np.abs(pa["b"][ids] - pb["b"][ids]).max() <= 5e-4 * np.abs(pa["b"][ids]).max()
    if kind != scenes.MAT_WATER:
        assert np.abs(pa["F"][ids] - pb["F"][ids]).max() <= 5e-5
    assert T.ps_err(pa["ps"][ids], pb["ps"][ids]) <= (1e-3 if kind == scenes.MAT_VISCO else 5e-5)
    # the pool is a valid reference state again: both go on with MPM<3>::substep and dump a frame with write_partio
    assert a.substep(3) == b.substep(3) == na
    qa, qb = a.particles(), b.particles()
    assert np.abs(qa["x"][ids] - qb["x"][ids]).max() <= 3e-6
    fa, fb = tmp_path / "a.bgeo", tmp_path / "b.bgeo"
    a.write_partio(fa); b.write_partio(fb)
    a.close(); b.close()
    from taichi_mpm_b200 import bgeo
    xa, aa = bgeo.read_bgeo(str(fa))
    xb, ab = bgeo.read_bgeo(str(fb))
    assert np.array_equal(dict((n, v) for n, _, v in aa)["index"], dict((n, v) for n, _, v in ab)["index"])
    assert np.abs(xa - xb).max() <= 3e-6


@pytest.mark.parametrize("kind", [scenes.MAT_SAND, scenes.MAT_SNOW, scenes.MAT_WATER, scenes.MAT_JELLY, scenes.MAT_LINEAR,
                                  scenes.MAT_ELASTIC, scenes.MAT_VON_MISES, scenes.MAT_VISCO])
def test_reference_solver_steps_through_the_c_abi_on_the_emulator(kind, tmp_path):
    from tests.simt import build_simt
    run_dropin(build_simt.build(), kind, tmp_path)


def run_dropin_rigid(lib_path, variant, nsub=4):
    """The same with rigid bodies: body A steps with the reference's own coupled substep
    (rasterize_rigid_boundary, gather_cdf, block_op_rigid), body B hands pool, boundary samples, particle colours and rigid state
    to libmpmb through the C-ABI every substep and takes velocities, colours and the reconstructed boundary back."""
    import importlib.util
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_rigid_golden", os.path.join(here, "golden", "make_rigid_golden.py"))
    G = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(G)
    scene, st, rigid = G.golden_scene(variant)
    a, b = O.RefSolver(scene, st), O.RefSolver(scene, st)
    for s in (a, b):
        s.set_rigid(rigid)
        if st.get("states") is not None:
            s.set_states(st["states"])
    a.substep(nsub)                                         # MPM<3>::substep() itself, has_rigid_body() true
    b.substep_via_mpmb(lib_path, nsub)
    pa, pb, ca, cb = a.particles(), b.particles(), a.cdf_particles(), b.cdf_particles()
    ids = pa["alive_ids"]
    assert len(ids) == len(st["x"]) and np.array_equal(ids, pb["alive_ids"])
    assert np.abs(pa["x"] - pb["x"]).max() <= 2e-6
    assert np.abs(pa["v"] - pb["v"]).max() <= 5e-4 * np.abs(pa["v"]).max()
    assert np.abs(pa["F"] - pb["F"]).max() <= 5e-5
    same = ca["states"] == cb["states"]
    assert same.mean() > 0.998                              # a colour decided by two nearly equal weighted distances may differ in fp32
    assert np.array_equal(ca["near"][same], cb["near"][same]) and np.abs(ca["bdist"] - cb["bdist"])[same].max() <= 1e-4 * scene["dx"] * 32
    ra, rb = a.rigid_state(), b.rigid_state()
    for k in range(1, len(rigid["inv_mass"])):
        dv = ra["velocity"][k] - rigid["velocity"][k]
        assert np.abs(ra["velocity"][k] - rb["velocity"][k]).max() <= 5e-3 * np.abs(dv).max() + 1e-6
        assert (np.abs(dv).max() > 1e-3) == (rigid["inv_mass"][k] > 0)
    # the pool is a valid reference state again (RigidBoundaryParticles back in the index vector): both carry on by themselves
    a.substep(2); b.substep(2)
    qa, qb = a.particles(), b.particles()
    assert np.abs(qa["x"] - qb["x"]).max() <= 4e-6
    a.close(); b.close()


@pytest.mark.parametrize("variant", ["dynamic", "two_bodies", "sand_preset"])
def test_reference_solver_with_rigid_bodies_steps_through_the_c_abi_on_the_emulator(variant):
    from tests.simt import build_simt
    run_dropin_rigid(build_simt.build(), variant)


def run_async(lib_path, steps=2, kind=scenes.MAT_SNOW):
    """SURVEY §8f row 3 through the drop-in: the reference's own AsyncMPM<3> object — its scheduler (per-block power-of-two time
    levels, backup pools, update_dt_limits / advance / step, src/async/async_mpm.cpp) compiled in place — twice: A with the
    reference's MPM<3>::substep(), B with every substep the scheduler asks for handed to ONE libmpmb engine through the C-ABI
    (mpmb_set_delta_t for the level's step, mpmb_upload_aos of the level's particle set, one substep, mpmb_download_aos)."""
    from tests import common as T
    scene, st = T.perturbed_scene(kind, res=32, cells=8, seed=7, strain=0.0, vel=0.0, with_floor=(kind == scenes.MAT_SAND))
    st["v"][:] = 0
    st["v"][st["x"][:, 0] > 0.55, 0] = 10.0          # a fast half: with cfl_dt_mul = 0.1 its blocks step at 8 units, the rest at 32
    unit = 2.5e-5 if kind == scenes.MAT_SNOW else 5e-6   # sand is stiffer: its strength limit (get_allowed_dt, src/particles.cpp:649) is smaller
    kw = dict(unit_delta_t=unit, max_units=64, cfl_dt_mul=0.1)
    a, b = O.RefAsyncSolver(scene, st, **kw), O.RefAsyncSolver(scene, st, **kw)
    b.route_through(lib_path)
    for _ in range(steps):
        ra, rb = a.step(80 * unit), b.step(80 * unit)
        # the scheduler made the same decisions on both sides
        assert {k: ra[k] for k in ("alive", "update_counter", "current_t_int", "min_level", "max_level")} == \
               {k: rb[k] for k in ("alive", "update_counter", "current_t_int", "min_level", "max_level")}
    assert ra["min_level"] < ra["max_level"] and rb["routed_substeps"] > 5 and ra["routed_substeps"] == 0   # really asynchronous, really routed
    # (in a scene this small — half the block fast, the other half its neighbour — the scheduler's copies of neighbouring blocks cost
    # more updates than the coarse levels save; the point here is who executes the substeps, not the saving)
    pa, pb = a.particles(), b.particles()
    m = pa["alive"].astype(bool)
    assert np.array_equal(pa["alive"], pb["alive"]) and (~m).sum() < 50
    assert np.abs(pa["x"] - pb["x"])[m].max() <= 2e-6
    assert np.abs(pa["v"] - pb["v"])[m].max() <= 2e-5 * np.abs(pa["v"]).max()
    assert np.abs(pa["F"] - pb["F"])[m].max() <= 5e-5 and np.abs(pa["ps"] - pb["ps"])[m].max() <= 5e-5
    a.close(); b.close()


@pytest.mark.parametrize("kind", [scenes.MAT_SNOW, scenes.MAT_SAND])
def test_reference_asyncmpm_scheduler_runs_on_libmpmb_on_the_emulator(kind):
    from tests.simt import build_simt
    run_async(build_simt.build(), kind=kind)
