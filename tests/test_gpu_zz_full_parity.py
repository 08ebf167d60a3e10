"""Per-particle parity at the BASELINE sizes and in the regimes the small scenes do not reach:

* one full-size substep of config 2 (128^3, 1.0 M jelly) and config 3 (256^3, 8.0 M sand), every particle perturbed
  (affine velocity field, apic_b, F, plastic scalar), against the fp32 OpenMP port of the reference's optimized path
  (oracle FastOracle: src/transfer.cpp:467-569,837-954 restated, pinned to the reference's own golden runs) and, for
  config 2, against the fp64 oracle as well;
* reduced-scale config 4 (snow) and config 5 (water) against the fp64 oracle;
* sand at strain 0.1 and 0.3, where sand_step_series() declines and every lane takes the Jacobi eigen path ON THE DEVICE
  (src/particles.cpp:599-647).

Tolerances are the single-substep bounds of SURVEY.md §8d (tests/common.py)."""
import numpy as np
import pytest

from taichi_mpm_b200 import scenes
from tests import common as T

pytestmark = pytest.mark.gpu


def _perturbed_config(name, scale, seed, strain, vel):
    cfg = scenes.config(name, scale)
    scene, st = cfg["scene"], cfg["state"]
    T.perturb_state(st, cfg["meta"]["kind"], scene["dx"], seed=seed, strain=strain, vel=vel)
    scene["sdf"] = scenes.planes_sdf(scene["res"], scene["planes"]) if scene["planes"] is not None else None
    return scene, st


def _engine_substep(scene, st):
    e = T.make_engine(dict(scene, sdf=None), st)   # the engine rasterises the planes on the device
    e.substep(1)
    got = e.download()
    c = e.get_counters()
    e.close()
    return got, c


def _errors(got, ref, alive):
    ids = got["id"].astype(np.int64)
    assert len(ids) == int(alive.sum()) and np.array_equal(ids, np.nonzero(alive)[0]), "different particles deleted"
    vmax = np.abs(ref["v"][alive]).max()
    bmax = max(np.abs(ref["b"][alive]).max(), 1e-30)
    return dict(x=np.abs(got["x"] - ref["x"][ids]).max(), v=np.abs(got["v"] - ref["v"][ids]).max() / vmax,
                b=np.abs(got["b"] - ref["b"][ids]).max() / bmax, F=np.abs(got["F"] - ref["F"][ids]).max(),
                ps=np.abs(got["ps"] - ref["ps"][ids]).max())


def _check(err, tag):
    print(tag, {k: float(v) for k, v in err.items()})
    assert err["x"] <= T.TOL_X_ABS and err["v"] <= T.TOL_V_REL and err["b"] <= T.TOL_V_REL, (tag, err)
    assert err["F"] <= T.TOL_F_ABS and err["ps"] <= T.TOL_PS_ABS, (tag, err)


def _port_substep(scene, st):
    from oracle import pyoracle as O
    f = O.FastOracle(scene, st)     # reorder_interval=0: storage index == caller index
    f.substeps(1)
    return f.st, f.st["alive"].astype(bool)


@pytest.mark.needs_cuda
def test_config2_full_size_substep_vs_port_and_fp64_oracle():
    from oracle import pyoracle as O
    scene, st = _perturbed_config("jelly128", 1.0, seed=11, strain=0.02, vel=0.5)
    assert len(st["x"]) == 1_000_000 and tuple(scene["res"]) == (128, 128, 128)
    got, c = _engine_substep(scene, st)
    assert c["alive"] == 1_000_000
    ref, alive = _port_substep(scene, st)
    _check(_errors(got, ref, alive), "cfg2 1.0M jelly vs fp32 port")
    ref64, _, _ = O.substep(scene, st, np.float64, want_grids=False)
    _check(_errors(got, ref64, ref64["alive"].astype(bool)), "cfg2 1.0M jelly vs fp64 oracle")


@pytest.mark.needs_cuda
def test_config3_full_size_substep_vs_port():
    scene, st = _perturbed_config("sand256", 1.0, seed=12, strain=0.01, vel=0.5)
    assert len(st["x"]) == 8_000_000 and tuple(scene["res"]) == (256, 256, 256)
    got, c = _engine_substep(scene, st)
    assert c["alive"] == 8_000_000
    ref, alive = _port_substep(scene, st)
    err = _errors(got, ref, alive)
    _check(err, "cfg3 8.0M sand vs fp32 port")
    # all three branches of SandParticle::project are populated at this size
    moved = np.abs(ref["ps"] - st["ps"]) > 0
    assert moved.any() and (~moved).any()


@pytest.mark.parametrize("name,scale", [("snow256", 0.25), ("water512", 0.125)])
def test_reduced_config4_config5_vs_fp64_oracle(name, scale):
    from oracle import pyoracle as O
    scene, st = _perturbed_config(name, scale, seed=13, strain=0.02, vel=0.5)
    got, _ = _engine_substep(scene, st)
    ref, _, _ = O.substep(scene, st, np.float64, want_grids=False)
    _check(_errors(got, ref, ref["alive"].astype(bool)), "%s x%g vs fp64 oracle" % (name, scale))


@pytest.mark.parametrize("strain", [0.1, 0.3])
def test_sand_large_strain_takes_the_eigen_path_on_the_device(strain):
    """|F F^T - I|_F > 0.15 for (nearly) every particle: sand_step_series returns false and material_step falls
    through to the Jacobi path.  det F > 0 is kept (the inverted-element convention is the documented difference)."""
    scene, st = T.perturbed_scene(scenes.MAT_SAND, res=32, cells=8, seed=21, strain=strain)
    F = st["F"].reshape(-1, 3, 3).transpose(0, 2, 1).astype(np.float64)
    keep = np.linalg.det(F) > 0.2
    E = F @ F.transpose(0, 2, 1) - np.eye(3)
    big = np.sqrt((E ** 2).sum((1, 2))) > 0.15
    assert big[keep].mean() > 0.7
    st = {k: v[keep] for k, v in st.items()}
    e = T.make_engine(scene, st)
    err, got, ref = T.compare_substep(e, scene, st)
    e.close()
    print("sand strain %g:" % strain, {k: float(v) for k, v in err.items() if not isinstance(v, bool)})
    assert err["alive_match"]
    assert err["grid_rast"] <= T.TOL_GRID_REL
    assert err["v"] <= T.TOL_V_REL and err["b"] <= T.TOL_V_REL and err["x"] <= T.TOL_X_ABS
    assert err["F"] <= T.TOL_F_ABS * max(1.0, 10 * strain) and err["ps"] <= T.TOL_PS_ABS * max(1.0, 10 * strain)


def test_full_size_config2_with_rigid_bodies_one_coupled_substep_vs_fp64_oracle():
    # BASELINE config 2 (128^3 grid, 1.0 M jelly particles, every particle perturbed) cut by a tilted free plate and holding a small
    # free box: one coupled substep — colour field, gather_cdf, both transfers with their impulses — per particle against the oracle's
    # restatement of the reference's CPIC path (itself pinned to src/rigid_transfer.cpp / block_op_rigid run in place)
    from oracle import pyoracle as O
    scene, st = _perturbed_config("jelly128", 1.0, seed=81, strain=0.01, vel=0.3)
    n = len(st["x"])
    assert n == 1000000
    c = st["x"].mean(0)
    dx = scene["dx"]
    plate = dict(tris=scenes.plate_mesh(0.27, 0.23, axis=1), position=c + np.array([0.004, 0.031, -0.003]), rotation=scenes.euler_rotation((7.0, 13.0, -5.0)),
                 velocity=(0.1, -0.8, 0.05), angular_velocity=(0.3, 0.0, -0.4), frictions=(0.3, 0.5), inv_mass=1 / 30.0, inv_inertia=np.diag([4.0, 2.5, 4.0]))
    box = dict(tris=scenes.box_mesh((0.05, 0.04, 0.06)), position=c + np.array([0.09, -0.08, 0.02]), rotation=scenes.euler_rotation((20.0, 5.0, 33.0)),
               velocity=(-0.5, 0.0, 0.2), friction=-1.0, inv_mass=0.2, inv_inertia=np.diag([30.0, 30.0, 30.0]))
    rigid = scenes.make_rigid([plate, box], dx, penalty=1e3)
    ref, grid_rast, _, rref, cdf = O.substep_coupled(scene, st, rigid, np.float64)
    e = T.make_engine(dict(scene, sdf=None), st)
    e.set_rigid(rigid)
    e.sort_particles_and_populate_grid()
    assert np.array_equal(e.download_cdf()["node_state"], cdf["node_state"])
    pc = e.get_particle_cdf(n)
    same = pc["states"] == ref["states"]
    assert (ref["states"] != 0).sum() > 50000 and same.mean() > 1 - 2e-5       # a colour decided by two nearly equal weighted distances may flip in fp32
    assert np.array_equal(pc["near"][same], ref["near"][same]) and ref["near"].sum() > 20000
    e.rasterize()
    g0 = e.download_grid(0).astype(np.float64)
    pmax = max(np.abs(grid_rast[..., :3]).max(), grid_rast[..., 3].max())
    flipped_cells = (~same).sum()
    assert np.abs(g0 - grid_rast).max() <= (T.TOL_GRID_REL if flipped_cells == 0 else 1e-3) * pmax
    e.resample()
    got = e.download()
    rs = e.get_rigid_state(3)
    e.close()
    alive = ref["alive"].astype(bool)
    ids = got["id"].astype(np.int64)
    assert np.array_equal(ids, np.nonzero(alive)[0])
    ok = same[ids]
    if flipped_cells:   # a flipped colour changes the node values its particle scatters to: compare away from those particles' stencils
        bad = np.zeros(n, bool)
        bx = np.floor(st["x"] / dx).astype(int)
        for p in np.nonzero(~same)[0]:
            bad |= (np.abs(bx - bx[p]).max(1) <= 3)
        ok &= ~bad[ids]
    vmax = np.abs(ref["v"]).max()
    err = dict(x=np.abs(got["x"] - ref["x"][ids])[ok].max(), v=np.abs(got["v"] - ref["v"][ids])[ok].max() / vmax,
               b=np.abs(got["b"] - ref["b"][ids])[ok].max() / np.abs(ref["b"]).max(), F=np.abs(got["F"] - ref["F"][ids])[ok].max(), ps=0.0)
    _check(err, "config 2 + rigid bodies vs fp64 oracle (%d colour flips of %d coloured particles)" % (flipped_cells, int((ref["states"] != 0).sum())))
    for b in (1, 2):
        dv = rref["velocity"][b] - rigid["velocity"][b]
        assert np.abs(dv).max() > 1e-4
        assert np.abs((rs["velocity"][b] - rigid["velocity"][b]) - dv).max() <= 5e-3 * np.abs(dv).max() + 1e-6
