"""GPU parity of the CPIC rigid-coupled path: the colour field, the particle colours, both transfers with
their impulses, against the oracle's restatement of src/rigid_transfer.cpp and the block_op_rigid branches of
src/transfer.cpp on identical inputs.  Runs on the SIMT emulator too (tests/test_simt_emulated.py)."""
import importlib.util
import os

import numpy as np
import pytest

from taichi_mpm_b200 import scenes
from tests import common as T

pytestmark = pytest.mark.gpu


HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_rigid_golden", os.path.join(HERE, "golden", "make_rigid_golden.py"))
G = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(G)


def _scene(dynamic=False, two_bodies=False, penalty=1e3):
    """The scenes of the reference's golden run (tests/golden/make_rigid_golden.py): a stirred block cut by a tilted plate."""
    scene, st, rigid = G.golden_scene("two_bodies" if two_bodies else ("dynamic" if dynamic else "kinematic"))
    rigid["penalty"] = penalty
    return scene, st, rigid


def _engine(scene, st, rigid, states=None):
    e = T.make_engine(scene, st)
    e.set_rigid(rigid)
    if states is not None:
        e.set_particle_states(states)
    return e


@pytest.mark.parametrize("variant", ["kinematic", "dynamic", "two_bodies"])
def test_coupled_substep_vs_fp64_oracle(variant):
    from oracle import pyoracle as O
    scene, st, rigid = _scene(dynamic=variant != "kinematic", two_bodies=variant == "two_bodies")
    n = len(st["x"])
    ref, grid_rast, grid_vel, rref, cdf = O.substep_coupled(scene, st, rigid, np.float64)
    e = _engine(scene, st, rigid)
    e.sort_particles_and_populate_grid()
    # the colour field: same nodes, same tags, same nearest body, same distances
    dcdf = e.download_cdf()
    assert (cdf["node_state"] >> 24 != 0).sum() > 200
    assert np.array_equal(dcdf["node_state"], cdf["node_state"])
    assert np.abs(dcdf["node_dist"] - cdf["node_dist"]).max() <= 2e-6 * scene["dx"] * 3
    # the particle colours and the reconstructed boundary
    pc = e.get_particle_cdf(n)
    assert np.array_equal(pc["states"], ref["states"]) and len(np.unique(ref["states"])) >= 3
    assert np.array_equal(pc["near"], ref["near"]) and 0 < ref["near"].sum() < n
    assert np.abs(pc["bdist"] - ref["bdist"]).max() <= 2e-4 * scene["dx"]
    assert np.abs(pc["bnormal"] - ref["bnormal"]).max() <= 2e-4
    e.rasterize()
    g0 = e.download_grid(0).astype(np.float64)
    pmax = max(np.abs(grid_rast[..., :3]).max(), grid_rast[..., 3].max())
    assert np.abs(g0 - grid_rast).max() <= T.TOL_GRID_REL * pmax
    uncoupled, ug, _ = O.substep(scene, st, np.float64)
    assert np.abs(ug - grid_rast).max() > 1e-2 * pmax            # the colour mask really removed contributions
    e.resample()
    got = e.download()
    ids = got["id"].astype(np.int64)
    assert np.array_equal(np.sort(ids), np.nonzero(ref["alive"])[0])
    vmax = np.abs(ref["v"]).max()
    assert np.abs(got["x"] - ref["x"][ids]).max() <= T.TOL_X_ABS
    assert np.abs(got["v"] - ref["v"][ids]).max() <= T.TOL_V_REL * vmax
    assert np.abs(got["b"] - ref["b"][ids]).max() <= T.TOL_V_REL * max(np.abs(ref["b"]).max(), 1e-30)
    assert np.abs(got["F"] - ref["F"][ids]).max() <= T.TOL_F_ABS
    assert np.abs(ref["v"] - uncoupled["v"]).max() > 0.05 * vmax  # ... and the coupling moved particles
    # the bodies: both transfers' impulses applied (scripted bodies keep their velocity)
    rs = e.get_rigid_state(len(rigid["inv_mass"]))
    for b in range(1, len(rigid["inv_mass"])):
        dv_ref = rref["velocity"][b] - rigid["velocity"][b]
        dw_ref = rref["angular_velocity"][b] - rigid["angular_velocity"][b]
        if rigid["inv_mass"][b] == 0:
            assert np.array_equal(rs["velocity"][b], rigid["velocity"][b]) and np.array_equal(rs["angular_velocity"][b], rigid["angular_velocity"][b])
        else:
            assert np.abs(dv_ref).max() > 1e-4
            assert np.abs((rs["velocity"][b] - rigid["velocity"][b]) - dv_ref).max() <= 2e-3 * np.abs(dv_ref).max() + 1e-6
            assert np.abs((rs["angular_velocity"][b] - rigid["angular_velocity"][b]) - dw_ref).max() <= 2e-3 * np.abs(dw_ref).max() + 1e-5
    e.close()


@pytest.mark.parametrize("variant", G.VARIANTS)
def test_engine_cpic_matches_golden_run_of_reference(variant):
    # the engine against the REFERENCE's own rigid_transfer.cpp / block_op_rigid run (fp32 both sides), no oracle in between
    scene, st, rigid = G.golden_scene(variant)
    z = np.load(os.path.join(HERE, "golden", "rigid_ref.npz"))
    nn = tuple(int(r) + 1 for r in scene["res"])
    n = len(st["x"])
    e = _engine(scene, st, rigid, states=st.get("states"))
    e.sort_particles_and_populate_grid()
    dcdf = e.download_cdf()
    assert np.array_equal(dcdf["node_state"], G.dense(z[variant + "_nstate_idx"], z[variant + "_nstate_val"], nn))
    assert np.abs(dcdf["node_dist"] - G.dense(z[variant + "_ndist_idx"], z[variant + "_ndist_val"], nn)).max() <= 3e-7
    pc = e.get_particle_cdf(n)
    assert np.array_equal(pc["states"], z[variant + "_states"]) and np.array_equal(pc["near"], z[variant + "_near"])
    assert np.abs(pc["bdist"] - z[variant + "_bdist"]).max() <= 2e-4 * scene["dx"] and np.abs(pc["bnormal"] - z[variant + "_bnormal"]).max() <= 3e-4
    e.rasterize()
    ref_grid = G.dense(z[variant + "_grid_idx"], z[variant + "_grid_val"], nn + (4,))
    pmax = max(np.abs(ref_grid[..., :3]).max(), ref_grid[..., 3].max())
    assert np.abs(e.download_grid(0) - ref_grid).max() <= T.TOL_GRID_REL * pmax
    e.resample()
    got = e.download()
    ids = got["id"].astype(np.int64)
    assert np.array_equal(np.sort(ids), np.nonzero(z[variant + "_alive"])[0])
    ref = {k: z["%s_%s" % (variant, k)][ids] for k in ("x", "v", "F", "b", "ps")}
    assert np.abs(got["x"] - ref["x"]).max() <= 2 * T.TOL_X_ABS
    assert np.abs(got["v"] - ref["v"]).max() <= 2 * T.TOL_V_REL * np.abs(ref["v"]).max()
    assert np.abs(got["b"] - ref["b"]).max() <= 2 * T.TOL_V_REL * np.abs(ref["b"]).max()
    assert np.abs(got["F"] - ref["F"]).max() <= 2 * T.TOL_F_ABS and T.ps_err(got["ps"], ref["ps"]) <= 2 * T.TOL_PS_ABS
    rs = e.get_rigid_state(len(rigid["inv_mass"]))
    for b in range(1, len(rigid["inv_mass"])):
        dv = z[variant + "_rigid_v"][b] - rigid["velocity"][b]
        dw = z[variant + "_rigid_w"][b] - rigid["angular_velocity"][b]
        assert np.abs((rs["velocity"][b] - rigid["velocity"][b]) - dv).max() <= 2e-3 * np.abs(dv).max() + 1e-6
        assert np.abs((rs["angular_velocity"][b] - rigid["angular_velocity"][b]) - dw).max() <= 2e-3 * np.abs(dw).max() + 1e-5
    e.close()


def test_particle_colours_persist_and_are_cleared_as_in_the_reference():
    # a particle keeps its colour from substep to substep; colours of bodies it no longer touches are dropped
    # (src/rigid_transfer.cpp:162), and a preset colour decides which side of the plate the particle counts on
    from oracle import pyoracle as O
    scene, st, rigid = _scene()
    n = len(st["x"])
    st["x"][-40:, 1] = 0.66 + 0.01 * np.linspace(0, 1, 40, dtype=np.float32)   # a few particles two blocks above the plate: outside every rigid page
    preset = np.zeros(n, np.uint32)
    preset[: n // 2] = 0b1000            # body 1, positive side — also for particles that sit on the negative side
    preset[n // 2:] = 0b110000           # a body that does not exist near them: dropped at the first gather
    st2 = dict(st, states=preset)
    ref, _, _, _, _ = O.substep_coupled(scene, st2, rigid, np.float64)
    e = _engine(scene, st, rigid, states=preset)
    e.substep(1)
    pc = e.get_particle_cdf(n)
    assert np.array_equal(pc["states"], ref["states"])
    stale = (pc["states"][n // 2:] & 0b110000) != 0
    # dropped for the particles of rigid pages; particles whose cell lies outside keep what they had — gather_cdf returns
    # before it looks at them (src/rigid_transfer.cpp:142-146)
    assert (~stale).any() and stale.any()
    fresh, _, _, _, _ = O.substep_coupled(scene, st, rigid, np.float64)
    assert (ref["states"] != fresh["states"]).any()          # the preset overrode what the distances would have said
    e.close()


def test_several_coupled_substeps_follow_the_oracle_with_host_side_advection():
    # the loop a host runs: set the pose, one substep, read the velocities back, advance the pose
    from oracle import pyoracle as O
    scene, st, rigid = _scene(dynamic=True, penalty=0.0)
    n = len(st["x"])
    e = _engine(scene, st, rigid)
    cur = dict(st, states=np.zeros(n, np.uint32))
    r_o = {k: (np.array(v, np.float64) if isinstance(v, np.ndarray) else v) for k, v in rigid.items()}
    r_d = {k: (np.array(v) if isinstance(v, np.ndarray) else v) for k, v in rigid.items()}
    for step in range(6):
        new, _, _, rv, _ = O.substep_coupled(scene, cur, r_o, np.float64)
        cur = dict(new)
        e.set_rigid_state(r_d)
        e.substep(1)
        rs = e.get_rigid_state(2)
        for r, vel, ang in ((r_o, rv["velocity"], rv["angular_velocity"]), (r_d, rs["velocity"], rs["angular_velocity"])):
            r["velocity"], r["angular_velocity"] = np.array(vel), np.array(ang)
            r["position"] = r["position"] + r["velocity"] * scene["dt"]          # translation only: enough for the loop's plumbing
    got = e.download()
    ids = got["id"].astype(np.int64)
    assert np.array_equal(np.sort(ids), np.nonzero(cur["alive"])[0])
    assert np.abs(got["x"] - cur["x"][ids]).max() <= 5e-6
    assert np.abs(got["v"] - cur["v"][ids]).max() <= 2e-3 * np.abs(cur["v"]).max()
    assert np.abs(r_d["velocity"][1] - r_o["velocity"][1]).max() <= 1e-3 * np.abs(r_o["velocity"][1]).max()
    pc = e.get_particle_cdf(n)
    assert (pc["states"] != cur["states"]).mean() < 0.002     # a colour decided by two nearly equal weighted distances may flip in fp32
    e.close()


def test_rigid_coupling_is_refused_where_it_is_not_implemented():
    from taichi_mpm_b200 import capi
    scene, st, rigid = _scene()
    e = T.make_engine(scene, st)
    bad = dict(rigid, sample_rigid=np.full(len(rigid["sample_rigid"]), 5, np.int32))
    with pytest.raises(capi.MpmbError):
        e.set_rigid(bad)                                       # samples naming a body that was not declared
    with pytest.raises(capi.MpmbError):
        e.get_rigid_state(2)                                   # no bodies yet
    e.set_rigid(rigid)
    e.set_rigid(dict(rigid, sample_offset=np.zeros((0, 3)), sample_tri=np.zeros((0, 9)), sample_rigid=np.zeros(0, np.int32)))   # off again
    e.substep(2)
    ref, _, _ = __import__("oracle.pyoracle", fromlist=["x"]).substep(scene, st, np.float64)
    e.close()


def test_mirror_runs_rigid_scenes_and_the_coupling_conserves_momentum():
    # a free box thrown into a jelly block, no gravity, no walls: what the body gains the particles lose (two-way coupling:
    # both transfers' impulses reach the body through apply_tmp_velocity, src/transfer.cpp:578-580, 967-969)
    from taichi_mpm_b200 import mpm as mpm_mod
    # (pushing_force = 0: the reference's artificial push along the boundary normal, src/transfer.cpp:781-782, has no
    # counter-impulse on the body and is the one term of the coupling that does not conserve momentum)
    m = mpm_mod.MPM(res=(32, 32, 32), base_delta_t=1e-4, gravity=(0, 0, 0), pushing_force=0.0)
    m.add_particles(type="jelly", benchmark_block=((12, 12, 12), (20, 18, 20)), density=400.0, jitter=0.2, E=2e4)
    bid = m.add_particles(type="rigid", tris=scenes.box_mesh((0.06, 0.05, 0.06)), codimensional=False, density=2000.0, friction=0.2,
                          initial_position=(0.5, 0.62, 0.5), initial_velocity=(0.05, -1.5, 0.0), initial_rotation=(10.0, 20.0, 5.0))
    assert bid == "1"
    body = m.rigids[0]
    p0 = m.get_particles()
    mom0 = (p0["mass"][:, None] * p0["v"]).sum(0) + body.mass * body.velocity
    y0 = body.position[1]
    for _ in range(3):
        m.step(0.004)
    p1 = m.get_particles()
    mom_p = (p1["mass"][:, None] * p1["v"]).sum(0)
    mom1 = mom_p + body.mass * body.velocity
    assert len(p1["x"]) == len(p0["x"]) and np.isfinite(p1["x"]).all()
    assert body.position[1] < y0 - 0.01                                # it moved down ...
    assert body.velocity[1] > -1.45 and mom_p[1] < -0.02 * abs(mom0[1])   # ... was slowed down by the block, which took momentum
    # the sum is conserved to a few per cent — as far as the reference's scheme conserves it: rasterize hands the body the
    # impulse of the velocity change it projects (src/transfer.cpp:436-445), resample projects again without a counter-impulse
    assert np.abs(mom1 - mom0).max() <= 0.05 * np.abs(mom0).max()
    assert abs(body.mass * (body.velocity[1] + 1.5)) > 10 * np.abs(mom1 - mom0).max()   # the exchange is much larger than the defect
    pc = m.engine.get_particle_cdf(len(p0["x"]))
    assert (pc["states"] != 0).sum() > 50                              # particles near the box carry its colour


def test_mirror_scripted_body_follows_its_functions_and_keeps_infinite_mass():
    from taichi_mpm_b200 import mpm as mpm_mod
    m = mpm_mod.MPM(res=(32, 32, 32), base_delta_t=1e-4, gravity=(0, -10, 0), penalty=1e3)
    ls = m.create_levelset()
    ls.add_plane((0, 1, 0), -0.3)
    ls.set_friction(0.4)
    m.set_levelset(ls, False)
    m.add_particles(type="sand", benchmark_block=((12, 10, 12), (20, 16, 20)), density=400.0, jitter=0.2)
    m.add_particles(type="rigid", tris=scenes.plate_mesh(0.03, 0.12, axis=0), codimensional=True, friction=0.3,
                    scripted_position=lambda t: (0.40 + 1.0 * t, 0.42, 0.5), scripted_rotation=lambda t: (0.0, 0.0, 15.0))
    body = m.rigids[0]
    assert body.inv_mass == 0.0 and not body.inv_inertia_body.any()
    x0 = m.get_particles()["x"].mean(0)
    m.step(0.02)
    assert abs(body.position[0] - (0.40 + float(m.current_t))) < 1e-6 and abs(body.velocity[0] - 1.0) < 1e-3
    p = m.get_particles()
    assert np.isfinite(p["x"]).all() and p["x"].mean(0)[0] > x0[0] + 1e-4     # the paddle pushes the sand along +x
    with pytest.raises(ValueError):
        m.add_particles(type="rigid", tris=scenes.plate_mesh(0.1, 0.1), codimensional=True, initial_position=(0.5, 0.5, 0.5),
                        scripted_position=lambda t: (0.5, 0.5, 0.5))


def test_rigid_state_survives_reupload_resampling_and_bodies_near_the_domain_edge():
    from oracle import pyoracle as O
    scene, st, rigid = _scene(dynamic=True)
    n = len(st["x"])
    e = _engine(scene, st, rigid)
    e.substep(2)
    # (1) new boundary samples with a different count (a second body appears), then a re-upload: colours restart at 0
    scene2, st2, rigid2 = _scene(two_bodies=True)
    e.set_rigid(rigid2)
    e.upload(st2["x"], st2["v"], st2["mass"], st2["vol"], st2["F"], st2["b"], st2["ps"], st2["group"])
    ref, _, _, rref, cdf = O.substep_coupled(scene2, st2, rigid2, np.float64)
    e.sort_particles_and_populate_grid()
    assert np.array_equal(e.download_cdf()["node_state"], cdf["node_state"])      # nothing of the first body's old field is left
    assert np.array_equal(e.get_particle_cdf(n)["states"], ref["states"])
    e.rasterize(); e.resample()
    got = e.download()
    assert np.abs(got["x"] - ref["x"][got["id"].astype(np.int64)]).max() <= T.TOL_X_ABS
    # (2) a body whose samples reach outside the node grid: those samples are skipped (no stencil), the rest still colour nodes
    far = dict(rigid2)
    far["position"] = rigid2["position"].copy()
    far["position"][2] = (0.5, 0.995, 0.5)
    e.set_rigid_state(far)
    e.substep(1)
    assert np.isfinite(e.download()["x"]).all()
    # (3) coupling off again: the plain substep, bit for bit what an engine that never saw a body computes
    e.set_rigid(dict(rigid, sample_offset=np.zeros((0, 3)), sample_tri=np.zeros((0, 9)), sample_rigid=np.zeros(0, np.int32)))
    e.upload(st["x"], st["v"], st["mass"], st["vol"], st["F"], st["b"], st["ps"], st["group"])
    f = T.make_engine(scene, st)
    e.substep(3); f.substep(3)
    a, b = e.download(), f.download()
    for k in ("x", "v", "F"):
        assert np.array_equal(a[k], b[k]), k
    e.close(); f.close()


def test_two_bodies_claiming_one_node_the_nearer_one_wins_and_both_leave_their_tags():
    # two parallel plates less than a cell apart: nodes between them carry both bodies' tag bits and the id of the nearer
    # (src/rigid_transfer.cpp:65-74)
    from oracle import pyoracle as O
    scene, st, _ = _scene()
    dx = scene["dx"]
    c = st["x"].mean(0)
    rot = scenes.euler_rotation((3.0, 8.0, 2.0))
    mk = lambda off: dict(tris=scenes.plate_mesh(0.1, 0.1, axis=1), position=c + np.array([0.0, off, 0.0]), rotation=rot, friction=0.1)
    rigid = scenes.make_rigid([mk(0.0), mk(0.6 * dx)], dx)
    ref, _, _, _, cdf = O.substep_coupled(scene, st, rigid, np.float64)
    tags = cdf["node_state"] & 0xFFFFFF
    both = ((tags & 0b1000) != 0) & ((tags & 0b100000) != 0)
    assert both.sum() > 50 and len(np.unique(cdf["node_state"][both] >> 24)) == 2     # both ids occur as the nearer body
    e = _engine(scene, st, rigid)
    e.sort_particles_and_populate_grid()
    d = e.download_cdf()
    assert np.array_equal(d["node_state"], cdf["node_state"])
    assert np.array_equal(e.get_particle_cdf(len(st["x"]))["states"], ref["states"])
    e.close()


def test_device_seeded_lattice_with_id_gaps_and_a_rigid_body():
    # mpmb_seed_lattice numbers particles by lattice index and skips the 7-cell boundary band: ids have gaps and exceed the particle
    # count; the by-id colour arrays must cover them
    from oracle import pyoracle as O
    from taichi_mpm_b200 import capi
    res, dx = 32, 1.0 / 32
    scene, _, _ = _scene()
    e = capi.Engine(scene["res"], scene["dx"], scene["dt"], scene["gravity"], 1, True)
    e.set_material(0, int(scene["mat_kind"][0]), scene["mat_params"][0])
    vol = dx ** 3 / 8
    n = e.seed_lattice((4, 12, 12), (14, 18, 18), vol, vol * 400.0, jitter=0.1, seed=5)     # cells 4..6 along x lie in the deletion band
    p = e.download()
    ids = p["id"].astype(np.int64)
    assert n == len(ids) and ids.max() + 1 > n                       # gaps
    plate = dict(tris=scenes.plate_mesh(0.12, 0.12, axis=1), position=p["x"].mean(0) + np.array([0.0, 0.004, 0.0]),
                 rotation=scenes.euler_rotation((5.0, 9.0, -4.0)), velocity=(0.0, -0.5, 0.0), friction=0.2)
    rigid = scenes.make_rigid([plate], dx, penalty=1e3)
    st = dict(x=p["x"], v=p["v"], F=p["F"], b=p["b"], mass=p["mass"], vol=p["vol"], ps=p["ps"], group=p["group"])
    ref, _, _, _, _ = O.substep_coupled(scene, st, rigid, np.float64)
    e.set_rigid(rigid)
    e.substep(1)
    pc = e.get_particle_cdf(int(ids.max()) + 1)
    assert np.array_equal(pc["states"][ids], ref["states"]) and (ref["states"] != 0).sum() > 100
    got = e.download()
    assert np.array_equal(got["id"].astype(np.int64), ids[ref["alive"].astype(bool)])
    assert np.abs(got["x"] - ref["x"][ref["alive"].astype(bool)]).max() <= T.TOL_X_ABS
    e.close()
