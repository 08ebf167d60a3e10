"""GPU parity: the CUDA path, called through the C-ABI, against the fp64 oracle on identical
seeded inputs (tolerances of SURVEY.md §8d, stated in tests/common.py)."""
import numpy as np
import pytest

from taichi_mpm_b200 import scenes
from tests import common as T

pytestmark = pytest.mark.gpu

KINDS = [("linear", scenes.MAT_LINEAR), ("jelly", scenes.MAT_JELLY), ("snow", scenes.MAT_SNOW), ("water", scenes.MAT_WATER),
         ("sand", scenes.MAT_SAND), ("elastic", scenes.MAT_ELASTIC), ("von_mises", scenes.MAT_VON_MISES), ("visco", scenes.MAT_VISCO)]
KIND_KW = T.KIND_KW


def _assert_parity(err):
    assert err["alive_match"]
    assert err["grid_rast"] <= T.TOL_GRID_REL, err
    assert err["grid_vel"] <= T.TOL_GRID_REL, err       # node velocities after normalise + boundary: same 2e-5 bound (measured 1e-6..5e-6 on a B200)
    assert err["grid_mass_outside"] == 0.0, err
    assert err["x"] <= T.TOL_X_ABS, err
    assert err["v"] <= T.TOL_V_REL, err
    assert err["b"] <= T.TOL_V_REL, err
    assert err["F"] <= T.TOL_F_ABS, err
    assert err["ps"] <= T.TOL_PS_ABS, err
    assert err["mass"] == 0.0


@pytest.mark.parametrize("name,kind", KINDS)
def test_single_substep_vs_fp64_oracle(name, kind):
    scene, st = T.perturbed_scene(kind, res=32, cells=8, seed=3, **KIND_KW.get(kind, {}))
    e = T.make_engine(scene, st)
    err, got, ref = T.compare_substep(e, scene, st)
    if kind == scenes.MAT_VON_MISES:   # both branches of the return map ran (src/particles.cpp:726-733)
        eps = np.log(np.linalg.svd(st["F"].reshape(-1, 3, 3).astype(np.float64), compute_uv=False))
        n2 = ((eps - eps.mean(1, keepdims=True)) ** 2).sum(1)
        yielding = n2 > scene["mat_params"][0][2] / (2 * scene["mat_params"][0][0])
        assert 0.1 < yielding.mean() < 0.9, yielding.mean()
    if kind == scenes.MAT_VISCO:
        moved = np.abs(ref["ps"] - st["ps"]) > 1e-3
        assert 0.05 < moved.mean() < 0.95, moved.mean()
    print(name, {k: (float(v) if not isinstance(v, bool) else v) for k, v in err.items()})
    _assert_parity(err)
    e.close()


@pytest.mark.parametrize("name,kind", KINDS)
def test_single_substep_vs_fp32_oracle(name, kind):
    # same bounds against the fp32 restatement (reference operation order)
    from oracle import pyoracle as O
    scene, st = T.perturbed_scene(kind, res=32, cells=6, seed=4, **KIND_KW.get(kind, {}))
    e = T.make_engine(scene, st)
    ref, _, _ = O.substep(scene, st, np.float32)
    e.substep(1)
    got = e.download()
    ids = got["id"].astype(np.int64)
    vmax = np.abs(ref["v"]).max()
    assert np.abs(got["v"] - ref["v"][ids]).max() <= 2 * T.TOL_V_REL * vmax
    assert np.abs(got["F"] - ref["F"][ids]).max() <= 2 * T.TOL_F_ABS
    assert np.abs(got["x"] - ref["x"][ids]).max() <= T.TOL_X_ABS
    e.close()


def test_sand_plastic_branches_are_exercised():
    # the three Drucker-Prager cases (expansion / inside cone / projected) all occur and agree
    from oracle import pyoracle as O
    scene, st = T.perturbed_scene(scenes.MAT_SAND, res=32, cells=8, seed=5, strain=0.01)
    ref, _, _ = O.substep(scene, st, np.float64)
    Ftrial_changed = np.abs(ref["F"] - st["F"]).max(1)
    assert (ref["ps"] > 0).any() and (ref["ps"] == 0).any()
    e = T.make_engine(scene, st)
    err, got, ref = T.compare_substep(e, scene, st, check_grid=False)
    assert err["F"] <= T.TOL_F_ABS and err["ps"] <= T.TOL_PS_ABS, err
    e.close()


def test_dense_sdf_equals_planes():
    scene, st = T.perturbed_scene(scenes.MAT_JELLY, res=32, cells=6, seed=6, friction=-1.0)
    e1 = T.make_engine(scene, st)
    scene2 = dict(scene)
    scene2["sdf_dense_upload"] = scene["sdf"]
    e2 = T.make_engine(scene2, st)
    e1.substep(2)
    e2.substep(2)
    a, b = e1.download(), e2.download()
    # bitwise: the engine has no float atomics, every sum has a fixed order
    for k in ("x", "v", "F", "b"):
        assert np.array_equal(a[k], b[k]), k
    e1.close(); e2.close()


@pytest.mark.parametrize("friction", [-1.0, -2.3, 0.0, 0.4])
def test_boundary_modes(friction):
    scene, st = T.perturbed_scene(scenes.MAT_JELLY, res=32, cells=6, seed=7, friction=friction, vel=1.0)
    e = T.make_engine(scene, st)
    err, _, _ = T.compare_substep(e, scene, st)
    _assert_parity(err)
    e.close()


def test_grid_gravity_mode():
    scene, st = T.perturbed_scene(scenes.MAT_JELLY, res=32, cells=6, seed=8)
    scene["particle_gravity"] = 0
    e = T.make_engine(scene, st)
    err, _, _ = T.compare_substep(e, scene, st)
    _assert_parity(err)
    e.close()


def test_two_materials_in_one_scene():
    sc1, st1 = T.perturbed_scene(scenes.MAT_SAND, res=32, cells=6, seed=9)
    x2, m2, v2 = scenes.lattice_block(32, (8, 18, 8), (12, 22, 12), jitter=0.1, seed=10)
    st2 = scenes.make_state(x2, m2, v2, scenes.MAT_WATER, group=1)
    st = {k: np.concatenate([st1[k], st2[k]]) for k in st1}
    scene = dict(sc1)
    scene["mat_kind"] = np.array([scenes.MAT_SAND, scenes.MAT_WATER], np.int32)
    scene["mat_params"] = np.stack([scenes.material_params(scenes.MAT_SAND), scenes.material_params(scenes.MAT_WATER)])
    e = T.make_engine(scene, st)
    err, _, _ = T.compare_substep(e, scene, st)
    _assert_parity(err)
    e.close()


def test_particle_deletion_matches_reference_band():
    scene, st = T.perturbed_scene(scenes.MAT_JELLY, res=32, cells=4, seed=11, with_floor=False)
    res = 32
    st["x"][0] = [6.9 / res, 0.5, 0.5]
    st["x"][1] = [0.5, (res - 6.9) / res, 0.5]
    st["x"][2] = [7.6 / res, 0.5, 0.5]
    st["v"][:3] = 0
    e = T.make_engine(scene, st)
    err, got, ref = T.compare_substep(e, scene, st, check_grid=False)
    assert err["alive_match"]
    assert 0 not in got["id"] and 1 not in got["id"] and 2 in got["id"]
    assert e.num_particles() == len(st["x"]) - 2
    # deleted particles stay deleted and the rest keep evolving
    e.substep(3)
    assert e.num_particles() == len(st["x"]) - 2
    e.close()


def test_empty_and_tiny_inputs():
    scene, st = T.perturbed_scene(scenes.MAT_JELLY, res=32, cells=4, seed=12)
    one = {k: v[:1] for k, v in st.items()}
    e = T.make_engine(scene, one)
    err, _, _ = T.compare_substep(e, scene, one)
    assert err["alive_match"] and err["v"] <= T.TOL_V_REL
    e.close()
    from taichi_mpm_b200 import capi
    e = capi.Engine(scene["res"], scene["dx"], scene["dt"])
    e.upload(np.zeros((0, 3)), np.zeros((0, 3)), np.zeros(0), np.zeros(0))
    e.substep(2)
    assert e.num_particles() == 0
    e.close()


def test_stage_order_is_enforced():
    from taichi_mpm_b200 import capi
    scene, st = T.perturbed_scene(scenes.MAT_JELLY, res=32, cells=4, seed=13)
    e = T.make_engine(scene, st)
    with pytest.raises(capi.MpmbError):
        e.rasterize()
    e.sort_particles_and_populate_grid()
    with pytest.raises(capi.MpmbError):
        e.resample()
    e.close()


def test_multi_step_invariants_sand():
    # >=100 substeps: trajectories are chaotic, compare invariants
    from oracle import pyoracle as O
    scene, st = T.perturbed_scene(scenes.MAT_SAND, res=32, cells=8, seed=14, strain=0.0, vel=0.2, with_floor=True)
    e = T.make_engine(scene, st)
    fast = O.FastOracle(scene, st, threads=4)
    nsub = 100
    e.substep(nsub)
    fast.substeps(nsub)
    got = e.download()
    alive = fast.st["alive"].astype(bool)
    assert len(got["id"]) == alive.sum()
    m = got["mass"].astype(np.float64)
    assert m.sum() == pytest.approx(fast.st["mass"][alive].astype(np.float64).sum(), rel=1e-12)
    com_g = (m[:, None] * got["x"]).sum(0) / m.sum()
    mo = fast.st["mass"][alive].astype(np.float64)
    com_o = (mo[:, None] * fast.st["x"][alive]).sum(0) / mo.sum()
    assert np.abs(com_g - com_o).max() < 2e-5
    mom_g = (m[:, None] * got["v"]).sum(0)
    mom_o = (mo[:, None] * fast.st["v"][alive]).sum(0)
    assert np.abs(mom_g - mom_o).max() <= 2e-3 * max(np.abs(mom_o).max(), m.sum() * 0.01)
    assert np.isfinite(got["F"]).all()
    e.close()


def test_aos_round_trip():
    # the reference's 320 B particle slots in, the same slots out (drop-in adapter path)
    from taichi_mpm_b200 import capi
    scene, st = T.perturbed_scene(scenes.MAT_SAND, res=32, cells=4, seed=15)
    n = len(st["x"])
    L = capi.MpmbAosLayout()
    L.stride, L.off_v_and_m, L.off_pos, L.off_dg_e, L.off_apic_b, L.col_pitch = 320, 16, 32, 48, 96, 16
    L.off_vol, L.off_scalar = 164, 200
    slots = n + 7
    pool = np.zeros((slots, 320), np.uint8)
    rng = np.random.default_rng(1)
    indices = rng.permutation(slots)[:n].astype(np.uint32)
    f = pool.view(np.float32).reshape(slots, 80)
    for k, s in enumerate(indices):
        f[s, 4:7] = st["v"][k]; f[s, 7] = st["mass"][k]
        f[s, 8:11] = st["x"][k]
        for c in range(3):
            f[s, 12 + 4 * c: 15 + 4 * c] = st["F"][k, 3 * c: 3 * c + 3]
            f[s, 24 + 4 * c: 27 + 4 * c] = st["b"][k, 3 * c: 3 * c + 3]
        f[s, 41] = st["vol"][k]
        f[s, 50] = st["ps"][k]
    e = capi.Engine(scene["res"], scene["dx"], scene["dt"], scene["gravity"])
    e.set_material(0, scenes.MAT_SAND, scene["mat_params"][0])
    e.set_planes(scene["planes"], scene["friction"])
    e.upload_aos(pool, indices, L)
    e2 = T.make_engine(scene, st)
    e.substep(1); e2.substep(1)
    ref = e2.download()
    idx = indices.copy()
    n_alive = e.download_aos(pool, idx, L)
    assert n_alive == len(ref["id"])
    f = pool.view(np.float32).reshape(slots, 80)
    for k in range(n_alive):
        s = idx[k]
        assert s == indices[ref["id"][k]]
        assert np.array_equal(f[s, 8:11], ref["x"][k]) and np.array_equal(f[s, 4:7], ref["v"][k])
        assert f[s, 50] == ref["ps"][k]
    e.close(); e2.close()


def test_many_movers_per_step_stay_consistent():
    # fast motion: dozens of particles change tile every substep, so the incremental ordering
    # (runs with holes + arrival lists) is exercised hard; compare with the CPU fast path
    from oracle import pyoracle as O
    from tests.test_gpu_slab import _scene
    scene, st = _scene()
    scene = dict(scene)
    scene["sdf"] = scenes.planes_sdf(scene["res"], scene["planes"])
    e = T.make_engine(scene, st)
    fast = O.FastOracle(scene, st, threads=4)
    nsub = 60
    e.substep(nsub)
    fast.substeps(nsub)
    got = e.download()
    alive = fast.st["alive"].astype(bool)
    assert len(got["id"]) == alive.sum() == len(st["x"])
    assert len(np.unique(got["id"])) == len(got["id"])
    ids = got["id"].astype(np.int64)
    # individual trajectories still agree closely over this horizon
    assert np.abs(got["x"] - fast.st["x"][ids]).max() < 5e-5
    assert np.abs(got["v"] - fast.st["v"][ids]).max() < 2e-2 * np.abs(fast.st["v"]).max()
    m = got["mass"].astype(np.float64)
    assert np.all(m > 0)
    e.close()


def test_dense_tiles_need_several_chunks():
    # 16 particles per cell -> ~1000 rows per tile: P2G (512-row chunks) and G2P (256-row chunks) both
    # take the multi-chunk path, with holes and arrivals appearing as the block deforms
    from oracle import pyoracle as O
    res = 32
    xa, ma, va = scenes.lattice_block(res, (10, 9, 10), (18, 17, 18), jitter=0.2, seed=21)
    xb, mb, vb = scenes.lattice_block(res, (10, 9, 10), (18, 17, 18), jitter=0.2, seed=22)
    x = np.concatenate([xa, xb]); mass = np.concatenate([ma, mb]) * 0.5; vol = np.concatenate([va, vb]) * 0.5
    st = scenes.make_state(x, mass, vol, scenes.MAT_SAND)
    rng = np.random.default_rng(23)
    st["v"] = (rng.normal(size=x.shape) * 0.5 + np.array([1.5, 0.0, -1.0])).astype(np.float32)
    planes = np.array([[0.0, 1.0, 0.0, -9.6]], np.float32)
    scene = dict(res=(res,) * 3, dx=1.0 / res, dt=1e-4, gravity=(0.0, -10.0, 0.0), particle_gravity=1,
                 mat_kind=np.array([scenes.MAT_SAND], np.int32), mat_params=scenes.material_params(scenes.MAT_SAND)[None],
                 planes=planes, friction=0.4, sdf=scenes.planes_sdf(res, planes))
    e = T.make_engine(scene, st)
    err, got, ref = T.compare_substep(e, scene, st)          # single substep, full parity vs fp64
    assert err["alive_match"] and err["grid_rast"] <= T.TOL_GRID_REL and err["v"] <= T.TOL_V_REL and err["F"] <= T.TOL_F_ABS, err
    from taichi_mpm_b200 import slab
    tb = np.stack([slab.base_tile_z(x[:, d], 1.0 / res) for d in range(3)], 1)
    per_tile = np.unique(tb[:, 0] * 10000 + tb[:, 1] * 100 + tb[:, 2], return_counts=True)[1]
    assert per_tile.max() > 2 * 512 - 64, per_tile.max()      # at least one tile needs 2 P2G chunks and 4 G2P chunks
    fast = O.FastOracle(scene, ref, threads=4)                # continue both for 40 more substeps
    e.substep(40)
    fast.substeps(40)
    got = e.download()
    ids = got["id"].astype(np.int64)
    assert len(ids) == len(x) and len(np.unique(ids)) == len(ids)
    assert np.abs(got["x"] - fast.st["x"][ids]).max() < 5e-5
    e.close()


def test_a_few_overfull_cells_among_ordinary_ones():
    """The cell-owner loop of k_p2g runs as long as the fullest cell of a warp has particles: a regular 8-per-cell block
    plus three cells holding 30..45 particles (what a developed flow looks like locally) — node momenta and mass must
    still match the fp64 oracle, and the ordering must keep the crowded cells consistent."""
    res = 32
    rng = np.random.default_rng(31)
    x, mass, vol = scenes.lattice_block(res, (10, 9, 10), (18, 17, 18), jitter=0.1, seed=30)
    extra = []
    for cell, n in (((11, 10, 12), 45), ((13, 12, 15), 30), ((16, 15, 11), 37)):     # base nodes inside different tiles / warps
        extra.append(((np.asarray(cell) + 0.55 + 0.9 * rng.random((n, 3))) / res).astype(np.float32))
    xe = np.concatenate(extra)
    x = np.concatenate([x, xe]); mass = np.concatenate([mass, np.full(len(xe), mass[0] * 0.2, np.float32)])
    vol = np.concatenate([vol, np.full(len(xe), vol[0] * 0.2, np.float32)])
    st = scenes.make_state(x, mass, vol, scenes.MAT_SAND)
    T.perturb_state(st, scenes.MAT_SAND, 1.0 / res, seed=32, strain=0.01, vel=0.5)
    planes = np.array([[0.0, 1.0, 0.0, -9.6]], np.float32)
    scene = dict(res=(res,) * 3, dx=1.0 / res, dt=2e-5, gravity=(0.0, -10.0, 0.0), particle_gravity=1,
                 mat_kind=np.array([scenes.MAT_SAND], np.int32), mat_params=scenes.material_params(scenes.MAT_SAND)[None],
                 planes=planes, friction=0.4, sdf=scenes.planes_sdf(res, planes))
    base = (x * res - 0.5).astype(np.int32)
    per_cell = np.unique(base[:, 0] * 10000 + base[:, 1] * 100 + base[:, 2], return_counts=True)[1]
    assert per_cell.max() >= 38 and np.sort(per_cell)[-4] <= 14          # three overfull cells, the rest ordinary
    e = T.make_engine(scene, st)
    err, got, ref = T.compare_substep(e, scene, st)
    assert err["alive_match"] and err["grid_rast"] <= T.TOL_GRID_REL and err["grid_vel"] <= 1e-4, err
    assert err["v"] <= T.TOL_V_REL and err["F"] <= T.TOL_F_ABS and err["x"] <= T.TOL_X_ABS, err
    e.substep(20)                                                         # and the ordering keeps them consistent
    assert e.num_particles() == len(x)
    e.close()


def test_reupload_replaces_the_resident_set():
    # the drop-in adapter re-uploads after host-side changes (add_particles, load): same engine, new set
    sc1, st1 = T.perturbed_scene(scenes.MAT_SAND, res=32, cells=6, seed=31)
    sc2, st2 = T.perturbed_scene(scenes.MAT_SAND, res=32, cells=8, seed=32)
    e = T.make_engine(sc1, st1)
    e.substep(5)
    assert e.num_particles() == len(st1["x"])
    e.upload(st2["x"], st2["v"], st2["mass"], st2["vol"], st2["F"], st2["b"], st2["ps"], st2["group"])   # larger set
    e.substep(7)
    a = e.download()
    f = T.make_engine(sc2, st2)
    f.substep(7)
    b = f.download()
    assert len(a["id"]) == len(b["id"]) == len(st2["x"])
    for k in ("x", "v", "F", "b", "ps"):
        assert np.array_equal(a[k], b[k]), k          # bit-reproducible: no float atomics anywhere
    e.upload(st1["x"], st1["v"], st1["mass"], st1["vol"], st1["F"], st1["b"], st1["ps"], st1["group"])   # smaller again
    e.substep(3)
    g = T.make_engine(sc1, st1)
    g.substep(3)
    c, d = e.download(), g.download()
    for k in ("x", "v", "F"):
        assert np.array_equal(c[k], d[k]), k
    e.close(); f.close(); g.close()


@pytest.mark.parametrize("name,kind", [("jelly", scenes.MAT_JELLY), ("snow", scenes.MAT_SNOW), ("water", scenes.MAT_WATER),
                                       ("elastic", scenes.MAT_ELASTIC), ("von_mises", scenes.MAT_VON_MISES), ("visco", scenes.MAT_VISCO)])
def test_multi_step_other_materials(name, kind):
    from oracle import pyoracle as O
    scene, st = T.perturbed_scene(kind, res=32, cells=6, seed=41, strain=0.0, vel=0.5)
    e = T.make_engine(scene, st)
    fast = O.FastOracle(scene, st, threads=4)
    e.substep(50)
    fast.substeps(50)
    got = e.download()
    ids = got["id"].astype(np.int64)
    assert len(ids) == fast.st["alive"].sum()
    assert np.abs(got["x"] - fast.st["x"][ids]).max() < 1e-4
    assert np.isfinite(got["F"]).all()
    e.close()


def test_graph_replay_is_bit_identical_to_host_launches_and_follows_state_changes():
    """mpmb_substep replays pairs of substeps as a CUDA graph (captured per buffer parity).  Same kernels in the same
    order: results must be bit-identical to launching every kernel from the host, across odd/even call lengths, a
    material change, a level-set change and a re-upload (each of which must invalidate the captured graphs)."""
    scene, st = T.perturbed_scene(scenes.MAT_SAND, res=32, cells=8, seed=41, vel=1.5)
    out = []
    for no_graph in (False, True):
        e = T.make_engine(scene, st, no_graph=no_graph)
        e.substep(9)                                           # fresh + pairs + last
        e.substep(4)
        e.set_material(0, scenes.MAT_SAND, scenes.material_params(scenes.MAT_SAND, friction_angle=40.0))
        e.substep(7)
        e.set_planes(np.array([[0.0, 1.0, 0.0, -9.9]], np.float32), 0.2)
        e.substep(6)
        a = e.download()
        e.upload(st["x"], st["v"], st["mass"], st["vol"], st["F"], st["b"], st["ps"], st["group"])   # same capacity: no realloc
        e.substep(8)
        b = e.download()
        out.append((a, b))
        e.close()
    for k in ("id", "x", "v", "F", "ps", "b"):
        assert np.array_equal(out[0][0][k], out[1][0][k]), k
        assert np.array_equal(out[0][1][k], out[1][1][k]), k


@pytest.mark.parametrize("name,kind", [("sand", scenes.MAT_SAND), ("jelly", scenes.MAT_JELLY)])
def test_delta_t_changes_between_substeps_as_asyncmpm_sets_it(name, kind):
    # AsyncMPM sets base_delta_t before every MPM::substep() it schedules (src/async/async_mpm.cpp:407-409): resident particles
    # keep going with the new step (their cached affine matrices are rebuilt), and an upload after the change needs nothing
    from oracle import pyoracle as O
    scene, st = T.perturbed_scene(kind, res=32, cells=6, seed=51, vel=0.4)
    dts = [scene["dt"], scene["dt"] * 0.5, scene["dt"] * 0.25, scene["dt"]]
    cur = st
    for dt in dts:
        cur, _, _ = O.substep(dict(scene, dt=dt), cur, np.float64)
    e = T.make_engine(scene, st)
    for dt in dts:
        e.set_delta_t(dt)
        e.substep(1)
    got = e.download()
    ids = got["id"].astype(np.int64)
    assert np.array_equal(np.sort(ids), np.nonzero(cur["alive"])[0])
    vmax = np.abs(cur["v"]).max()
    assert np.abs(got["x"] - cur["x"][ids]).max() <= 4 * T.TOL_X_ABS
    assert np.abs(got["v"] - cur["v"][ids]).max() <= 4 * T.TOL_V_REL * vmax
    assert np.abs(got["F"] - cur["F"][ids]).max() <= 4 * T.TOL_F_ABS
    # the same through a re-upload under the new step, and it is not what the old step would have given
    f = T.make_engine(scene, st)
    f.set_delta_t(dts[1])
    f.upload(st["x"], st["v"], st["mass"], st["vol"], st["F"], st["b"], st["ps"], st["group"])
    f.substep(1)
    one, _, _ = O.substep(dict(scene, dt=dts[1]), st, np.float64)
    same_dt, _, _ = O.substep(scene, st, np.float64)
    g = f.download()
    assert np.abs(g["x"] - one["x"][g["id"].astype(np.int64)]).max() <= T.TOL_X_ABS
    assert np.abs(one["x"] - same_dt["x"]).max() > 3 * T.TOL_X_ABS
    from taichi_mpm_b200 import capi
    with pytest.raises(capi.MpmbError):
        f.set_delta_t(0.0)
    e.close(); f.close()
