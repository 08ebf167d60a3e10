"""Device-side seeding: mpmb_seed_lattice builds the `benchmark` lattice of MPM<3>::add_particles
(src/mpm.cpp:149-186) on the device.  Parity reference = its numpy twin scenes.lattice_block_hashed (same float32
operations) and, at jitter 0, the reference's own lattice (scenes.lattice_block, pinned to add_particles(benchmark=...)
of the reference sources compiled in place by tests/test_oracle_ref_transfer.py)."""
import numpy as np
import pytest

from taichi_mpm_b200 import scenes
from tests import common as T

pytestmark = pytest.mark.gpu


def _engine(res=32, kind=scenes.MAT_SAND, dt=2e-5, **kw):
    from taichi_mpm_b200 import capi
    e = capi.Engine((res, res, res), 1.0 / res, dt, (0.0, -10.0, 0.0), True, True, **kw)
    e.set_material(0, kind, scenes.material_params(kind))
    e.set_planes(np.array([[0.0, 1.0, 0.0, -9.6]], np.float32), 0.4)
    return e


@pytest.mark.parametrize("jitter", [0.0, 0.05])
def test_seeded_lattice_equals_the_host_twin_bit_for_bit(jitter):
    res, lo, hi = 32, (10, 10, 11), (18, 16, 19)
    ids, x, mass, vol = scenes.lattice_block_hashed(res, lo, hi, jitter=jitter, seed=1234)
    e = _engine(res)
    n = e.seed_lattice(lo, hi, float(vol[0]), float(mass[0]), jitter=jitter, seed=1234, v0=(0.1, 0.0, -0.2))
    assert n == len(ids) == 8 * 8 * 6 * 8 and e.num_particles() == n
    d = e.download()
    assert np.array_equal(d["id"], ids) and np.array_equal(d["x"], x)
    assert np.array_equal(d["mass"], mass) and np.array_equal(d["vol"], vol)
    assert np.array_equal(d["v"], np.tile(np.array([0.1, 0.0, -0.2], np.float32), (n, 1)))
    assert np.array_equal(d["F"], np.tile(np.eye(3, dtype=np.float32).reshape(9), (n, 1))) and not d["b"].any() and not d["ps"].any()
    if jitter == 0.0:   # the reference's own lattice (src/mpm.cpp:164-180), same order
        xr, mr, vr = scenes.lattice_block(res, lo, hi)
        assert np.abs(x - xr).max() <= 1e-7 and np.array_equal(mass, mr)
    # seeded and uploaded engines then step identically (same storage order, same ids)
    st = scenes.make_state(x, mass, vol, scenes.MAT_SAND, v0=(0.1, 0.0, -0.2))
    e2 = _engine(res)
    e2.upload(st["x"], st["v"], st["mass"], st["vol"], st["F"], st["b"], st["ps"], st["group"])
    e.substep(5)
    e2.substep(5)
    a, b = e.download(), e2.download()
    for k in ("id", "x", "v", "F", "ps"):
        assert np.array_equal(a[k], b[k]), k
    e.close()
    e2.close()


def test_boundary_band_is_not_seeded_and_slabs_partition_the_lattice():
    res, lo, hi = 32, (5, 8, 6), (12, 12, 26)          # reaches into the 7-cell band in x and z
    ids, x, mass, vol = scenes.lattice_block_hashed(res, lo, hi, jitter=0.1, seed=5)
    assert 0 < len(ids) < 7 * 4 * 20 * 8
    e = _engine(res)
    assert e.seed_lattice(lo, hi, float(vol[0]), float(mass[0]), jitter=0.1, seed=5) == len(ids)
    d = e.download()
    assert np.array_equal(d["id"], ids) and np.array_equal(d["x"], x)
    e.close()
    # two z-slabs: every particle is created by exactly one rank, with the same id and position
    cut = 4
    parts = []
    for rank, (z0, z1) in enumerate(((0, cut), (cut, 10))):
        s = _engine(res, rank=rank, world=2, tile_z0=z0, tile_z1=z1, migrate_capacity=1024)
        s.seed_lattice(lo, hi, float(vol[0]), float(mass[0]), jitter=0.1, seed=5)
        parts.append(s.download())
        s.close()
    both = {k: np.concatenate([p[k] for p in parts]) for k in ("id", "x")}
    o = np.argsort(both["id"], kind="stable")
    assert np.array_equal(both["id"][o], ids) and np.array_equal(both["x"][o], x)
    assert len(parts[0]["id"]) > 0 and len(parts[1]["id"]) > 0


@pytest.mark.parametrize("friction", [0.4, -1.0])
def test_levelset_shapes_rasterised_on_the_device_match_the_host_twin(friction):
    """mpmb_set_levelset_shapes (plane + sphere + inside-out cuboid container) against its numpy twin scenes.shapes_sdf
    fed to the fp64 oracle as a dense level set: same node velocities after the boundary projection, same particles."""
    from taichi_mpm_b200 import capi
    res = 32
    scene, st = T.perturbed_scene(scenes.MAT_SAND, res=res, cells=8, seed=9, vel=1.0, friction=friction)
    shapes = [(capi.SHAPE_PLANE, False, [0.0, 1.0, 0.0, -9.6]),
              (capi.SHAPE_SPHERE, False, [16.0, 6.0, 16.0, 5.2]),                       # a ball poking into the block from below
              (capi.SHAPE_CUBOID, True, [10.3, 8.0, 10.5, 21.4, 24.0, 21.6])]           # a container: walls inside the block's edge cells
    scene = dict(scene, planes=None, shapes=shapes, sdf=scenes.shapes_sdf(res, shapes))
    band = (scene["sdf"][..., 3] >= -3.0) & (scene["sdf"][..., 3] <= 0.0)
    assert band.sum() > 2000                                                            # the boundary band really crosses the scene
    e = T.make_engine(scene, st)
    err, got, ref = T.compare_substep(e, scene, st)
    e.close()
    print("shapes friction %g:" % friction, {k: float(v) for k, v in err.items() if not isinstance(v, bool)})
    assert err["alive_match"]
    assert err["grid_vel"] <= 1e-4 and err["v"] <= T.TOL_V_REL and err["x"] <= T.TOL_X_ABS and err["F"] <= T.TOL_F_ABS


def test_mirror_levelset_verbs_reach_the_engine_in_grid_units():
    from taichi_mpm_b200 import MPM, capi
    m = MPM(res=(32, 32, 32), base_delta_t=2e-5)
    ls = m.create_levelset()
    ls.add_plane((0, 1, 0), -0.3)
    ls.add_sphere((0.5, 0.55, 0.5), 0.3, True)          # scripts/mls-cpic/sand_stir.py:9
    ls.add_cuboid((0, 0.2, 0.05), (0.95, 0.95, 0.95), True)   # scripts/async/sand.py:35
    ls.set_friction(0.5)
    g = ls.shapes_grid_units()
    assert [s[0] for s in g] == [capi.SHAPE_PLANE, capi.SHAPE_SPHERE, capi.SHAPE_CUBOID]
    assert np.allclose(g[0][2], [0, 1, 0, -0.3 * 32]) and np.allclose(g[1][2], [16, 17.6, 16, 9.6]) and g[1][1] and g[2][1]
    assert np.allclose(g[2][2], [0, 6.4, 1.6, 30.4, 30.4, 30.4])
    m.set_levelset(ls, False)                            # rasterised on the device (mpmb_set_levelset_shapes)
    m.add_particles(type="sand", benchmark_block=((12, 12, 12), (16, 16, 16)), density=400)
    m.step(1e-4)
    assert m.num_particles() == 4 * 4 * 4 * 8
