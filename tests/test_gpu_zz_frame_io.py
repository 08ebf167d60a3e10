"""Frame output and snapshots of the mirror on the real engine: the dumped file
holds exactly what the device holds, and a reloaded snapshot continues like the original run."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _make(tmp_path, sub):
    from taichi_mpm_b200 import MPM
    res = 48
    m = MPM(res=(res, res, res), base_delta_t=2e-5, gravity=(0, -10, 0), frame_directory=str(tmp_path / sub), verbose_bgeo=True)
    ls = m.create_levelset()
    ls.add_plane((0, 1, 0), -10.0 / res)
    ls.set_friction(0.4)
    m.set_levelset(ls, False)
    return m


def test_visualize_and_snapshot_round_trip_on_device(tmp_path):
    from taichi_mpm_b200 import bgeo
    m = _make(tmp_path, "frames")
    m.add_particles(type="sand", benchmark_block=((16, 10, 16), (26, 20, 26)), density=400, jitter=0.05)
    m.add_particles(type="water", benchmark_block=((30, 12, 30), (34, 16, 34)), density=400, k=1e4)
    m.step(2e-4)
    fn = m.visualize()
    assert fn.endswith("0001.bgeo") and m.frame_count == 1
    p = m.get_particles()
    pos, attrs = bgeo.read_bgeo(fn)
    d = {a[0]: a[2] for a in attrs}
    assert np.array_equal(pos, p["x"]) and np.array_equal(d["v"], p["v"])
    assert np.array_equal(d["index"].ravel(), p["id"].astype(np.int32))
    assert np.array_equal(d["m"].ravel(), p["mass"])
    # snapshot -> fresh solver -> identical resident state
    snap = str(tmp_path / "snap.npz")
    m.general_action(action="save", file_name=snap)
    m2 = _make(tmp_path, "frames2")
    m2.general_action(action="load", file_name=snap)
    q = m2.get_particles()
    for k in ("id", "x", "v", "F", "b", "mass", "vol", "ps", "group"):
        assert np.array_equal(p[k], q[k]), k
    assert m2.substep_counter == m.substep_counter and m2.current_t == m.current_t
    # both continue: same particles, same motion up to the summation order of a re-sorted upload
    m.step(2e-4)
    m2.step(2e-4)
    a, b = m.get_particles(), m2.get_particles()
    assert np.array_equal(a["id"], b["id"])
    assert np.abs(a["x"] - b["x"]).max() < 1e-6
    assert np.abs(a["v"] - b["v"]).max() < 1e-4
    with pytest.raises(ValueError):
        m.general_action(action="no_such_action")


def test_device_packed_bgeo_equals_the_host_writer_and_material_change_refreshes_the_stress(tmp_path):
    """(a) The non-verbose frame whose point records are packed on the device (mpmb_download_bgeo_points) is byte-identical
    to the host writer — which tests/test_bgeo.py pins to the reference's own Partio output — also after deletions.
    (b) mpmb_set_material on resident particles rebuilds the cached affine matrices: changing the material after the
    upload gives the same substep as uploading with the new material."""
    import filecmp
    from taichi_mpm_b200 import MPM, bgeo, scenes
    from tests import common as T
    res = 48
    m = MPM(res=(res, res, res), base_delta_t=2e-5, gravity=(0, -10, 0), frame_directory=str(tmp_path / "f"))
    m.add_particles(type="sand", benchmark_block=((8, 8, 16), (20, 20, 26)), density=400, jitter=0.05, initial_velocity=(-16.0, 0.0, 0.5))
    m.step(2e-3)    # 100 substeps: the block drifts into the x < 7 cells band, some particles are deleted
    fn = m.visualize()
    p = m.get_particles()
    assert 0 < len(p["id"]) < 12 * 12 * 10 * 8
    ref = str(tmp_path / "host.bgeo")
    bgeo.write_bgeo(ref, p["x"], bgeo.reference_attributes(p, verbose=False))
    assert filecmp.cmp(fn, ref, shallow=False)

    scene, st = T.perturbed_scene(scenes.MAT_JELLY, res=32, cells=6, seed=4)
    new = scenes.material_params(scenes.MAT_JELLY, E=3e5, nu=0.2)
    e1 = T.make_engine(scene, st)
    e1.set_material(0, scenes.MAT_JELLY, new)          # after the upload
    e1.substep(1)
    scene2 = dict(scene, mat_params=new[None])
    e2 = T.make_engine(scene2, st)                      # before the upload
    e2.substep(1)
    a, b = e1.download(), e2.download()
    for k in ("x", "v", "F"):
        assert np.array_equal(a[k], b[k]), k
    e0 = T.make_engine(scene, st)
    e0.substep(1)
    assert np.abs(e0.download()["v"] - a["v"]).max() > 1e-4   # the material does matter in this scene
