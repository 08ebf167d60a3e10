"""The reference-facing mirror (taichi_mpm_b200.mpm.MPM: same verbs and keywords as the reference's
Python driver) against the engine driven directly, and against the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_mirror_runs_a_sand_scene_like_the_reference_script():
    from oracle import pyoracle as O
    from taichi_mpm_b200 import MPM, scenes
    res = 48
    mpm = MPM(res=(res, res, res), base_delta_t=2e-5, gravity=(0, -10, 0), frame_dt=4e-4)
    ls = mpm.create_levelset()
    ls.add_plane((0, 1, 0), -10.0 / res)      # world units, like levelset.add_plane in sand_sweep.py:14
    ls.set_friction(0.4)
    mpm.set_levelset(ls, False)
    mpm.add_particles(type="sand", benchmark_block=((16, 10, 16), (28, 22, 28)), density=400, jitter=0.05)
    mpm.add_particles(type="water", benchmark_block=((30, 12, 30), (36, 18, 36)), density=400, k=1e4)
    n = mpm.num_particles()
    assert n == (12 ** 3 + 6 ** 3) * 8
    p0 = mpm.get_particles()
    mpm.step(4e-4)                            # 19 substeps: while (t + dt < request_t), src/mpm.cpp:435
    assert mpm.substep_counter == 19
    mpm.step(-1.0)                            # dt < 0: exactly one substep (src/mpm.cpp:429-432)
    assert mpm.substep_counter == 20
    p = mpm.get_particles()
    assert len(p["id"]) == n and set(np.unique(p["group"])) == {0, 1}
    # same scene through the CPU oracle fast path
    planes = np.array([[0.0, 1.0, 0.0, -10.0]], np.float32)
    scene = dict(res=(res,) * 3, dx=1.0 / res, dt=2e-5, gravity=(0.0, -10.0, 0.0), particle_gravity=1,
                 mat_kind=np.array([scenes.MAT_SAND, scenes.MAT_WATER], np.int32),
                 mat_params=np.stack([scenes.material_params(scenes.MAT_SAND), scenes.material_params(scenes.MAT_WATER, k=1e4)]),
                 sdf=scenes.planes_sdf(res, planes), friction=0.4)
    st = dict(x=p0["x"], v=p0["v"], F=p0["F"], b=p0["b"], mass=p0["mass"], vol=p0["vol"], ps=p0["ps"], group=p0["group"],
              alive=np.ones(n, np.uint8))
    fast = O.FastOracle(scene, st, threads=4)
    fast.substeps(20)
    ids = p["id"].astype(np.int64)
    assert np.abs(p["x"] - fast.st["x"][ids]).max() < 2e-6
    # resting sand: velocities of 4e-3 m/s; the fp32 CPU path and the GPU differ by ~1e-5 m/s there
    assert np.abs(p["v"] - fast.st["v"][ids]).max() < 1e-4
    # unsupported reference features are rejected, not ignored
    with pytest.raises(ValueError):
        mpm.add_particles(type="rigid", density=40)
    with pytest.raises(ValueError):
        MPM(res=(res, res, res), optimized=False)
