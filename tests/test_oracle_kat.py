"""Known-answer tests that pin the CPU oracle.

(1) The reference's own unit tests for this path, restated against the oracle:
    mpm_kernel (src/tests.cpp:10-33), mpm_fast_kernel32 (src/tests.cpp:35-51),
    mls_kernel (src/transfer.cpp:975-989), grid_pos_offset (src/transfer.cpp:353-359).
    These are the ONLY fixtures the reference holds for the path; transfers and constitutive
    models are "parity unpinned" (DESIGN.md).
(2) Convention-free KATs derived from the algorithm.
"""
import numpy as np
import pytest

from oracle import pyoracle as O
from taichi_mpm_b200 import scenes
from tests import common as T


# ------------------------------------------------------------------ (1) reference tests restated
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_mpm_kernel_partition_of_unity(dtype):
    rng = np.random.default_rng(1)
    for _ in range(100):
        pos = rng.random(3) * 10.0
        for x in pos:
            w, dw = O.quadratic_kernel(x, dtype)
            assert w.sum() == pytest.approx(1.0, rel=1e-5)
            assert abs(dw.sum()) < 1e-6
            w, dw = O.cubic_kernel(x, dtype)
            assert w.sum() == pytest.approx(1.0, rel=1e-5)
            assert abs(dw.sum()) < 2e-6


def test_fast_kernel_equals_slow_kernel():
    # mpm_fast_kernel32 / mls_kernel: product-form 27 weights == per-axis kernel products, 1e-6
    rng = np.random.default_rng(2)
    for _ in range(10000 // 20):
        pos = rng.random(3).astype(np.float32) + 0.5       # [0.5,1.5)^3 like src/transfer.cpp:978
        fast = O.mls_fast_kernel(pos, np.float32)
        ws = [O.quadratic_kernel(float(p) + 7.0, np.float32)[0] for p in pos]   # fract(pos-0.5) form
        slow = np.einsum("i,j,k->ijk", *ws)
        assert np.abs(fast - slow).max() < 1e-6


def test_grid_pos_offset_table():
    # stencil node n <-> (n/9, n/3%3, n%3): the layout both transfers rely on
    w = O.mls_fast_kernel([0.6, 0.9, 1.3], np.float64)
    wx = O.quadratic_kernel(0.6 + 3, np.float64)[0]
    wy = O.quadratic_kernel(0.9 + 3, np.float64)[0]
    wz = O.quadratic_kernel(1.3 + 3, np.float64)[0]
    flat = w.reshape(27)
    for n in range(27):
        assert flat[n] == pytest.approx(wx[n // 9] * wy[n // 3 % 3] * wz[n % 3], rel=1e-12)


# ------------------------------------------------------------------ (2) derived KATs
def test_weight_moments():
    # sum w (x_i - x_p) = 0 ; sum w (x_i-x_p)(x_i-x_p)^T = 1/4 I (grid units) -> inv_D = 4 (src/kernel.h:68-70)
    rng = np.random.default_rng(3)
    for _ in range(50):
        rel = rng.random(3) + 0.5
        w = O.mls_fast_kernel(rel, np.float64)
        idx = np.stack(np.meshgrid(range(3), range(3), range(3), indexing="ij"), -1).astype(np.float64)
        d = idx - rel
        assert abs(w.sum() - 1) < 1e-12
        assert np.abs(np.einsum("ijk,ijkc->c", w, d)).max() < 1e-12
        M = np.einsum("ijk,ijkc,ijkd->cd", w, d, d)
        assert np.abs(M - 0.25 * np.eye(3)).max() < 1e-12


def test_svd_and_polar():
    rng = np.random.default_rng(4)
    for dtype, tol in ((np.float64, 1e-13), (np.float32, 2e-6)):
        for _ in range(200):
            A = rng.normal(size=(3, 3))
            if np.linalg.det(A) < 0:
                A[:, 0] *= -1
            U, s, V = O.svd3(A, dtype)
            assert np.abs(U @ np.diag(s) @ V.T - A).max() < tol * 10
            assert np.abs(U.T @ U - np.eye(3)).max() < tol * 10
            assert np.abs(np.sort(s) - np.sort(np.linalg.svd(A, compute_uv=False))).max() < tol * 10
            R, S = O.polar3(A, dtype)
            assert np.abs(R @ S - A).max() < tol * 10
            assert np.abs(S - S.T).max() < tol * 10
            assert np.linalg.det(R.astype(np.float64)) == pytest.approx(1.0, abs=1e-5)


MATS = [(scenes.MAT_LINEAR, {}), (scenes.MAT_JELLY, {}), (scenes.MAT_SNOW, {}), (scenes.MAT_SAND, {})]


@pytest.mark.parametrize("kind,kw", MATS)
def test_zero_stress_at_identity(kind, kw):
    p = scenes.material_params(kind, **kw)
    f = O.calculate_force(kind, p, np.eye(3), scenes.default_scalar(kind), 1e-6)
    assert np.abs(f).max() < 1e-12


@pytest.mark.parametrize("kind,kw", [(scenes.MAT_JELLY, {}), (scenes.MAT_SNOW, {}), (scenes.MAT_SAND, {})])
def test_rotation_covariance(kind, kw):
    # F -> QF  =>  P F^T -> Q (P F^T) Q^T  (isotropic models)
    rng = np.random.default_rng(5)
    p = scenes.material_params(kind, **kw)
    for _ in range(20):
        F = np.eye(3) + rng.normal(size=(3, 3)) * 0.05
        Q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        if np.linalg.det(Q) < 0:
            Q[:, 0] *= -1
        a = O.calculate_force(kind, p, F, scenes.default_scalar(kind), 1e-6)
        b = O.calculate_force(kind, p, Q @ F, scenes.default_scalar(kind), 1e-6)
        assert np.abs(b - Q @ a @ Q.T).max() < 1e-9 * max(1.0, np.abs(a).max() * 1e6)


def test_water_closed_form():
    p = scenes.material_params(scenes.MAT_WATER)
    j, vol = 0.97, 2e-7
    f = O.calculate_force(scenes.MAT_WATER, p, np.eye(3), j, vol)
    pres = 1e4 * (j ** -7.0 - 1)
    assert np.allclose(f, vol * j * pres * np.eye(3), rtol=1e-12)
    cdg = np.eye(3) + np.diag([1e-3, -2e-3, 5e-4])
    _, j2 = O.plasticity(scenes.MAT_WATER, p, cdg, np.eye(3), j)
    assert j2 == pytest.approx(j * (1 + 1e-3 - 2e-3 + 5e-4), rel=1e-12)
    _, j3 = O.plasticity(scenes.MAT_WATER, p, np.eye(3) * 0.1, np.eye(3), 0.2)
    assert j3 == pytest.approx(0.1)


def test_snow_clamp_and_jp():
    p = scenes.material_params(scenes.MAT_SNOW)
    F = np.diag([1.05, 0.9, 1.0])
    F2, Jp = O.plasticity(scenes.MAT_SNOW, p, np.eye(3), F, 1.0)
    s = np.sort(np.linalg.svd(F2, compute_uv=False))
    tc, ts = float(p[3]), float(p[4])          # parameters are stored as fp32
    assert np.allclose(s, np.sort([1 + ts, 1 - tc, 1.0]), atol=1e-12)
    assert Jp == pytest.approx(1.05 * 0.9 / ((1 + ts) * (1 - tc)), rel=1e-12)
    # inside the box: unchanged
    F = np.diag([1.001, 0.999, 1.0])
    F2, Jp = O.plasticity(scenes.MAT_SNOW, p, np.eye(3), F, 1.3)
    assert np.allclose(F2, F, atol=1e-13) and Jp == pytest.approx(1.3)


def test_sand_return_map_cases():
    p = scenes.material_params(scenes.MAT_SAND)
    mu, lam, alpha = float(p[0]), float(p[1]), float(p[2])
    # expansion: tr >= 0 -> F' = I, logJp accumulates sum(eps)
    F = np.diag([1.01, 1.02, 1.0])
    F2, lj = O.plasticity(scenes.MAT_SAND, p, np.eye(3), F, 0.0)
    assert np.allclose(F2, np.eye(3), atol=1e-12)
    assert lj == pytest.approx(np.log(1.01) + np.log(1.02), rel=1e-10)
    # inside the cone: unchanged, logJp reset
    F = np.diag([0.99, 0.9895, 0.9905])
    F2, lj = O.plasticity(scenes.MAT_SAND, p, np.eye(3), F, 0.0)
    assert np.allclose(F2, F, atol=1e-12) and lj == 0.0
    # outside: lands on the cone  ||eps_hat'|| = -(3 lam + 2 mu)/(2 mu) tr alpha
    F = np.diag([0.97, 1.02, 0.999])
    F2, lj = O.plasticity(scenes.MAT_SAND, p, np.eye(3), F, 0.0)
    eps = np.log(np.diag(F2))
    tr = np.log(np.diag(F)).sum()
    hat = eps - eps.sum() / 3
    assert eps.sum() == pytest.approx(tr, rel=1e-9)     # return map is deviatoric
    assert np.linalg.norm(hat) == pytest.approx(-(3 * lam + 2 * mu) / (2 * mu) * tr * alpha, rel=1e-9)


def test_friction_project_cases():
    n = np.array([0.0, 1.0, 0.0])
    v = np.array([1.0, -2.0, 0.5])
    assert np.allclose(O.friction_project(v, np.zeros(3), n, -1.0), 0)                       # sticky
    assert np.allclose(O.friction_project(v, np.zeros(3), n, -2.0), [1.0, 0.0, 0.5])         # slip, no friction
    out = O.friction_project(v, np.zeros(3), n, 0.4)                                         # separate + Coulomb
    tn = np.hypot(1.0, 0.5)
    assert np.allclose(out, np.array([1.0, 0, 0.5]) * (tn - 2 * 0.4) / tn)
    out = O.friction_project(np.array([1.0, 2.0, 0.5]), np.zeros(3), n, 0.4)                 # leaving: untouched
    assert np.allclose(out, [1.0, 2.0, 0.5])


def _small_block(kind, res=24, cells=4, **kw):
    lo = np.array([(res - cells) // 2] * 3)
    x, mass, vol = scenes.lattice_block(res, lo, lo + cells, jitter=0.1, seed=7)
    st = scenes.make_state(x, mass, vol, kind)
    scene = dict(res=(res,) * 3, dx=1.0 / res, dt=1e-4, gravity=(0.0, 0.0, 0.0), particle_gravity=1,
                 mat_kind=np.array([kind], np.int32), mat_params=scenes.material_params(kind, **kw)[None], sdf=None, friction=0.0)
    return scene, st


def test_p2g_conserves_mass_and_momentum():
    scene, st = _small_block(scenes.MAT_JELLY)
    rng = np.random.default_rng(8)
    st["v"] = rng.normal(size=st["v"].shape).astype(np.float32)
    _, grid_rast, _ = O.substep(scene, st, np.float64)
    assert grid_rast[..., 3].sum() == pytest.approx(st["mass"].astype(np.float64).sum(), rel=1e-12)
    mom = (st["mass"][:, None].astype(np.float64) * st["v"].astype(np.float64)).sum(0)
    assert np.allclose(grid_rast[..., :3].sum((0, 1, 2)), mom, rtol=1e-10, atol=1e-18)


def test_affine_reproduction():
    # v_p = a + A x_p with apic_b consistent (b = -dx/4 * A... C = -4/dx b), F=I, no stress:
    # P2G -> normalise -> G2P returns v_p and C = A exactly (quadratic B-splines reproduce affine fields)
    scene, st = _small_block(scenes.MAT_JELLY, E=0.0)
    rng = np.random.default_rng(9)
    A = rng.normal(size=(3, 3))
    a = rng.normal(size=3)
    x = st["x"].astype(np.float64)
    # interior particles only see a full neighbourhood if the block is surrounded; use a uniform field instead:
    # every node then gets exactly a + A x_i, so G2P reproduces the field for ALL particles.
    st["v"] = (a + x @ A.T).astype(np.float64)
    dx = scene["dx"]
    C = A
    b = (-dx / 4.0) * C                    # column-major flatten: b[c*3+r] = B[r,c]
    st["b"] = np.tile(b.T.reshape(9), (len(x), 1))
    new, _, grid_vel = O.substep(scene, st, np.float64)
    assert np.abs(new["v"] - st["v"]).max() < 1e-9
    Bn = new["b"].reshape(-1, 3, 3).transpose(0, 2, 1)       # back to math layout
    assert np.abs(-4.0 / dx * Bn - A).max() < 1e-8
    # F' = (I + dt C) F
    Fn = new["F"].reshape(-1, 3, 3).transpose(0, 2, 1)
    assert np.abs(Fn - (np.eye(3) + scene["dt"] * A)).max() < 1e-10


def test_fast_path_matches_scalar_oracle():
    # the OpenMP tile-cache / 8-colour restatement == the scalar restatement (same fp32 arithmetic,
    # same summation order inside a block up to the block ordering)
    for kind in (scenes.MAT_JELLY, scenes.MAT_SAND, scenes.MAT_SNOW, scenes.MAT_WATER, scenes.MAT_LINEAR):
        scene, st = _small_block(kind, res=40, cells=6)
        scene["gravity"] = (0.0, -10.0, 0.0)
        scene["sdf"] = scenes.floor_sdf(40, 17.5)
        scene["friction"] = 0.4
        rng = np.random.default_rng(10)
        st["v"] = (rng.normal(size=st["v"].shape) * 0.3).astype(np.float32)
        if kind != scenes.MAT_WATER:
            st["F"] = (st["F"] + rng.normal(size=st["F"].shape) * 0.01).astype(np.float32)
        ref, _, grid_vel = O.substep(scene, st, np.float32)
        fast = O.FastOracle(scene, st, threads=2)
        upd, t = fast.substeps(1)
        assert upd == len(st["mass"])
        g = fast.download_grid()
        assert np.abs(g - grid_vel).max() <= 2e-5 * np.abs(grid_vel).max()
        for k in ("x", "v", "F", "b", "ps"):
            scale = max(np.abs(ref[k]).max(), 1e-30)
            assert np.abs(fast.st[k] - ref[k]).max() <= 5e-5 * scale, (kind, k)
        assert np.array_equal(fast.st["alive"], ref["alive"])


def test_fast_path_storage_reorder_keeps_caller_indexing():
    # sort_allocator (src/mpm.cpp:753-768): physically re-ordered storage changes neither the
    # caller-visible indexing nor (beyond fp32 summation order) the result; deleted particles stay deleted
    scene, st = T.perturbed_scene(scenes.MAT_SAND, res=32, cells=8, seed=3)
    res = scene["res"][0]
    st["x"][5] = [6.5 / res, 0.5, 0.5]        # inside the deletion band
    a = O.FastOracle(scene, st, threads=2)
    b = O.FastOracle(scene, st, threads=2, reorder_interval=5)
    na = nb = 0
    for _ in range(3):                         # several calls: gather/scatter through `origin` each time
        na += a.substeps(7)[0]
        nb += b.substeps(7)[0]
    assert na == nb
    assert np.array_equal(a.st["alive"], b.st["alive"]) and a.st["alive"][5] == 0
    live = a.st["alive"] > 0
    for k in ("x", "v", "F", "b", "ps"):
        scale = max(np.abs(a.st[k][live]).max(), 1e-30)
        assert np.abs(a.st[k][live] - b.st[k][live]).max() <= 2e-5 * scale, k


def test_clear_boundary_band():
    scene, st = _small_block(scenes.MAT_JELLY)
    res = scene["res"][0]
    st["x"][0] = [6.9 / res, 0.5, 0.5]        # inside the 7-cell band -> deleted
    st["x"][1] = [0.5, (res - 6.9) / res, 0.5]
    st["x"][2] = [7.5 / res, 0.5, 0.5]        # just outside the band -> kept
    st["v"][3] = [np.nan, 0, 0]               # abnormal -> deleted (but only detected after G2P overwrote v)
    new, _, _ = O.substep(scene, st, np.float64)
    assert new["alive"][0] == 0 and new["alive"][1] == 0 and new["alive"][2] == 1


def test_mpm88_2d_config1():
    # config 1: the 88-line 2D algorithm, elastic jelly (plastic=False): invariants of one step
    rng = np.random.default_rng(11)
    n = 80
    centres = [(0.55, 0.45), (0.45, 0.65), (0.55, 0.85)]
    x = np.concatenate([(rng.random((1000, 2)) * 2 - 1) * 0.08 + c for c in centres])
    N = len(x)
    v = np.zeros((N, 2))
    F = np.tile([1.0, 0, 0, 1.0], (N, 1))
    C = np.zeros((N, 4))
    Jp = np.ones(N)
    x1, v1, F1, C1, Jp1, grid = O.mpm88_advance(n, 1e-4, x, v, F, C, Jp, plastic=False)
    # free fall of an unstressed body: every particle gains exactly g*dt, no deformation
    assert np.allclose(v1[:, 1], -200 * 1e-4, atol=1e-12) and np.allclose(v1[:, 0], 0, atol=1e-12)
    assert np.allclose(F1, F, atol=1e-12) and np.allclose(Jp1, 1.0)
    assert np.allclose(x1, x + 1e-4 * v1)
    # 200 steps: bodies hit the floor, stay in the box, momentum changes only through the boundary
    for _ in range(200):
        x1, v1, F1, C1, Jp1, grid = O.mpm88_advance(n, 1e-4, x1, v1, F1, C1, Jp1, plastic=False)
    assert np.isfinite(x1).all() and x1.min() > 0.03 and x1.max() < 0.97
    J = F1[:, 0] * F1[:, 3] - F1[:, 1] * F1[:, 2]
    assert (J > 0.5).all() and (J < 1.5).all()
    # fp32 path stays close to fp64 over a short horizon
    xs, vs, Fs, Cs, Js, _ = O.mpm88_advance(n, 1e-4, x, v, F, C, Jp, plastic=False, dtype=np.float32)
    xd, vd, Fd, Cd, Jd, _ = O.mpm88_advance(n, 1e-4, x, v, F, C, Jp, plastic=False, dtype=np.float64)
    assert np.abs(xs - xd).max() < 1e-6 and np.abs(vs - vd).max() < 1e-5


def test_fast_path_is_finite_for_resting_sand_next_to_water():
    # regression: F = I exactly made the fast SVD's pivot underflow (NaN) — sand at rest is the
    # headline scene's initial state
    res = 32
    xa, ma, va = scenes.lattice_block(res, (10, 10, 10), (16, 16, 16), 400.0, 0.05)
    xb, mb, vb = scenes.lattice_block(res, (18, 11, 18), (22, 15, 22), 400.0, 0.0)
    sa = scenes.make_state(xa, ma, va, scenes.MAT_SAND, 0)
    sb = scenes.make_state(xb, mb, vb, scenes.MAT_WATER, 1)
    st = {k: np.concatenate([sa[k], sb[k]]) for k in sa}
    planes = np.array([[0, 1, 0, -10.0]], np.float32)
    scene = dict(res=(res,) * 3, dx=1.0 / res, dt=2e-5, gravity=(0.0, -10.0, 0.0), particle_gravity=1,
                 mat_kind=np.array([scenes.MAT_SAND, scenes.MAT_WATER], np.int32),
                 mat_params=np.stack([scenes.material_params(scenes.MAT_SAND), scenes.material_params(scenes.MAT_WATER)]),
                 sdf=scenes.planes_sdf(res, planes), friction=0.4)
    f = O.FastOracle(scene, st, threads=2)
    f.substeps(10)
    ref = st
    for _ in range(10):
        ref, _, _ = O.substep(scene, ref, np.float32)
    for k in ("x", "v", "F"):
        assert np.isfinite(f.st[k]).all(), k
        assert np.abs(f.st[k] - ref[k]).max() <= 2e-3 * max(np.abs(ref[k]).max(), 1e-6), k   # different fp32 SVDs, tiny velocities
