"""Shared helpers for the parity tests: seeded scenes for both the oracle and the engine."""
import numpy as np

from taichi_mpm_b200 import scenes

# Parity tolerances, single substep, GPU fp32 vs fp64 oracle on identical inputs.
TOL_GRID_REL = 2e-5     # |d(p,m)| <= TOL * max(|p|_inf, m) (maxima over the grid)
TOL_V_REL = 1e-4        # particle v and apic_b: relative to max|v| (resp. max|b|)
TOL_F_ABS = 2e-5
TOL_X_ABS = 1e-6        # times the domain size (=1)
TOL_PS_ABS = 1e-5


# per-kind arguments of perturbed_scene for the single-substep comparisons: von Mises at a strain level that puts particles
# on both sides of its yield surface (|dev eps|^2 vs yield_stress / 2 mu = 2.8e-4), visco with hardening (visco_tau moves)
KIND_KW = {scenes.MAT_VON_MISES: dict(strain=0.006), scenes.MAT_VISCO: dict(kappa=0.3)}


def ps_err(a, b):
    """Error of the plastic scalar: absolute up to 1, relative above (visco_tau is O(1e3), the others O(1))."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float((np.abs(a - b) / np.maximum(1.0, np.abs(b))).max()) if a.size else 0.0


def perturb_state(st, kind, dx, seed=0, strain=0.02, vel=0.5):
    """Random affine velocity field, random apic_b, random F perturbation and plastic scalars on top of a
    lattice state (in place): every term of P2G, the grid update and G2P + return map is exercised."""
    rng = np.random.default_rng(seed)
    x = st["x"]
    n = len(x)
    A = rng.normal(size=(3, 3)) * vel * 4
    a = rng.normal(size=3) * vel
    st["v"] = (a + (x - x.mean(0)) @ A.T + rng.normal(size=(n, 3)) * 0.05 * vel).astype(np.float32)
    st["b"] = (rng.normal(size=(n, 9)) * vel * dx * 0.1).astype(np.float32)
    if kind != scenes.MAT_WATER:
        G = rng.normal(size=(n, 9)) * strain
        st["F"] = (st["F"] + G).astype(np.float32)
    if kind == scenes.MAT_SNOW:
        st["ps"] = (1.0 + rng.normal(size=n) * 0.02).astype(np.float32)
    if kind == scenes.MAT_WATER:
        st["ps"] = (1.0 + rng.normal(size=n) * 0.01).astype(np.float32)
    if kind == scenes.MAT_SAND:
        st["ps"] = (np.abs(rng.normal(size=n)) * 1e-3 * (rng.random(n) < 0.3)).astype(np.float32)
    if kind == scenes.MAT_VISCO:   # visco_tau around the first Piola norms of this strain level: some particles flow, some do not
        st["ps"] = (1000.0 * (0.2 + 1.6 * rng.random(n))).astype(np.float32)
    return st


def perturbed_scene(kind, res=32, cells=8, seed=0, strain=0.02, vel=0.5, with_floor=True, friction=0.4, dt=None, **matkw):
    """A small block with random affine velocity field, random apic_b and random F perturbation."""
    lo = np.array([(res - cells) // 2, 9, (res - cells) // 2])
    hi = lo + cells
    x, mass, vol = scenes.lattice_block(res, lo, hi, jitter=0.2, seed=seed + 1)
    st = scenes.make_state(x, mass, vol, kind)
    dx = 1.0 / res
    perturb_state(st, kind, dx, seed, strain, vel)
    if dt is None:
        dt = {scenes.MAT_SAND: 2e-5, scenes.MAT_WATER: 5e-5}.get(kind, 1e-4)
    planes = np.array([[0.0, 1.0, 0.0, -(lo[1] + 0.6)]], np.float32) if with_floor else None
    scene = dict(res=(res, res, res), dx=dx, dt=dt, gravity=(0.0, -10.0, 0.0), particle_gravity=1,
                 mat_kind=np.array([kind], np.int32), mat_params=scenes.material_params(kind, **matkw)[None],
                 planes=planes, friction=friction if with_floor else 0.0)
    scene["sdf"] = scenes.planes_sdf(res, planes) if with_floor else None
    return scene, st


def make_engine(scene, state, **kw):
    from taichi_mpm_b200 import capi
    e = capi.Engine(scene["res"], scene["dx"], scene["dt"], scene["gravity"], scene.get("particle_gravity", 1),
                    kw.pop("clean_boundary", True), **kw)
    for g, (k, p) in enumerate(zip(scene["mat_kind"], scene["mat_params"])):
        e.set_material(g, int(k), p)
    if scene.get("shapes") is not None:
        e.set_levelset_shapes(scene["shapes"], scene["friction"])
    elif scene.get("sdf_dense_upload") is not None:
        e.set_sdf(scene["sdf_dense_upload"], scene["friction"])
    elif scene.get("planes") is not None:
        e.set_planes(scene["planes"], scene["friction"])
    e.upload(state["x"], state["v"], state["mass"], state["vol"], state["F"], state["b"], state["ps"], state["group"])
    return e


def compare_substep(e, scene, state, check_grid=True):
    """Runs one substep on the engine stage by stage and on the fp64 oracle; returns dict of errors."""
    from oracle import pyoracle as O
    ref, grid_rast, grid_vel = O.substep(scene, state, np.float64)
    out = {}
    e.sort_particles_and_populate_grid()
    e.rasterize()
    if check_grid:
        g0 = e.download_grid(0).astype(np.float64)
        g1 = e.download_grid(1).astype(np.float64)
        pmax = max(np.abs(grid_rast[..., :3]).max(), grid_rast[..., 3].max())
        out["grid_rast"] = np.abs(g0 - grid_rast).max() / pmax
        active = grid_rast[..., 3] > 0
        vmax = np.abs(grid_vel[..., :3]).max()
        out["grid_vel"] = np.abs(g1[..., :3] - grid_vel[..., :3])[active].max() / vmax
        out["grid_mass_outside"] = np.abs(g0[~active]).max() if (~active).any() else 0.0
    e.resample()
    got = e.download()
    alive = ref["alive"].astype(bool)
    ids = got["id"].astype(np.int64)
    out["alive_match"] = (len(ids) == alive.sum()) and np.array_equal(np.sort(ids), np.nonzero(alive)[0])
    sel = ids
    vmax = np.abs(ref["v"][alive]).max()
    bmax = max(np.abs(ref["b"][alive]).max(), 1e-30)
    out["x"] = np.abs(got["x"] - ref["x"][sel]).max()
    out["v"] = np.abs(got["v"] - ref["v"][sel]).max() / vmax
    out["b"] = np.abs(got["b"] - ref["b"][sel]).max() / bmax
    out["F"] = np.abs(got["F"] - ref["F"][sel]).max()
    out["ps"] = (np.abs(got["ps"] - ref["ps"][sel]) / np.maximum(1.0, np.abs(ref["ps"][sel]))).max()   # relative above 1 (visco_tau ~ 1e3)
    out["mass"] = np.abs(got["mass"] - ref["mass"][sel]).max()
    return out, got, ref
