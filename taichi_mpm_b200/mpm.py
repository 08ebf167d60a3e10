"""Host-side mirror of the reference's solver surface for the accelerated path.

Same verbs and keyword names as the reference's Python driver (`tc.dynamics.MPM`, mirrored in the
reference tree by scripts/async/async_mpm.py:17-300) and the `MPM<3>` plugin behind it
(src/mpm.cpp:27-75 initialize, 77-270 add_particles, 428-450 step, 452-575 substep):

    mpm = MPM(res=(256,256,256), base_delta_t=2e-5, gravity=(0,-10,0))
    ls = mpm.create_levelset(); ls.add_plane((0,1,0), -0.1); ls.set_friction(0.4); mpm.set_levelset(ls)
    mpm.add_particles(type='sand', benchmark_block=(lo, hi), density=400)
    mpm.step(frame_dt)            # runs substeps on the GPU engine through the C-ABI
    p = mpm.get_particles()       # what visualize() would dump (src/visualize.cpp:16-100)

Everything numerical happens in libmpmb.so (capi.Engine); this file is glue: kwargs -> MpmbConfig,
registered particle type names -> material groups, level-set planes -> the engine's dense SDF.
Only the fast path of the reference is mirrored (optimized=True, no rigid bodies, static level set).
"""
import numpy as np

from . import capi, scenes


class LevelSet:
    """Static level set made of half-spaces (tc.core LevelSet3D.add_plane / set_friction as used by
    scripts/mls-cpic/sand_sweep.py:13-19).  World units: phi(x) = n.x + d, inside is phi>0."""

    def __init__(self, res, delta_x):
        self.res = tuple(res)
        self.delta_x = delta_x
        self.planes = []
        self.friction = 0.0
        self.dense = None

    def add_plane(self, normal, d):
        n = np.asarray(normal, np.float64)
        n = n / np.linalg.norm(n)
        self.planes.append((n[0], n[1], n[2], float(d)))

    def set_friction(self, f):
        self.friction = float(f)

    def set_dense(self, sdf4):
        """Arbitrary static level set sampled at the nodes: [nx][ny][nz][4] = (n, phi in grid units)."""
        self.dense = np.ascontiguousarray(sdf4, np.float32)

    def planes_grid_units(self):
        # phi_grid(X) = phi(x)/dx with X = x/dx  ->  n.X + d/dx
        return np.array([[p[0], p[1], p[2], p[3] / self.delta_x] for p in self.planes], np.float32)


class MPM:
    """Mirror of `tc.dynamics.MPM(**kwargs)` for 3D scenes on the accelerated path."""

    PARTICLE_TYPES = capi.MATERIAL_BY_NAME  # registered names, src/particles.cpp:845-856

    def __init__(self, **kwargs):
        res = kwargs["res"]
        if len(res) != 3:
            raise ValueError("the accelerated path is MPM<3>; 2D scenes run on the CPU reference")
        self.res = tuple(int(r) for r in res)
        self.delta_x = float(kwargs.get("delta_x", 1.0 / self.res[0]))          # async_mpm.py:40-41
        self.base_delta_t = float(kwargs.get("base_delta_t", 1e-4)) * float(kwargs.get("dt_multiplier", 1.0))  # mpm.cpp:42-43
        g = kwargs.get("gravity", (0.0, -10.0, 0.0))                            # mpm.cpp:38
        if np.isscalar(g):
            g = (0.0, float(g), 0.0)
        self.gravity = tuple(float(x) for x in g)
        self.frame_dt = float(kwargs.get("frame_dt", 0.01))
        self.particle_gravity = bool(kwargs.get("particle_gravity", True))      # mpm.cpp:47
        self.clean_boundary = bool(kwargs.get("clean_boundary", True))          # mpm.cpp:563
        if not kwargs.get("optimized", True):
            raise ValueError("optimized=False (the scalar transfers) is not accelerated")
        for key in ("apic_damping", "rpic_damping", "affine_damping", "penalty"):
            if float(kwargs.get(key, 0.0)) != 0.0:
                raise ValueError("%s != 0 is outside the accelerated fast path" % key)
        self.engine = capi.Engine(self.res, self.delta_x, self.base_delta_t, self.gravity, self.particle_gravity,
                                  self.clean_boundary, device=int(kwargs.get("device", 0)), capacity=int(kwargs.get("capacity", 0)))
        self.current_t = 0.0
        self.request_t = 0.0
        self.update_counter = 0   # "Times of particle updating" (mpm.cpp:436,449)
        self.substep_counter = 0
        self._groups = []         # (kind, params) per material group
        self._host = None         # pending host-side particle set (dict of arrays)
        self._dirty = False
        self._n_uploaded = 0

    # ---- level set
    def create_levelset(self):
        return LevelSet(self.res, self.delta_x)

    def set_levelset(self, levelset, is_dynamic_levelset=False):
        if is_dynamic_levelset:
            raise ValueError("dynamic level sets are outside the accelerated fast path")
        if levelset.dense is not None:
            self.engine.set_sdf(levelset.dense, levelset.friction)
        elif levelset.planes:
            self.engine.set_planes(levelset.planes_grid_units(), levelset.friction)
        else:
            self.engine.set_sdf(None, 0.0)

    # ---- particles
    def _group_of(self, kind, params):
        for g, (k, p) in enumerate(self._groups):
            if k == kind and np.array_equal(p, params):
                return g
        if len(self._groups) >= capi.MPMB_MAX_GROUPS:
            raise ValueError("more than %d distinct materials" % capi.MPMB_MAX_GROUPS)
        self._groups.append((kind, params))
        self.engine.set_material(len(self._groups) - 1, kind, params)
        return len(self._groups) - 1

    def add_particles(self, **kwargs):
        """type=<registered name>, plus either positions=[n,3] (world units) or
        benchmark_block=(lo_cell, hi_cell) for the reference's 8-per-cell lattice
        (src/mpm.cpp:164-180).  density (400), initial_velocity, and the material keys of the
        reference type (E, nu, youngs_modulus, friction_angle, ...) are honoured."""
        name = kwargs["type"]
        if name == "rigid":
            raise ValueError("rigid bodies are outside the accelerated fast path")
        if name not in self.PARTICLE_TYPES:
            raise ValueError("unknown particle type %r" % name)
        kind = self.PARTICLE_TYPES[name]
        params = scenes.material_params(kind, **kwargs)
        group = self._group_of(kind, params)
        density = float(kwargs.get("density", 400.0))                           # mpm.cpp:135
        if "benchmark_block" in kwargs:
            lo, hi = kwargs["benchmark_block"]
            x, mass, vol = scenes.lattice_block(self.res[0], lo, hi, density, kwargs.get("jitter", 0.0))
        else:
            x = np.ascontiguousarray(kwargs["positions"], np.float32).reshape(-1, 3)
            ppc = float(kwargs.get("maximum", 8.0))
            vol = np.full(len(x), self.delta_x ** 3 / ppc, np.float32)          # mpm.cpp:134
            mass = (vol * density).astype(np.float32)
        # near_boundary particles are ignored (mpm.cpp:129-132)
        X = x / np.float32(self.delta_x)
        keep = (X.min(1) >= 7.0) & ((X - np.asarray(self.res, np.float32)).max(1) <= -7.0)
        x, mass, vol = x[keep], mass[keep], vol[keep]
        st = scenes.make_state(x, mass, vol, kind, group, kwargs.get("initial_velocity", (0.0, 0.0, 0.0)))
        cur = self._pull_host()
        if cur is None:
            self._host = st
        else:
            self._host = {k: np.concatenate([cur[k], st[k]]) for k in st if k != "alive"}
        self._dirty = True
        return ""

    def _pull_host(self):
        """Host copy of the resident particles (download -> mutate -> upload contract)."""
        if self._host is not None:
            return self._host
        if self._n_uploaded == 0:
            return None
        d = self.engine.download()
        self._host = dict(x=d["x"], v=d["v"], F=d["F"], b=d["b"], mass=d["mass"], vol=d["vol"], ps=d["ps"], group=d["group"])
        return self._host

    def _push(self):
        if self._dirty and self._host is not None:
            h = self._host
            self.engine.upload(h["x"], h["v"], h["mass"], h["vol"], h["F"], h["b"], h["ps"], h["group"])
            self._n_uploaded = len(h["x"])
            self._dirty = False
            self._host = None

    # ---- time stepping
    def substep(self):
        self._push()
        self.engine.substep(1)
        self.current_t += self.base_delta_t
        self.substep_counter += 1

    def step(self, dt):
        """MPM<dim>::step (src/mpm.cpp:428-450): dt<0 runs exactly one substep."""
        self._push()
        if dt < 0:
            self.substep()
            self.request_t = self.current_t
            return
        self.request_t += dt
        n = 0
        t = self.current_t
        while t + self.base_delta_t < self.request_t:
            t += self.base_delta_t
            n += 1
        if n:
            self.engine.substep(n)
            self.current_t = t
            self.substep_counter += n
        self._host = None

    def get_current_time(self):
        return self.current_t

    def get_particles(self):
        """Live particles as numpy arrays (the data visualize() writes: src/visualize.cpp:16-100)."""
        self._push()
        return self.engine.download()

    def num_particles(self):
        self._push()
        return self.engine.num_particles()
