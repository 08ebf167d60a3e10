"""ctypes binding of the C-ABI in include/mpmb.h (libmpmb.so).

This is the only way Python reaches the engine: plain pointers and sizes, no torch types.
The library must exist (build.py builds it in-tree); there is NO CPU fallback — if the CUDA
library cannot be loaded, import of this module's `lib()` raises.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

MPMB_MAX_GROUPS = 16
MPMB_MAT_PARAMS = 8
MPMB_N_STAGES = 5
MAT_LINEAR, MAT_JELLY, MAT_SNOW, MAT_WATER, MAT_SAND, MAT_ELASTIC, MAT_VON_MISES, MAT_VISCO = range(8)
MATERIAL_BY_NAME = {"linear": MAT_LINEAR, "jelly": MAT_JELLY, "snow": MAT_SNOW, "water": MAT_WATER, "sand": MAT_SAND,
                    "elastic": MAT_ELASTIC, "von_mises": MAT_VON_MISES, "visco": MAT_VISCO}

# every symbol include/mpmb.h declares
EXPORTS = [
    "mpmb_create", "mpmb_destroy", "mpmb_last_error", "mpmb_version", "mpmb_set_stream", "mpmb_synchronize",
    "mpmb_set_material", "mpmb_set_delta_t", "mpmb_set_sdf", "mpmb_set_planes", "mpmb_set_levelset_shapes", "mpmb_set_id_base",
    "mpmb_upload_particles", "mpmb_upload_aos", "mpmb_seed_lattice", "mpmb_num_particles", "mpmb_get_update_count", "mpmb_download_bgeo_points", "mpmb_download_particles", "mpmb_download_aos",
    "mpmb_substep", "mpmb_sort_particles_and_populate_grid", "mpmb_rasterize", "mpmb_resample", "mpmb_rasterize_part",
    "mpmb_resample_part", "mpmb_download_grid",
    "mpmb_set_profiling", "mpmb_get_profile", "mpmb_get_counters", "mpmb_get_ordering_stats",
    "mpmb_halo_bytes", "mpmb_halo_pack", "mpmb_halo_unpack", "mpmb_migrate_bytes", "mpmb_migrate_pack", "mpmb_migrate_unpack",
    "mpmb_set_rigid_samples", "mpmb_set_rigid_coupling", "mpmb_set_rigid_state", "mpmb_get_rigid_state", "mpmb_set_particle_states",
    "mpmb_get_particle_cdf", "mpmb_download_cdf",
    "mpmb_xchg_buffer", "mpmb_xchg_ipc_handle", "mpmb_xchg_connect", "mpmb_halo_send", "mpmb_halo_recv", "mpmb_migrate_send", "mpmb_migrate_recv",
]


class MpmbConfig(C.Structure):
    _fields_ = [
        ("res", C.c_int32 * 3),
        ("dx", C.c_float),
        ("dt", C.c_float),
        ("gravity", C.c_float * 3),
        ("particle_gravity", C.c_int32),
        ("clean_boundary", C.c_int32),
        ("device", C.c_int32),
        ("capacity", C.c_int64),
        ("rank", C.c_int32),
        ("world", C.c_int32),
        ("tile_z0", C.c_int32),
        ("tile_z1", C.c_int32),
        ("migrate_capacity", C.c_int64),
        ("halo_capacity", C.c_int32),
        ("no_graph", C.c_int32),
        ("reserved", C.c_int32 * 6),
    ]


class MpmbAosLayout(C.Structure):
    _fields_ = [
        ("stride", C.c_int32), ("off_pos", C.c_int32), ("off_v_and_m", C.c_int32), ("off_dg_e", C.c_int32),
        ("off_apic_b", C.c_int32), ("col_pitch", C.c_int32), ("off_vol", C.c_int32), ("off_scalar", C.c_int32),
        ("reserved", C.c_int32 * 4),
    ]


class MpmbShape(C.Structure):
    _fields_ = [("kind", C.c_int32), ("inside_out", C.c_int32), ("p", C.c_float * 6)]


class MpmbRigidBody(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("rot", C.c_float * 9), ("velocity", C.c_float * 3), ("angular_velocity", C.c_float * 3),
                ("inv_mass", C.c_float), ("inv_inertia", C.c_float * 9), ("frictions", C.c_float * 2)]


SHAPE_PLANE, SHAPE_SPHERE, SHAPE_CUBOID = range(3)
_LIB = None


def lib_path():
    """The product library, or the A/B variant named by MPMB_LIB (a build of the same source with
    experiment defines, `python -m taichi_mpm_b200.build --define ...`; kernel tuning only)."""
    return os.environ.get("MPMB_LIB") or _build.LIB


def lib():
    """Loads libmpmb.so (building it in-tree if the source is newer).  Raises if unavailable."""
    global _LIB
    if _LIB is None:
        path = _build.build()
        if os.environ.get("MPMB_LIB"):
            path = os.environ["MPMB_LIB"]
        if not os.path.exists(path):
            raise RuntimeError("libmpmb.so is missing: the CUDA engine has not been built (python -m taichi_mpm_b200.build)")
        L = C.CDLL(path)
        L.mpmb_last_error.restype = C.c_char_p
        L.mpmb_last_error.argtypes = [C.c_void_p]
        L.mpmb_halo_bytes.restype = C.c_int64
        L.mpmb_migrate_bytes.restype = C.c_int64
        L.mpmb_halo_bytes.argtypes = [C.c_void_p]
        L.mpmb_migrate_bytes.argtypes = [C.c_void_p]
        L.mpmb_create.argtypes = [C.POINTER(MpmbConfig), C.POINTER(C.c_void_p)]
        for name in EXPORTS:
            getattr(L, name)  # AttributeError if a declared symbol is not exported
        _LIB = L
    return _LIB


class MpmbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("mpmb error %d: %s" % (code, msg))
        self.code = code


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, np.float32)
    if shape is not None:
        a = a.reshape(shape)
    return a


class Engine:
    """Thin object wrapper over one MpmbHandle."""

    def __init__(self, res, dx, dt, gravity=(0.0, -10.0, 0.0), particle_gravity=True, clean_boundary=True, device=0,
                 capacity=0, rank=0, world=1, tile_z0=0, tile_z1=0, migrate_capacity=0, halo_capacity=0, no_graph=False):
        self.L = lib()
        cfg = MpmbConfig()
        if np.isscalar(res):
            res = (res, res, res)
        cfg.res[:] = [int(r) for r in res]
        cfg.dx, cfg.dt = float(dx), float(dt)
        cfg.gravity[:] = [float(g) for g in gravity]
        cfg.particle_gravity = int(bool(particle_gravity))
        cfg.clean_boundary = int(bool(clean_boundary))
        cfg.device = int(device)
        cfg.capacity = int(capacity)
        cfg.rank, cfg.world, cfg.tile_z0, cfg.tile_z1 = int(rank), int(world), int(tile_z0), int(tile_z1)
        cfg.migrate_capacity = int(migrate_capacity)
        cfg.halo_capacity = int(halo_capacity)
        cfg.no_graph = int(bool(no_graph))
        self.res = tuple(int(r) for r in res)
        self.cfg = cfg
        self.h = C.c_void_p()
        rc = self.L.mpmb_create(C.byref(cfg), C.byref(self.h))
        if rc != 0:
            raise MpmbError(rc, self.L.mpmb_last_error(None).decode())

    def _check(self, rc):
        if rc != 0:
            raise MpmbError(rc, self.L.mpmb_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.L.mpmb_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- scene
    def set_stream(self, cuda_stream):
        self._check(self.L.mpmb_set_stream(self.h, C.c_void_p(int(cuda_stream))))

    def synchronize(self):
        self._check(self.L.mpmb_synchronize(self.h))

    def set_material(self, group, kind, params):
        p = np.zeros(MPMB_MAT_PARAMS, np.float32)
        p[: len(params)] = params
        self._check(self.L.mpmb_set_material(self.h, C.c_int32(group), C.c_int32(kind), _ptr(p), C.c_int32(len(params))))

    def set_delta_t(self, dt):
        self._check(self.L.mpmb_set_delta_t(self.h, C.c_float(float(dt))))

    def set_sdf(self, sdf4, friction):
        if sdf4 is not None:
            sdf4 = _f32(sdf4)
            assert sdf4.shape == tuple(r + 1 for r in self.res) + (4,)
        self._check(self.L.mpmb_set_sdf(self.h, _ptr(sdf4), C.c_float(friction)))

    def set_planes(self, planes4, friction):
        planes4 = _f32(planes4).reshape(-1, 4)
        self._check(self.L.mpmb_set_planes(self.h, C.c_int32(len(planes4)), _ptr(planes4), C.c_float(friction)))

    def set_levelset_shapes(self, shapes, friction):
        """shapes: [(kind, inside_out, params[<=6])] in GRID units (mpmb.h)."""
        arr = (MpmbShape * len(shapes))()
        for k, (kind, io, prm) in enumerate(shapes):
            arr[k].kind, arr[k].inside_out = int(kind), int(bool(io))
            for j, v in enumerate(prm):
                arr[k].p[j] = float(v)
        self._check(self.L.mpmb_set_levelset_shapes(self.h, C.c_int32(len(shapes)), arr, C.c_float(friction)))

    def set_id_base(self, base):
        self._check(self.L.mpmb_set_id_base(self.h, C.c_int64(int(base))))

    # --- particles
    def upload(self, x, v, mass, vol, F=None, b=None, scalar=None, group=None):
        x = _f32(x).reshape(-1, 3)
        n = len(x)
        v, F, b = _f32(v, (n, 3)), _f32(F, (n, 9)), _f32(b, (n, 9))
        mass, vol, scalar = _f32(mass, (n,)), _f32(vol, (n,)), _f32(scalar, (n,))
        group = None if group is None else np.ascontiguousarray(group, np.int32).reshape(n)
        self._check(self.L.mpmb_upload_particles(self.h, C.c_int64(n), _ptr(x), _ptr(v), _ptr(F), _ptr(b), _ptr(mass), _ptr(vol),
                                                 _ptr(scalar), _ptr(group)))

    def upload_ptrs(self, n, x, v, F, b, mass, vol, scalar, group):
        """Raw-pointer variant (ints), e.g. pinned host buffers owned by the caller."""
        vp = lambda p: C.c_void_p(p) if p else None
        self._check(self.L.mpmb_upload_particles(self.h, C.c_int64(n), vp(x), vp(v), vp(F), vp(b), vp(mass), vp(vol), vp(scalar), vp(group)))

    def download_ptrs(self, cap, id_, x, v, F, b, mass, vol, scalar, group):
        vp = lambda p: C.c_void_p(p) if p else None
        n = C.c_int64(0)
        self._check(self.L.mpmb_download_particles(self.h, C.c_int64(cap), C.byref(n), vp(id_), vp(x), vp(v), vp(F), vp(b), vp(mass),
                                                   vp(vol), vp(scalar), vp(group)))
        return n.value

    def seed_lattice(self, lo_cell, hi_cell, vol, mass, jitter=0.0, seed=0, group=0, v0=(0.0, 0.0, 0.0)):
        """Device-side `benchmark` lattice (src/mpm.cpp:149-186); returns the number of particles this engine created."""
        lo = (C.c_int32 * 3)(*[int(v) for v in lo_cell])
        hi = (C.c_int32 * 3)(*[int(v) for v in hi_cell])
        vv = (C.c_float * 3)(*[float(v) for v in v0])
        n = C.c_int64(0)
        self._check(self.L.mpmb_seed_lattice(self.h, lo, hi, C.c_float(vol), C.c_float(mass), C.c_float(jitter), C.c_uint32(int(seed)),
                                             C.c_int32(group), vv, C.byref(n)))
        return n.value

    def upload_aos(self, pool, indices, layout, group=None):
        pool = np.ascontiguousarray(pool, np.uint8)
        indices = np.ascontiguousarray(indices, np.uint32)
        slots = pool.size // layout.stride
        group = None if group is None else np.ascontiguousarray(group, np.int32)
        self._check(self.L.mpmb_upload_aos(self.h, C.c_int64(len(indices)), _ptr(pool), C.c_int64(slots), _ptr(indices), C.byref(layout),
                                           _ptr(group)))

    def download_aos(self, pool, indices, layout):
        assert pool.dtype == np.uint8 and pool.flags.c_contiguous and indices.dtype == np.uint32
        n = C.c_int64(0)
        self._check(self.L.mpmb_download_aos(self.h, _ptr(pool), C.c_int64(pool.size // layout.stride), _ptr(indices), C.c_int64(len(indices)),
                                             C.byref(layout), C.byref(n)))
        return n.value

    def num_particles(self):
        n = C.c_int64(0)
        self._check(self.L.mpmb_num_particles(self.h, C.byref(n)))
        return n.value

    def download_bgeo_points(self, id_range, cap=None):
        """Device-packed BGEO point block (48 big-endian bytes per particle, id order): (n, bytes ndarray)."""
        if cap is None:
            cap = max(self.num_particles(), 1)
        buf = np.empty(cap * 48, np.uint8)
        n = C.c_int64(0)
        self._check(self.L.mpmb_download_bgeo_points(self.h, C.c_int64(int(id_range)), _ptr(buf), C.c_int64(cap), C.byref(n)))
        return n.value, buf[: n.value * 48]

    def update_count(self):
        """Particle updates so far, as the reference's update_counter counts them (src/mpm.cpp:436)."""
        n = C.c_int64(0)
        self._check(self.L.mpmb_get_update_count(self.h, C.byref(n)))
        return n.value

    def download(self, cap=None, sort_by_id=True):
        """Returns dict(id,x,v,F,b,mass,vol,ps,group) of the live particles."""
        if cap is None:
            cap = max(self.num_particles(), 1)
        out = dict(id=np.zeros(cap, np.uint32), x=np.zeros((cap, 3), np.float32), v=np.zeros((cap, 3), np.float32),
                   F=np.zeros((cap, 9), np.float32), b=np.zeros((cap, 9), np.float32), mass=np.zeros(cap, np.float32),
                   vol=np.zeros(cap, np.float32), ps=np.zeros(cap, np.float32), group=np.zeros(cap, np.int32))
        n = C.c_int64(0)
        self._check(self.L.mpmb_download_particles(self.h, C.c_int64(cap), C.byref(n), _ptr(out["id"]), _ptr(out["x"]), _ptr(out["v"]),
                                                   _ptr(out["F"]), _ptr(out["b"]), _ptr(out["mass"]), _ptr(out["vol"]), _ptr(out["ps"]),
                                                   _ptr(out["group"])))
        out = {k: a[: n.value] for k, a in out.items()}
        if sort_by_id:
            o = np.argsort(out["id"], kind="stable")
            out = {k: a[o] for k, a in out.items()}
        return out

    # --- hot path
    def substep(self, nsub=1):
        self._check(self.L.mpmb_substep(self.h, C.c_int32(nsub)))

    # ---- rigid bodies (CPIC): `rigid` is the dict scenes.make_rigid builds (row 0 = the background body)
    def _rigid_records(self, rigid):
        nb = len(rigid["inv_mass"])
        arr = (MpmbRigidBody * nb)()
        for b in range(nb):
            arr[b].position[:] = [float(v) for v in rigid["position"][b]]
            arr[b].rot[:] = [float(v) for v in np.asarray(rigid["rot"][b]).reshape(9)]
            arr[b].velocity[:] = [float(v) for v in rigid["velocity"][b]]
            arr[b].angular_velocity[:] = [float(v) for v in rigid["angular_velocity"][b]]
            arr[b].inv_mass = float(rigid["inv_mass"][b])
            arr[b].inv_inertia[:] = [float(v) for v in np.asarray(rigid["inv_inertia"][b]).reshape(9)]
            arr[b].frictions[:] = [float(v) for v in rigid["frictions"][b]]
        return nb, arr

    def set_rigid(self, rigid):
        """Boundary samples + coupling constants + the current state of every body."""
        off, tri = _f32(rigid["sample_offset"], (-1, 3)), _f32(rigid["sample_tri"], (-1, 9))
        rid = np.ascontiguousarray(rigid["sample_rigid"], np.int32)
        self._check(self.L.mpmb_set_rigid_samples(self.h, C.c_int32(len(rigid["inv_mass"])), C.c_int64(len(rid)), _ptr(off), _ptr(tri), _ptr(rid)))
        if len(rid):
            self._check(self.L.mpmb_set_rigid_coupling(self.h, C.c_float(rigid.get("penalty", 0.0)), C.c_float(rigid.get("pushing_force", 20000.0))))
            self.set_rigid_state(rigid)

    def set_rigid_state(self, rigid):
        nb, arr = self._rigid_records(rigid)
        self._check(self.L.mpmb_set_rigid_state(self.h, C.c_int32(nb), arr))

    def get_rigid_state(self, n_bodies):
        arr = (MpmbRigidBody * n_bodies)()
        self._check(self.L.mpmb_get_rigid_state(self.h, C.c_int32(n_bodies), arr))
        return dict(velocity=np.array([list(a.velocity) for a in arr], np.float32), angular_velocity=np.array([list(a.angular_velocity) for a in arr], np.float32),
                    position=np.array([list(a.position) for a in arr], np.float32))

    def set_particle_states(self, states):
        st = np.ascontiguousarray(states, np.uint32)
        self._check(self.L.mpmb_set_particle_states(self.h, C.c_int64(len(st)), _ptr(st)))

    def get_particle_cdf(self, n):
        out = dict(states=np.zeros(n, np.uint32), bnormal=np.zeros((n, 3), np.float32), bdist=np.zeros(n, np.float32), near=np.zeros(n, np.uint8))
        self._check(self.L.mpmb_get_particle_cdf(self.h, C.c_int64(n), _ptr(out["states"]), _ptr(out["bnormal"]), _ptr(out["bdist"]), _ptr(out["near"])))
        return out

    def download_cdf(self):
        nn = tuple(r + 1 for r in self.res)
        st, d = np.zeros(nn, np.uint32), np.zeros(nn, np.float32)
        self._check(self.L.mpmb_download_cdf(self.h, _ptr(st), _ptr(d)))
        return dict(node_state=st, node_dist=d)

    def sort_particles_and_populate_grid(self):
        self._check(self.L.mpmb_sort_particles_and_populate_grid(self.h))

    def rasterize(self):
        self._check(self.L.mpmb_rasterize(self.h))

    def resample(self):
        self._check(self.L.mpmb_resample(self.h))

    def rasterize_part(self, part):
        self._check(self.L.mpmb_rasterize_part(self.h, C.c_int32(part)))

    def resample_part(self, part):
        self._check(self.L.mpmb_resample_part(self.h, C.c_int32(part)))

    def download_grid(self, which):
        g = np.zeros(tuple(r + 1 for r in self.res) + (4,), np.float32)
        self._check(self.L.mpmb_download_grid(self.h, C.c_int32(which), _ptr(g)))
        return g

    # --- z-slab exchange (device pointers owned by the caller)
    def halo_bytes(self):
        return int(self.L.mpmb_halo_bytes(self.h))

    def migrate_bytes(self):
        return int(self.L.mpmb_migrate_bytes(self.h))

    def halo_pack(self, face, dev_ptr):
        self._check(self.L.mpmb_halo_pack(self.h, C.c_int32(face), C.c_void_p(int(dev_ptr))))

    def halo_unpack(self, face, dev_ptr):
        self._check(self.L.mpmb_halo_unpack(self.h, C.c_int32(face), C.c_void_p(int(dev_ptr))))

    def migrate_pack(self, face, dev_ptr):
        self._check(self.L.mpmb_migrate_pack(self.h, C.c_int32(face), C.c_void_p(int(dev_ptr))))

    def migrate_unpack(self, face, dev_ptr):
        self._check(self.L.mpmb_migrate_unpack(self.h, C.c_int32(face), C.c_void_p(int(dev_ptr))))

    # --- peer-memory exchange
    def xchg_buffer(self, kind, face):
        p = C.c_void_p()
        self._check(self.L.mpmb_xchg_buffer(self.h, C.c_int32(kind), C.c_int32(face), C.byref(p)))
        return p.value

    def xchg_ipc_handle(self, kind, face):
        buf = (C.c_ubyte * 64)()
        self._check(self.L.mpmb_xchg_ipc_handle(self.h, C.c_int32(kind), C.c_int32(face), buf))
        return bytes(buf)

    def xchg_connect(self, kind, face, handle=None, ptr=None):
        hb = (C.c_ubyte * 64).from_buffer_copy(handle) if handle is not None else None
        self._check(self.L.mpmb_xchg_connect(self.h, C.c_int32(kind), C.c_int32(face), hb, C.c_void_p(ptr) if ptr else None))

    def halo_send(self, face):
        self._check(self.L.mpmb_halo_send(self.h, C.c_int32(face)))

    def halo_recv(self, face):
        self._check(self.L.mpmb_halo_recv(self.h, C.c_int32(face)))

    def migrate_send(self, face):
        self._check(self.L.mpmb_migrate_send(self.h, C.c_int32(face)))

    def migrate_recv(self, face):
        self._check(self.L.mpmb_migrate_recv(self.h, C.c_int32(face)))

    # --- profiling
    def set_profiling(self, enabled):
        self._check(self.L.mpmb_set_profiling(self.h, C.c_int32(int(enabled))))

    def get_profile(self, reset=True):
        ms = (C.c_double * MPMB_N_STAGES)()
        ln = (C.c_int64 * MPMB_N_STAGES)()
        self._check(self.L.mpmb_get_profile(self.h, ms, ln, C.c_int32(int(reset))))
        return list(ms), list(ln)

    def get_ordering_stats(self):
        a, b, c = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        self._check(self.L.mpmb_get_ordering_stats(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(rows=a.value, movers=b.value, ghost_tiles=c.value)

    def get_counters(self):
        a, b, c = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        self._check(self.L.mpmb_get_counters(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(active_tiles=a.value, alive=b.value, kernel_launches=c.value)
