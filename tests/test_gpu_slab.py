"""z-slab path on real GPUs (needs >= 2 devices; `gpurun --gpus 2`): two ranks with halo exchange and
particle migration must reproduce the single-GPU run of the same scene."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
SIMT = os.environ.get("MPMB_SIMT", "") == "1"   # the SIMT emulator build: device memory is host memory, no torch CUDA


class _HostBuf:
    """numpy-backed stand-in for a torch CUDA byte tensor (emulator runs)"""

    def __init__(self, n):
        self.a = np.zeros(n, np.uint8)

    def data_ptr(self):
        return self.a.ctypes.data


def _scene():
    from taichi_mpm_b200 import scenes
    res = (32, 32, 64)
    dx = 1.0 / 32
    x, mass, vol = scenes.lattice_block(32, (11, 9, 14), (21, 17, 50), jitter=0.2, seed=3)
    st = scenes.make_state(x, mass, vol, scenes.MAT_SAND)
    rng = np.random.default_rng(5)
    n = len(x)
    # motion along z so that particles cross the slab boundary in both directions
    vz = np.where(x[:, 0] < 0.5, 2.5, -2.5)
    st["v"] = np.stack([0.3 * rng.normal(size=n), 0.3 * rng.normal(size=n), vz + 0.2 * rng.normal(size=n)], 1).astype(np.float32)
    scene = dict(res=res, dx=dx, dt=1e-4, gravity=(0.0, -10.0, 0.0), particle_gravity=1, mat_kind=np.array([scenes.MAT_SAND], np.int32),
                 mat_params=scenes.material_params(scenes.MAT_SAND)[None], planes=np.array([[0.0, 1.0, 0.0, -9.6]], np.float32), friction=0.4)
    return scene, st


def _worker(rank, world, port, nsub, out_path):
    import torch
    import torch.distributed as dist
    from taichi_mpm_b200 import capi, slab
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    scene, st = _scene()
    tz = slab.base_tile_z(st["x"][:, 2], scene["dx"])
    cuts = slab.slab_partition(tz, slab.tile_layers(scene["res"][2]), world)
    z0, z1 = cuts[rank]
    mine = np.nonzero((tz >= z0) & (tz < z1))[0]
    e = capi.Engine(scene["res"], scene["dx"], scene["dt"], scene["gravity"], 1, True, device=rank, rank=rank, world=world, tile_z0=z0, tile_z1=z1,
                    migrate_capacity=4096, halo_capacity=64)
    e.set_stream(torch.cuda.current_stream().cuda_stream)
    e.set_material(0, int(scene["mat_kind"][0]), scene["mat_params"][0])
    e.set_planes(scene["planes"], scene["friction"])
    # global ids: upload in global order restricted to this rank, id_base such that ranges are disjoint
    counts = [int(((tz >= a) & (tz < b)).sum()) for a, b in cuts]
    e.set_id_base(sum(counts[:rank]))
    sub = {k: v[mine] for k, v in st.items()}
    e.upload(sub["x"], sub["v"], sub["mass"], sub["vol"], sub["F"], sub["b"], sub["ps"], sub["group"])
    r = slab.SlabRunner(slab.EngineAdapter(e), rank, world, torch.device("cuda", rank), dist=dist)
    r.substep(nsub)
    torch.cuda.synchronize()
    got = e.download()
    # map slab ids back to indices of the global scene
    order = np.concatenate([np.nonzero((tz >= a) & (tz < b))[0] for a, b in cuts])
    got["gid"] = order[got["id"].astype(np.int64)]
    np.savez(out_path % rank, **got, n_start=len(mine))
    dist.barrier()
    dist.destroy_process_group()
    e.close()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.needs_cuda
def test_two_slabs_match_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from tests import common as T
    nsub = 60
    out_path = str(tmp_path / "rank%d.npz")
    mp.spawn(_worker, args=(2, _free_port(), nsub, out_path), nprocs=2, join=True)
    scene, st = _scene()
    e = T.make_engine(scene, st)
    e.substep(nsub)
    ref = e.download()
    e.close()
    parts = [np.load(out_path % r) for r in range(2)]
    gid = np.concatenate([p["gid"] for p in parts])
    assert len(gid) == len(ref["id"]) and len(np.unique(gid)) == len(gid)
    o = np.argsort(gid)
    assert np.array_equal(gid[o], ref["id"].astype(np.int64))
    # particles really moved between ranks
    moved = sum(abs(len(p["gid"]) - int(p["n_start"])) for p in parts)
    in0 = set(parts[0]["gid"].tolist())
    tz = None
    for k, tol in (("x", 2e-6), ("v", 5e-4), ("F", 5e-5), ("ps", 1e-5)):
        got = np.concatenate([p[k] for p in parts])[o]
        scale = max(np.abs(ref[k]).max(), 1e-30) if k == "v" else 1.0
        assert np.abs(got - ref[k]).max() <= tol * scale, k
    from taichi_mpm_b200 import slab
    tz0 = slab.base_tile_z(st["x"][:, 2], scene["dx"])
    cuts = slab.slab_partition(tz0, slab.tile_layers(scene["res"][2]), 2)
    started0 = set(np.nonzero(tz0 < cuts[0][1])[0].tolist())
    assert len(in0 - started0) > 0 and len(started0 - in0) > 0, "no migration happened in either direction"


def _run_two_slabs_one_process(scene, st, nsub, device=0):
    """Both slabs of a 2-rank run in ONE process on one GPU (no NCCL): the exchange buffers are
    copied engine to engine.  Exercises halo ghost tiles and migration with the default 1-GPU suite."""
    if not SIMT:
        import torch
    from taichi_mpm_b200 import capi, slab
    world = 2
    tz = slab.base_tile_z(st["x"][:, 2], scene["dx"])
    cuts = slab.slab_partition(tz, slab.tile_layers(scene["res"][2]), world)
    counts = [int(((tz >= a) & (tz < b)).sum()) for a, b in cuts]
    engines, adapters, bufs = [], [], []
    dev = None if SIMT else torch.device("cuda", device)
    for rank in range(world):
        z0, z1 = cuts[rank]
        e = capi.Engine(scene["res"], scene["dx"], scene["dt"], scene["gravity"], 1, True, device=device, rank=rank, world=world,
                        tile_z0=z0, tile_z1=z1, migrate_capacity=4096, halo_capacity=64)
        if not SIMT:
            e.set_stream(torch.cuda.current_stream().cuda_stream)
        e.set_material(0, int(scene["mat_kind"][0]), scene["mat_params"][0])
        e.set_planes(scene["planes"], scene["friction"])
        e.set_id_base(sum(counts[:rank]))
        mine = np.nonzero((tz >= z0) & (tz < z1))[0]
        e.upload(*(st[k][mine] for k in ("x", "v", "mass", "vol", "F", "b", "ps", "group")))
        engines.append(e)
        adapters.append(slab.EngineAdapter(e))
        mk = (lambda n: _HostBuf(max(int(n), 16))) if SIMT else (lambda n: torch.zeros(max(int(n), 16), dtype=torch.uint8, device=dev))
        bufs.append(dict(halo=[mk(e.halo_bytes()), mk(e.halo_bytes())], mig=[mk(e.migrate_bytes()), mk(e.migrate_bytes())]))
    a0, a1 = adapters
    for _ in range(nsub):
        for a in adapters:
            a.sort(); a.rasterize_part(1)
        a0.halo_pack(1, bufs[0]["halo"][1]); a1.halo_pack(0, bufs[1]["halo"][0])
        for a in adapters:
            a.rasterize_part(2); a.resample_part(2)
        a1.halo_unpack(0, bufs[0]["halo"][1]); a0.halo_unpack(1, bufs[1]["halo"][0])
        for a in adapters:
            a.resample_part(1)
        a0.migrate_pack(1, bufs[0]["mig"][1]); a1.migrate_pack(0, bufs[1]["mig"][0])
        a1.migrate_unpack(0, bufs[0]["mig"][1]); a0.migrate_unpack(1, bufs[1]["mig"][0])
    if not SIMT:
        torch.cuda.synchronize()
    order = np.concatenate([np.nonzero((tz >= a) & (tz < b))[0] for a, b in cuts])
    parts = []
    for e in engines:
        got = e.download()
        got["gid"] = order[got["id"].astype(np.int64)]
        parts.append(got)
        e.close()
    return parts, cuts, tz


def test_two_slabs_on_one_gpu_match_single_run():
    from tests import common as T
    scene, st = _scene()
    nsub = 60
    parts, cuts, tz0 = _run_two_slabs_one_process(scene, st, nsub)
    e = T.make_engine(scene, st)
    e.substep(nsub)
    ref = e.download()
    e.close()
    gid = np.concatenate([p["gid"] for p in parts])
    assert len(gid) == len(ref["id"]) and len(np.unique(gid)) == len(gid)
    o = np.argsort(gid)
    assert np.array_equal(gid[o], ref["id"].astype(np.int64))
    for k, tol in (("x", 2e-6), ("v", 5e-4), ("F", 5e-5), ("ps", 1e-5)):
        got = np.concatenate([p[k] for p in parts])[o]
        scale = max(np.abs(ref[k]).max(), 1e-30) if k == "v" else 1.0
        assert np.abs(got - ref[k]).max() <= tol * scale, k
    in0 = set(parts[0]["gid"].tolist())
    started0 = set(np.nonzero(tz0 < cuts[0][1])[0].tolist())
    assert len(in0 - started0) > 0 and len(started0 - in0) > 0, "no migration happened in either direction"


def test_two_slabs_peer_memory_exchange_one_gpu():
    """Same two slabs, but the exchange runs through the engine's peer-memory path (pack kernels write
    into the other engine's receive buffer, seq flag, wait kernel) instead of host-moved buffers."""
    from taichi_mpm_b200 import capi, slab
    from tests import common as T
    scene, st = _scene()
    nsub, world = 40, 2
    tz = slab.base_tile_z(st["x"][:, 2], scene["dx"])
    cuts = slab.slab_partition(tz, slab.tile_layers(scene["res"][2]), world)
    counts = [int(((tz >= a) & (tz < b)).sum()) for a, b in cuts]
    eng = []
    for rank in range(world):
        z0, z1 = cuts[rank]
        e = capi.Engine(scene["res"], scene["dx"], scene["dt"], scene["gravity"], 1, True, device=0, rank=rank, world=world,
                        tile_z0=z0, tile_z1=z1, migrate_capacity=4096, halo_capacity=64)
        e.set_material(0, int(scene["mat_kind"][0]), scene["mat_params"][0])
        e.set_planes(scene["planes"], scene["friction"])
        e.set_id_base(sum(counts[:rank]))
        mine = np.nonzero((tz >= z0) & (tz < z1))[0]
        e.upload(*(st[k][mine] for k in ("x", "v", "mass", "vol", "F", "b", "ps", "group")))
        eng.append(e)
    e0, e1 = eng
    for k in (0, 1):
        e0.xchg_connect(k, 1, ptr=e1.xchg_buffer(k, 0))
        e1.xchg_connect(k, 0, ptr=e0.xchg_buffer(k, 1))
    for _ in range(nsub):   # one stream, one process: interleave the stages so that every wait finds its flag
        for e in eng:
            e.sort_particles_and_populate_grid(); e.rasterize()
        e0.halo_send(1); e1.halo_send(0)
        e0.halo_recv(1); e1.halo_recv(0)
        for e in eng:
            e.resample()
        e0.migrate_send(1); e1.migrate_send(0)
        e0.migrate_recv(1); e1.migrate_recv(0)
    order = np.concatenate([np.nonzero((tz >= a) & (tz < b))[0] for a, b in cuts])
    parts = []
    for e in eng:
        g = e.download()
        g["gid"] = order[g["id"].astype(np.int64)]
        parts.append(g)
        e.close()
    ref_e = T.make_engine(scene, st)
    ref_e.substep(nsub)
    ref = ref_e.download()
    ref_e.close()
    gid = np.concatenate([p["gid"] for p in parts])
    o = np.argsort(gid)
    assert np.array_equal(gid[o], ref["id"].astype(np.int64))
    for k, tol in (("x", 2e-6), ("v", 5e-4), ("F", 5e-5)):
        got = np.concatenate([p[k] for p in parts])[o]
        scale = max(np.abs(ref[k]).max(), 1e-30) if k == "v" else 1.0
        assert np.abs(got - ref[k]).max() <= tol * scale, k


def _long_scene():
    """The same kind of block, four times longer in z (25 tile layers): room for eight slabs."""
    from taichi_mpm_b200 import scenes
    res = (32, 32, 128)
    dx = 1.0 / 32
    x, mass, vol = scenes.lattice_block(32, (12, 9, 14), (20, 15, 114), jitter=0.2, seed=7)
    st = scenes.make_state(x, mass, vol, scenes.MAT_SAND)
    rng = np.random.default_rng(8)
    n = len(x)
    vz = np.where(x[:, 0] < 0.5, 2.5, -2.5)
    st["v"] = np.stack([0.3 * rng.normal(size=n), 0.3 * rng.normal(size=n), vz + 0.2 * rng.normal(size=n)], 1).astype(np.float32)
    scene = dict(res=res, dx=dx, dt=1e-4, gravity=(0.0, -10.0, 0.0), particle_gravity=1, mat_kind=np.array([scenes.MAT_SAND], np.int32),
                 mat_params=scenes.material_params(scenes.MAT_SAND)[None], planes=np.array([[0.0, 1.0, 0.0, -9.6]], np.float32), friction=0.4)
    return scene, st


def _run_slabs_peer_one_process(scene, st, world, nsub):
    from taichi_mpm_b200 import capi, slab
    from tests import common as T
    tz = slab.base_tile_z(st["x"][:, 2], scene["dx"])
    cuts = slab.slab_partition(tz, slab.tile_layers(scene["res"][2]), world)
    counts = [int(((tz >= a) & (tz < b)).sum()) for a, b in cuts]
    assert min(counts) > 0
    eng = []
    for rank in range(world):
        z0, z1 = cuts[rank]
        e = capi.Engine(scene["res"], scene["dx"], scene["dt"], scene["gravity"], 1, True, device=0, rank=rank, world=world,
                        tile_z0=z0, tile_z1=z1, migrate_capacity=4096, halo_capacity=64)
        e.set_material(0, int(scene["mat_kind"][0]), scene["mat_params"][0])
        e.set_planes(scene["planes"], scene["friction"])
        e.set_id_base(sum(counts[:rank]))
        mine = np.nonzero((tz >= z0) & (tz < z1))[0]
        e.upload(*(st[k][mine] for k in ("x", "v", "mass", "vol", "F", "b", "ps", "group")))
        eng.append(e)
    for r in range(world - 1):                      # my face 1 -> upper neighbour's face-0 buffer, and back
        for k in (0, 1):
            eng[r].xchg_connect(k, 1, ptr=eng[r + 1].xchg_buffer(k, 0))
            eng[r + 1].xchg_connect(k, 0, ptr=eng[r].xchg_buffer(k, 1))
    faces = [[f for f in (0, 1) if 0 <= r + (1 if f else -1) < world] for r in range(world)]
    for _ in range(nsub):                           # one stream, one process: every wait finds its flag already published
        for e in eng:
            e.sort_particles_and_populate_grid(); e.rasterize()
        for r, e in enumerate(eng):
            for f in faces[r]:
                e.halo_send(f)
        for r, e in enumerate(eng):
            for f in faces[r]:
                e.halo_recv(f)
        for e in eng:
            e.resample()
        for r, e in enumerate(eng):
            for f in faces[r]:
                e.migrate_send(f)
        for r, e in enumerate(eng):
            for f in faces[r]:
                e.migrate_recv(f)
    order = np.concatenate([np.nonzero((tz >= a) & (tz < b))[0] for a, b in cuts])
    parts = []
    for e in eng:
        g = e.download()
        g["gid"] = order[g["id"].astype(np.int64)]
        parts.append(g)
        e.close()
    ref_e = T.make_engine(scene, st)
    ref_e.substep(nsub)
    ref = ref_e.download()
    ref_e.close()
    gid = np.concatenate([p["gid"] for p in parts])
    o = np.argsort(gid)
    assert np.array_equal(gid[o], ref["id"].astype(np.int64))
    for k, tol in (("x", 2e-6), ("v", 5e-4), ("F", 5e-5)):
        got = np.concatenate([p[k] for p in parts])[o]
        scale = max(np.abs(ref[k]).max(), 1e-30) if k == "v" else 1.0
        assert np.abs(got - ref[k]).max() <= tol * scale, k
    # every interior rank really traded particles with both neighbours
    start = [set(np.nonzero((tz >= a) & (tz < b))[0].tolist()) for a, b in cuts]
    for r in range(1, world - 1):
        came = set(parts[r]["gid"].tolist()) - start[r]
        assert came & start[r - 1] and came & start[r + 1], "rank %d did not receive from both sides" % r


def test_four_slabs_peer_memory_exchange_one_process():
    """Four slabs in one process on one device: the two middle ranks exchange halos and migrating particles with BOTH
    neighbours every substep (the topology of the 4- and 8-GPU runs), through the peer-memory path."""
    scene, st = _scene()
    _run_slabs_peer_one_process(scene, st, world=4, nsub=25)


def test_eight_slabs_peer_memory_exchange_one_process():
    """Eight slabs (the 8-GPU topology: id bases, cuts, six interior ranks) on a four times longer block."""
    scene, st = _long_scene()
    _run_slabs_peer_one_process(scene, st, world=8, nsub=15)


def _peer_worker(rank, world, port, nsub, out_path):
    import torch
    import torch.distributed as dist
    from taichi_mpm_b200 import capi, slab
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    scene, st = _scene()
    tz = slab.base_tile_z(st["x"][:, 2], scene["dx"])
    cuts = slab.slab_partition(tz, slab.tile_layers(scene["res"][2]), world)
    z0, z1 = cuts[rank]
    mine = np.nonzero((tz >= z0) & (tz < z1))[0]
    e = capi.Engine(scene["res"], scene["dx"], scene["dt"], scene["gravity"], 1, True, device=rank, rank=rank, world=world, tile_z0=z0, tile_z1=z1,
                    migrate_capacity=4096, halo_capacity=64)
    e.set_material(0, int(scene["mat_kind"][0]), scene["mat_params"][0])
    e.set_planes(scene["planes"], scene["friction"])
    counts = [int(((tz >= a) & (tz < b)).sum()) for a, b in cuts]
    e.set_id_base(sum(counts[:rank]))
    e.upload(*(st[k][mine] for k in ("x", "v", "mass", "vol", "F", "b", "ps", "group")))
    slab.connect_peers(e, rank, world, dist)
    e.substep(nsub)            # whole z-slab substeps inside the C-ABI, exchanges over NVLink peer memory
    e.synchronize()
    got = e.download()
    order = np.concatenate([np.nonzero((tz >= a) & (tz < b))[0] for a, b in cuts])
    got["gid"] = order[got["id"].astype(np.int64)]
    np.savez(out_path % rank, **got)
    dist.barrier()
    e.close()
    dist.destroy_process_group()


@pytest.mark.needs_cuda
def test_two_ranks_peer_memory_exchange(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from tests import common as T
    nsub = 60
    out_path = str(tmp_path / "peer%d.npz")
    mp.spawn(_peer_worker, args=(2, _free_port(), nsub, out_path), nprocs=2, join=True)
    scene, st = _scene()
    e = T.make_engine(scene, st)
    e.substep(nsub)
    ref = e.download()
    e.close()
    parts = [np.load(out_path % r) for r in range(2)]
    gid = np.concatenate([p["gid"] for p in parts])
    o = np.argsort(gid)
    assert np.array_equal(gid[o], ref["id"].astype(np.int64))
    for k, tol in (("x", 2e-6), ("v", 5e-4), ("F", 5e-5), ("ps", 1e-5)):
        got = np.concatenate([p[k] for p in parts])[o]
        scale = max(np.abs(ref[k]).max(), 1e-30) if k == "v" else 1.0
        assert np.abs(got - ref[k]).max() <= tol * scale, k
