"""Host-side logic of the z-slab multi-GPU path on CPU: partitioning, and the exchange
choreography with world_size-2/3/8 gloo processes and a mock engine (no CUDA here)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from taichi_mpm_b200 import slab


def test_tile_layers_and_base_tile():
    assert slab.tile_layers(256) == 66
    dx = 1.0 / 256
    z = np.array([7.49 * dx, 7.51 * dx, 11.5 * dx, 100.3 * dx], np.float32)
    assert list(slab.base_tile_z(z, dx)) == [1, 1, 2, 24]   # base = int(X-0.5) -> 6,7,11,99


def test_partition_balances_particles_over_occupied_extent():
    rng = np.random.default_rng(0)
    tz = rng.integers(20, 45, 100000)
    for world in (1, 2, 3, 4, 8):
        cuts = slab.slab_partition(tz, 66, world)
        assert cuts[0][0] == 0 and cuts[-1][1] == 66 and len(cuts) == world
        assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:])) and all(z1 > z0 for z0, z1 in cuts)
        counts = [((tz >= z0) & (tz < z1)).sum() for z0, z1 in cuts]
        assert sum(counts) == len(tz)
        if world > 1:
            assert max(counts) <= 1.35 * len(tz) / world


def test_partition_degenerate_inputs():
    assert slab.slab_partition(np.array([], np.int64), 10, 3) == [(0, 3), (3, 6), (6, 10)]
    cuts = slab.slab_partition(np.full(1000, 5), 10, 4)          # everything in one layer
    assert len(cuts) == 4 and all(z1 > z0 for z0, z1 in cuts) and cuts[-1][1] == 10


class MockAdapter:
    """Stand-in engine: 'arenas' and 'migrants' are just tagged byte buffers so the test can check who
    received what, in which phase order."""

    def __init__(self, rank):
        self.rank = rank
        self.log = []
        self.received = []

    def halo_bytes(self):
        return 64

    def migrate_bytes(self):
        return 32

    def sort(self):
        self.log.append("sort")

    def rasterize(self):
        self.log.append("p2g")

    def resample(self):
        self.log.append("g2p")

    def rasterize_part(self, part):
        self.log.append("p2g%d" % part)

    def resample_part(self, part):
        self.log.append("g2p%d" % part)

    def halo_pack(self, face, buf):
        buf.fill_(10 * self.rank + face)
        self.log.append("hp%d" % face)

    def halo_unpack(self, face, buf):
        self.received.append(("halo", face, int(buf[0]), bool((buf == buf[0]).all())))
        self.log.append("hu%d" % face)

    def migrate_pack(self, face, buf):
        buf.fill_(100 + 10 * self.rank + face)
        self.log.append("mp%d" % face)

    def migrate_unpack(self, face, buf):
        self.received.append(("mig", face, int(buf[0]), bool((buf == buf[0]).all())))
        self.log.append("mu%d" % face)


def _worker(rank, world, port, q, overlap=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a = MockAdapter(rank)
    r = slab.SlabRunner(a, rank, world, torch.device("cpu"), dist=dist, overlap=overlap)
    r.substep(2)
    q.put((rank, a.log, a.received))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,overlap", [(2, False), (3, False), (2, True), (8, False)])
def test_exchange_choreography_gloo(world, overlap):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, overlap)) for r in range(world)]
    for p in procs:
        p.start()
    out = {}
    for _ in range(world):
        rank, log, recv = q.get(timeout=120)
        out[rank] = (log, recv)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank in range(world):
        log, recv = out[rank]
        lo, hi = rank > 0, rank < world - 1
        plain = ["sort", "p2g"] + (["hp0"] if lo else []) + (["hp1"] if hi else []) + (["hu0"] if lo else []) + (["hu1"] if hi else []) + \
                ["g2p"] + (["mp0"] if lo else []) + (["mp1"] if hi else []) + (["mu0"] if lo else []) + (["mu1"] if hi else [])
        # overlapped schedule: boundary P2G, pack, [exchange in flight] interior P2G + G2P, unpack, boundary G2P
        step = ["sort", "p2g1"] + (["hp0"] if lo else []) + (["hp1"] if hi else []) + ["p2g2", "g2p2"] + (["hu0"] if lo else []) + \
               (["hu1"] if hi else []) + ["g2p1"] + (["mp0"] if lo else []) + (["mp1"] if hi else []) + (["mu0"] if lo else []) + (["mu1"] if hi else [])
        assert log == (step if overlap else plain) * 2
        # face 0 receives what rank-1 packed for ITS face 1 (its top layer); face 1 what rank+1 packed for its face 0
        exp = []
        for _ in range(2):
            if lo:
                exp.append(("halo", 0, 10 * (rank - 1) + 1, True))
            if hi:
                exp.append(("halo", 1, 10 * (rank + 1) + 0, True))
            if lo:
                exp.append(("mig", 0, 100 + 10 * (rank - 1) + 1, True))
            if hi:
                exp.append(("mig", 1, 100 + 10 * (rank + 1) + 0, True))
        assert recv == exp
