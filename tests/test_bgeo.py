"""Frame output: the product's BGEO writer against the reference's own writer.

Golden files under tests/golden/ were written by the Partio copy vendored in the reference tree
(tests/golden/make_bgeo_golden.py); where that tree is present (the build container) the writer is
also compared live, including the 65536-point switch of the primitive's index width."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import pyoracle as O
from taichi_mpm_b200 import bgeo, mpm

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_bgeo_golden", os.path.join(HERE, "golden", "make_bgeo_golden.py"))
G = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(G)


@pytest.mark.parametrize("name,n,seed,verbose", [("frame_plain_17", 17, 1, False), ("frame_verbose_9", 9, 2, True), ("frame_empty", 0, 3, False)])
def test_writer_reproduces_reference_files(tmp_path, name, n, seed, verbose):
    p = G.golden_particles(n, seed)
    out = tmp_path / (name + ".bgeo")
    nbytes = bgeo.write_bgeo(str(out), p["x"], mpm.frame_attributes(p, G.GROUP_KINDS, verbose))
    golden = open(os.path.join(HERE, "golden", name + ".bgeo"), "rb").read()
    assert nbytes == len(golden)
    assert out.read_bytes() == golden                     # bit-exact: byte work


def test_reader_round_trip_of_reference_file():
    p = G.golden_particles(9, 2)
    pos, attrs = bgeo.read_bgeo(os.path.join(HERE, "golden", "frame_verbose_9.bgeo"))
    assert np.array_equal(pos, p["x"])
    want = mpm.frame_attributes(p, G.GROUP_KINDS, True)
    assert [a[0] for a in attrs] == [a[0] for a in want]
    assert [a[0] for a in attrs[:4]] == ["type", "index", "limit", "v"]          # creation order, visualize.cpp:24-28
    for got, ref in zip(attrs, want):
        assert got[1] == ref[1] and np.array_equal(got[2].ravel(), np.asarray(ref[2]).ravel()), got[0]
    d = {a[0]: a[2] for a in attrs}
    assert np.array_equal(d["index"].ravel(), p["id"].astype(np.int32))
    assert set(np.unique(d["debug"][:, 1])) <= {4.0, 5.0, 6.0}                   # jelly / water / sand codes
    water = np.asarray(G.GROUP_KINDS)[p["group"]] == 3
    assert np.array_equal(d["debug"][water, 0], p["ps"][water]) and not d["debug"][~water, 0].any()


@pytest.mark.skipif(not O.partio_ref_available(), reason="reference tree absent: golden files only")
@pytest.mark.parametrize("n,verbose", [(1, False), (65536, False), (65537, False), (70001, True)])
def test_writer_equals_reference_writer_live(tmp_path, n, verbose):
    p = G.golden_particles(n, 40 + n % 7)
    attrs = mpm.frame_attributes(p, G.GROUP_KINDS, verbose)
    mine, ref = tmp_path / "mine.bgeo", tmp_path / "ref.bgeo"
    bgeo.write_bgeo(str(mine), p["x"], attrs)
    d = {a[0]: a[2] for a in attrs}
    vb = {k: d[k] for k in ("m", "boundary_normal", "debug", "states", "boundary_distance", "near_boundary", "apic_frobenius_norm")} if verbose else None
    O.ref_write_partio(str(ref), p["x"], d["v"], d["type"], d["index"], d["limit"], vb)
    assert mine.read_bytes() == ref.read_bytes()


@pytest.mark.skipif(not O.ref_transfer_available(), reason="reference tree absent")
def test_writer_equals_write_partio_of_the_reference_solver_live(tmp_path):
    # the reference's MPM<3> (oracle/transfer_ref.cpp) steps a scene, deletes a particle, and dumps the frame with
    # its own write_partio; the product's writer, given the same particle state, produces the same bytes
    from tests import common as T
    from taichi_mpm_b200 import scenes
    scene, st = T.perturbed_scene(scenes.MAT_SAND, res=24, cells=3, seed=2)
    st["x"][0] = [6.5 / 24, 0.5, 0.5]
    s = O.RefSolver(scene, st)
    alive = s.substep(5)
    ref = tmp_path / "ref.bgeo"
    s.write_partio(ref)
    p = s.particles()
    s.close()
    ids = p["alive_ids"]
    assert alive == len(ids) == len(st["x"]) - 1
    d = dict(id=ids.astype(np.uint32), x=p["x"][ids], v=p["v"][ids], b=p["b"][ids], mass=st["mass"][ids], ps=p["ps"][ids],
             group=np.zeros(len(ids), np.int32))
    mine = tmp_path / "mine.bgeo"
    bgeo.write_bgeo(str(mine), d["x"], mpm.frame_attributes(d, [scenes.MAT_SAND], verbose=False))
    assert mine.read_bytes() == ref.read_bytes()


def test_read_rejects_other_files(tmp_path):
    f = tmp_path / "x.bgeo"
    f.write_bytes(b"PK\x03\x04 not a bgeo")
    with pytest.raises(ValueError):
        bgeo.read_bgeo(str(f))
