"""Regenerates tests/golden/frame_*.bgeo with the REFERENCE's own writer: the Partio copy vendored
under /root/reference/external/partio, compiled in place by `make -C oracle ref`, driven as
MPM<dim>::write_partio drives it (src/visualize.cpp:16-100; oracle/partio_ref.cpp).
Run in the build container (the reference tree is not present on the GPU box):

    python tests/golden/make_bgeo_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyoracle as O  # noqa: E402


def golden_particles(n, seed):
    """Deterministic stand-in for a `download()` dict (ids not contiguous, as after deletions)."""
    rng = np.random.default_rng(seed)
    return dict(
        id=np.sort(rng.choice(4 * n + 7, size=n, replace=False)).astype(np.uint32),
        x=(0.1 + 0.8 * rng.random((n, 3))).astype(np.float32),
        v=rng.normal(size=(n, 3)).astype(np.float32),
        b=(rng.normal(size=(n, 9)) * 1e-3).astype(np.float32),
        mass=(1e-6 * (1 + rng.random(n))).astype(np.float32),
        ps=(1 + 0.01 * rng.normal(size=n)).astype(np.float32),
        group=rng.integers(0, 3, size=n).astype(np.int32),
    )


GROUP_KINDS = [4, 3, 1]  # sand, water, jelly


def main():
    from taichi_mpm_b200 import mpm
    for name, n, seed, verbose in (("frame_plain_17", 17, 1, False), ("frame_verbose_9", 9, 2, True), ("frame_empty", 0, 3, False)):
        p = golden_particles(n, seed)
        attrs = {a[0]: a[2] for a in mpm.frame_attributes(p, GROUP_KINDS, verbose)}   # VALUES come from the product's host code;
        vb = None                                                                   # the BYTES from the reference's writer
        if verbose:
            vb = {k: attrs[k] for k in ("m", "boundary_normal", "debug", "states", "boundary_distance", "near_boundary", "apic_frobenius_norm")}
        O.ref_write_partio(os.path.join(HERE, name + ".bgeo"), p["x"], attrs["v"], attrs["type"], attrs["index"], attrs["limit"], vb)
        print(name, os.path.getsize(os.path.join(HERE, name + ".bgeo")), "bytes")


if __name__ == "__main__":
    main()
