"""The C-ABI driven from a pure C++ host (examples/host_substep.cpp): no Python, no torch in the
process that owns the engine.  Result must agree with the Python-driven engine."""
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.needs_cuda]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_host_matches_python_host(tmp_path):
    from taichi_mpm_b200 import build, capi, scenes
    lib = build.build()
    exe = str(tmp_path / "host_substep")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-O2", "-std=c++17", os.path.join(ROOT, "examples", "host_substep.cpp"), "-I" + os.path.join(ROOT, "include"),
                           "-L" + os.path.dirname(lib), "-lmpmb", "-Wl,-rpath," + os.path.dirname(lib), "-o", exe])
    out = subprocess.check_output([exe, "40"], text=True)
    m = re.search(r"alive=(\d+) com=\(([-\d.e]+) ([-\d.e]+) ([-\d.e]+)\) mean_vy=([-\d.e]+)", out)
    assert m, out
    alive, com, vy = int(m.group(1)), np.array([float(m.group(i)) for i in (2, 3, 4)]), float(m.group(5))
    # same scene through the Python host
    res, dx = 64, 1.0 / 64
    x, mass, vol = scenes.lattice_block(res, (24, 10, 24), (40, 26, 40))
    e = capi.Engine(res, dx, 2e-5, (0, -10, 0))
    e.set_material(0, scenes.MAT_SAND, scenes.material_params(scenes.MAT_SAND))
    e.set_planes(np.array([[0, 1, 0, -10.0]], np.float32), 0.4)
    e.upload(x, np.zeros_like(x), mass, vol)
    e.substep(40)
    p = e.download()
    assert alive == len(p["id"]) == len(x)
    assert np.abs(p["x"].astype(np.float64).mean(0) - com).max() < 1e-6
    assert abs(p["v"][:, 1].astype(np.float64).mean() - vy) < 1e-5
    e.close()
