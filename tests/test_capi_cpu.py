"""CPU-only checks of the drop-in boundary: libmpmb.so builds, loads and exports every symbol
include/mpmb.h declares; no compute call is made (no GPU here)."""
import ctypes
import os
import re

import pytest

from taichi_mpm_b200 import capi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "mpmb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mpmb_[a-z_0-9]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build()
    assert os.path.exists(path)
    L = ctypes.CDLL(path)
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(L, name), "libmpmb.so does not export %s" % name
    assert sorted(capi.EXPORTS) == declared
    # ... and nothing else under the mpmb_ prefix (debug readers exist only in -DMPMB_DEBUG_EXPORTS builds)
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", path], stdout=subprocess.PIPE, text=True).stdout
    exported = sorted(line.split()[-1] for line in out.splitlines() if line.split()[-1].startswith("mpmb_"))
    assert exported == declared


def test_version_and_struct_sizes():
    L = capi.lib()
    assert L.mpmb_version() == 1
    # MpmbConfig: 3*4 + 4 + 4 + 12 + 4 + 4 + 4 (+pad) + 8 + 4*4 + 8 + 32
    assert ctypes.sizeof(capi.MpmbConfig) == 112
    assert ctypes.sizeof(capi.MpmbAosLayout) == 48


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.MpmbError) as ei:
        capi.Engine(32, 1.0 / 32, 1e-4)
    assert ei.value.code == -2   # MPMB_ERR_CUDA: the product path fails loudly, it never routes to the oracle


def test_create_rejects_bad_arguments():
    L = capi.lib()
    h = ctypes.c_void_p()
    assert L.mpmb_create(None, ctypes.byref(h)) == -1
    cfg = capi.MpmbConfig()
    cfg.res[:] = [4, 4, 4]
    cfg.dx, cfg.dt = 0.25, 1e-4
    assert L.mpmb_create(ctypes.byref(cfg), ctypes.byref(h)) == -1
    assert b"res" in L.mpmb_last_error(None)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "taichi_mpm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in text and "liboracle" not in text and "mpm_oracle" not in text, f


def test_host_mirror_kwargs_without_gpu():
    # keyword handling of the reference-facing mirror that does not need a device
    from taichi_mpm_b200.mpm import LevelSet
    ls = LevelSet((64, 64, 64), 1 / 64)
    ls.add_plane((0, 2, 0), -0.1)
    ls.set_friction(0.4)
    pl = ls.planes_grid_units()
    assert pl.shape == (1, 4) and pl[0, 1] == pytest.approx(1.0) and pl[0, 3] == pytest.approx(-6.4)


def test_header_is_plain_c(tmp_path):
    # the drop-in boundary is a C ABI: include/mpmb.h must compile as C99 (and as C++11) on its own, warnings as errors
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "hdr.c"
    src.write_text('#include "include/mpmb.h"\nint main(void) { MpmbConfig c; MpmbRigidBody r; (void)c; (void)r; return MPMB_VERSION ? 0 : 1; }\n')
    for cmd in (["gcc", "-std=c99"], ["g++", "-std=c++11", "-x", "c++"]):
        r = subprocess.run(cmd + ["-Wall", "-Wextra", "-pedantic", "-Werror", "-I", root, "-fsyntax-only", str(src)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 0, r.stdout
