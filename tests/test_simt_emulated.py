"""The engine's CUDA source on a CPU: tests/simt compiles taichi_mpm_b200/csrc/mpmb_engine.cu (kernels, device math
AND the C-ABI host code; only kernel launches and inline PTX are rewritten, see tests/simt/build_simt.py) on top of
a small SIMT emulator — CUDA threads as coroutines, real __syncthreads / warp-collective rendezvous, shared memory,
atomics, cp.async as plain copies.  The gpu-marked parity tests then run against that library through the same
ctypes binding (MPMB_SIMT=1).  This checks kernel LOGIC (indexing, barriers, orderings, host orchestration), not
codegen, timing or memory-model races: the `-m gpu` run on a B200 stays the parity gate; this run means a logic
error is caught in the CPU suite, and build-time kernel experiments can be debugged without GPU minutes:

    MPMB_SIMT=1 MPMB_SIMT_DEFINES=<NAME> python -m pytest tests -m gpu -q
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env, args):
    env = dict(os.environ, MPMB_SIMT="1", **extra_env)
    env.pop("MPMB_LIB", None)
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-m", "gpu", "-p", "no:cacheprovider"] + args, cwd=ROOT, env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    tail = "\n".join(r.stdout.strip().splitlines()[-15:])
    assert r.returncode == 0, tail
    return tail


def test_gpu_parity_suite_passes_on_the_simt_emulator():
    # everything that does not need a real device, minus the two longest multi-step runs (they pass too: ~1 min more)
    tail = _run({}, ["tests/test_gpu_parity.py", "tests/test_gpu_zz_reference_golden.py", "tests/test_gpu_mpm_mirror.py",
                     "tests/test_gpu_zz_frame_io.py", "tests/test_gpu_rigid.py", "tests/test_gpu_async.py", "-k", "not many_movers and not multi_step_invariants"])
    assert " passed" in tail and "failed" not in tail


import pytest  # noqa: E402


_ORDER_SCRIPT = r"""
import os, sys
sys.path.insert(0, %r)
import numpy as np
from tests.simt import build_simt
defines = [d for d in os.environ.get("MPMB_SIMT_DEFINES", "").split(",") if d]
os.environ["MPMB_LIB"] = build_simt.build(defines)
from tests import common as T
from tests.test_gpu_slab import _scene
scene, st = _scene()           # fast motion: dozens of particles change tile every substep
e = T.make_engine(scene, st)
e.substep(10)
d = e.download()
e.close()
np.savez(sys.argv[1], **d)
"""


@pytest.mark.parametrize("defines", [""])   # build-time variants (MPMB_SIMT_DEFINES) can be added here while they are being developed
def test_results_do_not_depend_on_cta_or_thread_scheduling_order(tmp_path, defines):
    # the emulator runs CTAs and threads in index order, reversed, or pseudo-randomly shuffled per launch
    # (MPMB_SIMT_ORDER): a result that depended on who runs first — an inter-CTA race such as two CTAs writing one
    # outpos entry, an unordered float reduction — would change bits.  The engine claims bit-reproducibility.
    import numpy as np
    outs = []
    for order in ("0", "1", "2"):
        out = str(tmp_path / ("order%s.npz" % order))
        env = dict(os.environ, MPMB_SIMT="1", MPMB_SIMT_ORDER=order, MPMB_SIMT_DEFINES=defines)
        env.pop("MPMB_LIB", None)
        r = subprocess.run([sys.executable, "-c", _ORDER_SCRIPT % ROOT, out], cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 0, r.stdout[-2000:]
        outs.append(np.load(out))
    for k in outs[0].files:
        assert np.array_equal(outs[0][k], outs[1][k]) and np.array_equal(outs[0][k], outs[2][k]), k


_EXAMPLES_SCRIPT = r"""
import os, sys, importlib.util
sys.path.insert(0, %r)
from tests.simt import build_simt
os.environ["MPMB_LIB"] = build_simt.build()
for name, args in (("rigid_paddle", (1, 32)), ("async_snow", (1,))):
    spec = importlib.util.spec_from_file_location(name, os.path.join(%r, "examples", name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    m.main(*args)
"""


def test_example_scripts_run_on_the_emulator():
    # examples/rigid_paddle.py (a scripted paddle through sand: CPIC) and examples/async_snow.py (the AsyncMPM scheduler)
    env = dict(os.environ, MPMB_SIMT="1")
    env.pop("MPMB_LIB", None)
    r = subprocess.run([sys.executable, "-c", _EXAMPLES_SCRIPT % (ROOT, ROOT)], cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "coloured" in r.stdout and "time levels 8..32 units" in r.stdout, r.stdout[-500:]
