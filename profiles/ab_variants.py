"""A/B of build-time kernel experiments on one GPU box (run under gpurun).

  python profiles/ab_variants.py TILE_XYZ DUAL_ARENA TILE_XYZ+DUAL_ARENA [--parity] [--steps 200]

Builds libmpmb_<variant>.so with -DMPMB_EXP_<NAME> for every '+'-joined name, then runs the default
library and each variant through `bench.py` (device-timed substeps only: --frames 0 --no-cpu-baseline)
in the SAME process order twice (A B C A B C) so that box-to-box and warm-up differences cancel, and
prints ms/substep with the per-stage split.  --parity also runs the GPU parity tests on each variant.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from taichi_mpm_b200 import build  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    steps = next((sys.argv[i + 1] for i, a in enumerate(sys.argv) if a == "--steps"), "200")
    reps = next((sys.argv[i + 1] for i, a in enumerate(sys.argv) if a == "--reps"), "2")
    args = [a for a in args if a not in (steps, reps)]
    libs = {"default": build.build()}
    for v in args:
        if ":" in v:   # DEFINE[+DEFINE...]:libname  (full macro names, NAME=VALUE allowed)
            dd, name = v.split(":")
            defs = dd.split("+")
        else:
            defs, name = ["MPMB_EXP_" + n for n in v.split("+")], v.lower().replace("+", "_")
        libs[v] = build.build(defines=defs, out=os.path.join(os.path.dirname(build.LIB), "libmpmb_" + name + ".so"))
    rows = {k: [] for k in libs}
    for rep in range(int(reps)):
        for name, lib in libs.items():
            env = dict(os.environ, MPMB_LIB=lib)
            if "--parity" in sys.argv and rep == 0 and name != "default":
                r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_parity.py", "-x", "-q", "-m", "gpu"], cwd=ROOT, env=env,
                                   stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
                print("parity[%s]: %s" % (name, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.returncode), flush=True)
            r = subprocess.run([sys.executable, "bench.py", "--steps", steps, "--warmup", "20", "--frames", "0", "--no-cpu-baseline"], cwd=ROOT,
                               env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
            try:
                line = json.loads(r.stdout.strip().splitlines()[-1])
                st = line.get("states") or {}
                both = {k: (round(v["ms_per_step"], 4), {a: round(b, 4) for a, b in v["stage_ms_per_step"].items()}) for k, v in st.items() if isinstance(v, dict)}
                rows[name].append((line["ms_per_step"], both or line["roofline"]["stage_ms_per_step"]))
            except Exception:
                rows[name].append((None, r.stderr[-300:]))
            print(name, rows[name][-1], flush=True)
    print(json.dumps(rows))


if __name__ == "__main__":
    main()
