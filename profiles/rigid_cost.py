"""Cost of the CPIC rigid-coupled path at BASELINE config 3's size (256^3 grid, 8.0 M sand), run under gpurun:

    python profiles/rigid_cost.py [--steps 100]

The resting column of the headline bench with a scripted paddle (a 0.5 x 0.25 plate of boundary samples) pushed into it.
Prints one JSON line: ms per substep without bodies (graph replay off, like the coupled loop), with the body (one
mpmb_substep(h, 1) per pose, state set / read back every substep as a host does), the per-stage split of the coupled substep,
and how many tiles took the rigid kernels."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from taichi_mpm_b200 import capi, scenes  # noqa: E402


def main():
    steps = int(next((sys.argv[i + 1] for i, a in enumerate(sys.argv) if a == "--steps"), "100"))
    cfg = scenes.config("sand256", state=False)
    sc, m = cfg["scene"], cfg["meta"]
    out = {}
    for mode in ("plain", "coupled"):
        e = capi.Engine(sc["res"], sc["dx"], sc["dt"], sc["gravity"], sc["particle_gravity"], True, no_graph=True)
        e.set_material(0, m["kind"], sc["mat_params"][0])
        e.set_planes(sc["planes"], sc["friction"])
        n = e.seed_lattice(m["lo"], m["hi"], m["vol"], m["mass"], jitter=m["jitter"], seed=m["seed"])
        rigid = None
        if mode == "coupled":
            plate = dict(tris=scenes.plate_mesh(0.25, 0.125, axis=0), position=(0.42, 0.25, 0.5), rotation=scenes.euler_rotation((0, 0, 12.0)),
                         velocity=(1.0, 0.0, 0.0), friction=0.3)
            rigid = scenes.make_rigid([plate], sc["dx"], penalty=1e3)
            e.set_rigid(rigid)
            out["samples"] = int(len(rigid["sample_rigid"]))
        e.substep(20)
        e.synchronize()
        e.set_profiling(True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for s in range(steps):
            if rigid is not None:
                rigid["position"][1] = rigid["position"][1] + rigid["velocity"][1] * sc["dt"]
                e.set_rigid_state(rigid)
            e.substep(1)
            if rigid is not None:
                e.get_rigid_state(2)
        e.synchronize()
        wall = (time.perf_counter() - t0) / steps * 1e3
        ms, launches = e.get_profile(reset=True)
        names = ["sort+tiles(+cdf)", "p2g", "g2p", "exchange", "grid"]
        out[mode] = dict(wall_ms_per_substep=round(wall, 4), device_ms_per_substep=round(float(sum(ms)) / steps, 4),
                         stage_ms={k: round(float(v) / steps, 4) for k, v in zip(names, ms)}, launches_per_substep=float(sum(launches)) / steps)
        if rigid is not None:
            pc = e.get_particle_cdf(n)
            out["coloured_particles"] = int((pc["states"] != 0).sum())
            out["near_boundary_particles"] = int(pc["near"].sum())
        out["particles"] = int(n)
        e.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
