"""What the tile kernels see in a developed flow (run under gpurun, one GPU):
rows per tile (run + arrivals, holes included), particles per cell, and the balance of the cell-owner P2G
(mean / max particles per cell within each warp's 32 cells).   python profiles/flow_stats.py [--scale 0.5] [--develop 20000]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from taichi_mpm_b200 import capi, scenes  # noqa: E402


def stats(e, res, tag):
    d = e.download(sort_by_id=False)
    X = d["x"].astype(np.float32) * np.float32(res)
    base = (X - np.float32(0.5)).astype(np.int32)
    tile = base >> 2
    nt = (res + 1 + 3) // 4 + 1
    tid = (tile[:, 0] * nt + tile[:, 1]) * nt + tile[:, 2]
    cell = ((base[:, 0] & 3) << 4) | ((base[:, 1] & 3) << 2) | (base[:, 2] & 3)
    per_tile = np.bincount(tid)
    per_tile = per_tile[per_tile > 0]
    key = tid.astype(np.int64) * 64 + cell
    tiles = np.unique(tid)
    remap = np.searchsorted(tiles, tid)
    cc = np.bincount(remap * 64 + cell, minlength=len(tiles) * 64).reshape(len(tiles), 2, 32)   # [tile][warp][lane]
    mx = cc.max(2)
    mean = cc.mean(2)
    out = {"state": tag, "particles": int(len(X)), "tiles": int(len(tiles)),
           "rows_per_tile": {"mean": float(per_tile.mean()), "p50": float(np.percentile(per_tile, 50)), "p90": float(np.percentile(per_tile, 90)),
                             "max": int(per_tile.max()), "frac_tiles_over_512": float((per_tile > 512).mean()), "frac_tiles_over_640": float((per_tile > 640).mean()),
                             "frac_tiles_under_64": float((per_tile < 64).mean())},
           "particles_per_cell": {"mean_nonempty": float(cc[cc > 0].mean()), "p99": float(np.percentile(cc[cc > 0], 99)), "max": int(cc.max())},
           "p2g_cell_owner_balance": {"sum_of_warp_max": int(mx.sum()), "sum_of_warp_mean": float(mean.sum()),
                                      "efficiency_mean_over_max": float(mean.sum() / mx.sum())}}
    print(json.dumps(out))
    return out


def main():
    scale = float(next((sys.argv[i + 1] for i, a in enumerate(sys.argv) if a == "--scale"), 0.5))
    develop = int(next((sys.argv[i + 1] for i, a in enumerate(sys.argv) if a == "--develop"), 20000))
    cfg = scenes.config("sand256", scale, state=False)
    sc, m = cfg["scene"], cfg["meta"]
    e = capi.Engine(sc["res"], sc["dx"], sc["dt"], sc["gravity"], 1, True)
    e.set_material(0, m["kind"], sc["mat_params"][0])
    e.set_planes(sc["planes"], sc["friction"])
    e.seed_lattice(m["lo"], m["hi"], m["vol"], m["mass"], jitter=m["jitter"], seed=m["seed"])
    e.substep(20)
    stats(e, m["res"], "quiescent")
    for k in range(4):
        e.substep(develop // 4)
        stats(e, m["res"], "after %d substeps" % ((k + 1) * (develop // 4) + 20))


if __name__ == "__main__":
    main()
