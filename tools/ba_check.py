"""GPU check / timing of the device bundle adjustment on the golden Dino scenes (run on the B200 box)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tsbb15_b200 as rg  # noqa: E402
from oracle import ba_path as oba  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "ba_golden.npz"))
rt = rg.runtime
out = {}
for nv in (3, 8, 36):
    p = f"v{nv}_"
    cams, pts, uv, ci, pi = g[p + "cams0"], g[p + "pts0"], g[p + "uv"], g[p + "cam_idx"], g[p + "pt_idx"]
    for ftol in (1e-4, 1e-9):
        r = rt.bundle_adjust(cams, pts, uv, ci, pi, ftol=ftol)
        row = dict(cost=r["cost"], iters=r["iters"], status=r["status"], scipy_cost=float(g[p + "scipy_cost"]),
                   cost_check=oba.cost(r["cams"], r["pts"], uv, ci, pi))
        for l2 in (0, 1):
            rt.set_option(4, l2)
            for cl in (1, 2, 4, 8):
                rt.set_option(3, cl)
                rt.bundle_adjust(cams, pts, uv, ci, pi, ftol=ftol)
                t = time.perf_counter()
                for _ in range(5):
                    rt.bundle_adjust(cams, pts, uv, ci, pi, ftol=ftol)
                row[f"ms_{'l2' if l2 else 'dsmem'}_cluster{cl}"] = (time.perf_counter() - t) / 5 * 1e3
        rt.set_option(3, 0)
        rt.set_option(4, 0)
        out[f"v{nv}_ftol{ftol:g}"] = row
        print(nv, ftol, row, flush=True)
    if nv <= 8:
        tr = []
        Co, Xo, co, ito, sto = oba.bundle_adjust_lm(cams, pts, uv, ci, pi, ftol=1e-9, trace=tr)
        print("  numpy LM:", co, ito, sto)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ba_check.json"), "w"), indent=1)
