"""The -DRG_DEBUG build of the library (device-side bounds / invariant assertions in the scorer, the flag list, the fix-up
and the solvers; compute-sanitizer is closed on this pool) runs a medley of ragged, multi-pass, overflowing and degenerate
calls in a subprocess without tripping an assertion, and gives the same answers as the release build."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEBUG_LIB = os.path.join(ROOT, "tsbb15-3d-reconstruction-project_b200", "librg_b200_debug.so")

SCRIPT = r"""
import json, sys
import numpy as np
sys.path.insert(0, %r)
import tsbb15_b200 as rg
from tsbb15_b200 import runtime as rt, sampling, synth
out = {}
sizes = [900, 1500, 64, 2100, 8, 1000, 33000]
pts = [synth.two_view(n, seed=40 + k)[0] for k, n in enumerate(sizes)]
r = rt.f_ransac_batched(pts, None, n_hyp=[96, 700, 33, 1100, 5, 513, 300], sample_seed=3, want_counts=True)
out["f"] = [r["best_idx"].tolist(), r["best_count"].tolist(), [int(c.sum()) for c in r["counts"]]]
rt.set_option(6, 1)
r2 = rt.f_ransac_batched(pts, None, n_hyp=[96, 700, 33, 1100, 5, 513, 300], sample_seed=3)
rt.set_option(6, 0)
rt.set_option(8, 50)
r3 = rt.f_ransac_batched(pts, None, n_hyp=[96, 700, 33, 1100, 5, 513, 300], sample_seed=3)
rt.set_option(8, 0)
out["multipass_equal"] = bool(np.array_equal(r2["best_idx"], r["best_idx"]) and np.array_equal(r3["best_count"], r["best_count"]))
bad = pts[0].copy(); bad[5] = np.nan
rb = rt.f_ransac_batched([bad, pts[1][:0]], [sampling.fast(900, 64, 8, seed=1), np.zeros((0, 8), np.int32)])
out["nan"] = rb["best_idx"].tolist()
for n in (6, 7, 8):
    X, y, _ = synth.pnp_scene(3000 + n, seed=n, sigma_px=0.3, outlier_frac=0.25)
    for solver in (0, 1):
        rt.set_option(7, solver)
        p = rt.pnp_ransac(X, y, sampling.fast(3000 + n, 301, n, seed=n), (1.5 / 3217.0) ** 2, want_counts=True)
        out["pnp%%d_%%d" %% (n, solver)] = [p["best_idx"], p["best_count"], int(p["counts"].sum())]
rt.set_option(7, 0)
views = [synth.pnp_scene(200 + 31 * v, seed=v)[:2] for v in range(4)]
vb = rt.pnp_ransac_batched([v[0] for v in views], [v[1] for v in views],
                           [sampling.fast(len(v[0]), 64 + 7 * k, 6, seed=k) for k, v in enumerate(views)], (1.5 / 3217.0) ** 2)
out["pnp_batched"] = vb["best_count"].tolist()
print("RESULT " + json.dumps(out))
""" % ROOT


def _run(lib=None):
    env = dict(os.environ)
    if lib:
        env["RG_LIB"] = lib
    else:
        env.pop("RG_LIB", None)
    res = subprocess.run([sys.executable, "-c", SCRIPT], capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-3000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("RESULT ")][-1]
    return json.loads(line[7:])


def test_debug_build_runs_clean_and_agrees_with_release():
    if not os.path.isfile(DEBUG_LIB):
        pytest.skip("librg_b200_debug.so not built (python __graft_entry__.py builds it)")
    dbg = _run(DEBUG_LIB)
    rel = _run(None)
    assert dbg == rel
    assert dbg["multipass_equal"] and dbg["nan"][1] == -1
