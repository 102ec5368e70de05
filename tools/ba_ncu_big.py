"""One device bundle adjustment of a synthetic 36-view / 20 000-point / 160 000-observation scene (ncu launch lists)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tsbb15_b200 as rg  # noqa: E402

sc = rg.synth.ba_scene(36, 20000, track=8)
r = rg.runtime.bundle_adjust(*sc[:5], ftol=1e-6)
print(r["cost"], r["iters"], r["status"])
