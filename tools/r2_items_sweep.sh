#!/bin/bash
# scorer work items per resident block (planner target) with the 128 x 4 shape and pipelined passes, 512-pair sweep
run() {
  out=$(env "$@" python bench.py --pairs 512 --steps 4 --warmup 3 --no-extras --no-cpu --no-split --no-oracle-check 2>/dev/null | tail -1)
  python - "$*" "$out" <<'PY'
import json, sys
d = json.loads(sys.argv[2]); r = d["roofline"]
print(f"{sys.argv[1]:24s} step {d['ms_per_step']:.3f} e2e {d['e2e']['ms_per_step']:.3f}  score/launch {r['kernel_ms_per_launch']:.3f} frac {r['frac']:.4f} alone {r['kernel_alone']['kernel_ms_per_launch']:.3f} {r['kernel_alone']['frac']:.4f}")
PY
}
for v in 8 12 16 24 32 48 96; do run RG_ITEMS_PER_BLOCK=$v; done
