#!/bin/bash
# full config-5 sweep, device-resident against host entry point, serial passes (RG_NOPIPE set) against pipelined ones (default)
run() {
  out=$(env "$@" python bench.py --steps 5 --warmup 3 --no-extras --no-cpu --no-split --no-oracle-check 2>/dev/null | tail -1)
  python - "$*" "$out" <<'PY'
import json, sys
d = json.loads(sys.argv[2]); r = d["roofline"]
print(f"{sys.argv[1]:20s} step {d['ms_per_step']:.3f} e2e {d['e2e']['ms_per_step']:.3f}  score/launch {r['kernel_ms_per_launch']:.3f} frac {r['frac']:.4f} alone {r['kernel_alone']['frac']:.4f}")
PY
}
run RG_NOPIPE=1
run RG_DUMMY=0
