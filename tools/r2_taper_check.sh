#!/bin/bash
# pipeline fill / drain: first and last pass of a call cut into 1/8 .. 1/2 pieces, 512-pair sweep (= one rank's share at 8 GPUs)
run() {
  out=$(env "$@" python bench.py --pairs 512 --steps 5 --warmup 3 --no-extras --no-cpu --no-split --no-oracle-check 2>/dev/null | tail -1)
  python - "$*" "$out" <<'PY'
import json, sys
d = json.loads(sys.argv[2]); r = d["roofline"]
print(f"{sys.argv[1]:20s} step {d['ms_per_step']:.3f} e2e {d['e2e']['ms_per_step']:.3f}  score/launch {r['kernel_ms_per_launch']:.3f} launches {r['launches_timed']} frac {r['frac']:.4f} alone {r['kernel_alone']['frac']:.4f} serial step {r['kernel_alone']['ms_per_step_serial_passes']:.3f}")
PY
}
run RG_NO_TAPER=1
run RG_NO_TAPER=0
