#!/bin/bash
# runs bench.py (kernel-only legs, 256-pair sweep) once per scorer variant built into build/variants/*.so; prints ms/step + score kernel ms
for so in build/variants/*.so; do
  out=$(RG_LIB=$PWD/$so python bench.py --pairs 256 --steps 3 --warmup 3 --no-extras --no-cpu --no-split --no-oracle-check 2>/dev/null | tail -1)
  python - "$so" "$out" <<'PY'
import json, sys
so, line = sys.argv[1], sys.argv[2]
try:
    d = json.loads(line)
    print(f"{so}: step {d['ms_per_step']:.3f} ms  score {d['roofline']['kernel_ms_per_launch']:.3f} ms  frac {d['roofline']['frac']:.3f}")
except Exception as e:
    print(so, "FAILED", line[:200])
PY
done
