#!/bin/bash
# runs bench.py (kernel-only legs) once per scorer variant built into build/variants/*.so; prints ms/step + score kernel ms
for so in build/variants/*.so; do
  out=$(RG_LIB=$PWD/$so python bench.py --steps 10 --warmup 3 --no-extras --no-cpu 2>/dev/null | tail -1)
  python - "$so" "$out" <<'PY'
import json, sys
so, line = sys.argv[1], sys.argv[2]
try:
    d = json.loads(line)
    print(f"{so}: step {d['ms_per_step']:.3f} ms  score {d['roofline']['kernel_ms_per_launch']:.3f} ms  frac {d['roofline']['frac']:.3f}  fixup {d['phases_ms_per_step']['fixup_ms']:.3f}")
except Exception as e:
    print(so, "FAILED", line[:200])
PY
done
