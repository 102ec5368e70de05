#!/bin/bash
# pass pipelining (option 10): the fix-up grid beside the next pass's scorer against the per-pass phase times, 1024-pair sweep
run() {
  out=$(env "$@" python bench.py --pairs 1024 --steps 3 --warmup 3 --no-extras --no-cpu --no-split --no-oracle-check 2>/dev/null | tail -1)
  python - "$*" "$out" <<'PY'
import json, sys
d = json.loads(sys.argv[2]); p = d["phases_ms_per_pass"]
print(f"{sys.argv[1]:44s} step {d['ms_per_step']:.3f} e2e {d['e2e']['ms_per_step']:.3f}  score {p['score_ms']:.3f} solve {p['solve_ms']:.3f} fixup {p['fixup_ms']:.3f} select {p['select_ms']:.3f}")
PY
}
run RG_NOPIPE=1
run RG_TAIL_GRID=148
run RG_TAIL_GRID=111
run RG_TAIL_GRID=222
