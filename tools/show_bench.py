"""Pretty-print the interesting parts of a bench.py JSON line."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
for k in ["metric", "value", "ms_per_step", "e2e", "gpu_launches", "roofline", "phases_ms_per_step", "guard_band",
          "clocks", "cpu_baseline"]:
    if k in d:
        print(k, json.dumps(d[k])[:900])
for k in d:
    if k.startswith("config") and k != "config":
        v = d[k]
        if "epi_max_thr1.5" in v:
            print(k, json.dumps({kk: v[kk] for kk in ("epi_max_thr1.5", "epi_max_thr0.25", "sampson_thr1.5")})[:1200])
        else:
            print(k, json.dumps(v)[:900])
