"""Compact one-line summary of a bench.py JSON line (stdin).  usage: python bench.py ... | tail -1 | python tools/bsum.py [tag]"""
import json, sys
tag = sys.argv[1] if len(sys.argv) > 1 else ""
try:
    d = json.loads(sys.stdin.read())
except Exception as e:
    print(tag, "bad json", e); sys.exit(0)
p = {k: round(v["ms_per_scan"], 3) for k, v in d.get("profile_chain0", {}).items() if v["ms_per_scan"] > 0}
print(tag, "chains", d["config"]["chains_per_gpu"], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1),
      "launches", d["gpu_launches"], p)
