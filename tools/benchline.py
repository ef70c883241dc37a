#!/usr/bin/env python
"""Condense bench.py's JSON line (stdin) to one short line:  python bench.py ... | python tools/benchline.py TAG"""
import json, sys
tag = sys.argv[1] if len(sys.argv) > 1 else ""
txt = sys.stdin.read().strip().splitlines()
d = json.loads(txt[-1])
r = d.get("roofline", {})
st = r.get("stage_ms_per_step", {})
print(tag, "value=%.3e" % d["value"], "ms/step=%.3f" % d["ms_per_step"], "e2e=%.3e" % d["e2e"]["value"],
      "stages(ms):", " ".join(f"{k}={v:.3f}" for k, v in st.items()), "launch_ms=%.4f" % r.get("launch_ms", 0),
      "frac=%.3f" % r.get("frac", 0), "launches=%d" % d.get("gpu_launches", 0), "chk=%.4f" % d["e2e"].get("density_checksum", 0))
