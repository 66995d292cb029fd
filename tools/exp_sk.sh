#!/bin/bash
(timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -4)
for s in 0 1; do
  IAN_STREAMK=$s timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > /tmp/line.json
  python - "$s" <<'PY'
import json, sys
d = json.loads(open('/tmp/line.json').read())
f = d.get("full_ian") or {}
print("STREAMK=%s ms/step %.3f img/s %.0f edit %.0f frac %.3f" % (sys.argv[1], d["ms_per_step"], d["value"], d["edit"]["value"], d["roofline"]["frac"]), d["roofline"]["layer_ms"])
print("   full bf16 %.0f (%.2f ms) fp32 %.0f (%.2f ms)" % (f["value"], f["bf16"]["ms_per_step"], f["fp32_split"]["value"], f["fp32_split"]["ms_per_step"]))
PY
done
