#!/bin/bash
# accuracy + speed of the split (main|cross) vs merged TMEM accumulator modes
(timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -5)
for m in 0 1; do
  echo "=== IAN_TC_MERGED=$m"
  IAN_TC_MERGED=$m DIAG_N=24 timeout 400 python tools/diag_gpu.py 2>&1 | grep "^tc"
  IAN_TC_MERGED=$m timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > /tmp/line.json
  python - <<'PY'
import json
d = json.loads(open('/tmp/line.json').read())
print("ms/step %.3f  img/s %.0f  e2e %.0f  edit %.0f" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["edit"]["value"]), d["roofline"]["layer_ms"], "frac %.3f" % d["roofline"]["frac"], d["clocks"])
PY
done
