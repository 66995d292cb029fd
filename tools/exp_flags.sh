#!/bin/bash
# perf experiment: bench with IAN_DEBUG_FLAGS variants (results invalid for flags != 0)
for f in ${FLAGS:-0 16 32 48}; do
  IAN_DEBUG_FLAGS=$f timeout 200 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-edit 2>&1 | tail -1 > /tmp/line.json
  python - "$f" <<'PY'
import sys, json
d = json.loads(open('/tmp/line.json').read())
print("flags", sys.argv[1], "ms/step %.3f" % d["ms_per_step"], d["roofline"]["layer_ms"], d["clocks"])
PY
done
