#!/bin/bash
(timeout 900 python -m pytest tests/test_gpu_full.py -x -q -m gpu 2>&1 | tail -30)
timeout 300 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-edit 2>&1 | tail -1 > gpurun_out/bench_full.json
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_full.json').read())
print("simple ms/step %.3f" % d["ms_per_step"]); print("full_ian:", d.get("full_ian"))
PY
