#!/bin/bash
# GPU check: parity tests, error statistics, bench line
(timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -6)
DIAG_N=${DIAG_N:-16} timeout 400 python tools/diag_gpu.py 2>&1 | grep -E "^(tc|simt)"
timeout 300 python bench.py --steps 30 --warmup 5 ${BENCH_ARGS:---no-cpu-baseline} 2>&1 | tail -1 > gpurun_out/bench_last.json
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_last.json').read())
print("ms/step %.3f  img/s %.0f  e2e %.0f  edit %s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["edit"] and round(d["edit"]["value"])), d["roofline"]["layer_ms"], d["roofline"]["edge_kernel_ms"], "frac %.3f" % d["roofline"]["frac"], "launches", d["gpu_launches"], d["clocks"])
f = d.get("full_ian") or {}
print("full_ian bf16 %.0f img/s, fp32 %.0f img/s" % (f.get("value", 0), (f.get("fp32_split") or {}).get("value", 0)), "| latency b1:", d.get("single_image_latency"))
PY
