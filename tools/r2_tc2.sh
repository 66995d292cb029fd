#!/bin/bash
# pair-kernel bring-up: the two dedicated tests first (bounded), then the suite, then a bench line with per-layer times
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py::test_pair_kernel_equals_one_cta_kernel tests/test_gpu_full.py::test_pair_kernel_on_the_flow_models -x -q 2>&1 | tail -25) > gpurun_out/tc2_tests.log
cat gpurun_out/tc2_tests.log
(timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8) > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_tc2.json 2> gpurun_out/bench_tc2.err
tail -2 gpurun_out/bench_tc2.err
IAN_TC2=0 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-config5 --no-edit --no-full > gpurun_out/bench_tc2off.json 2>/dev/null
python - <<'PY'
import json
for f in ("gpurun_out/bench_tc2.json", "gpurun_out/bench_tc2off.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), d["ms_per_step"], d["roofline"]["frac_burst"], d["roofline"]["layer_ms"])
        if d.get("full_ian"): print(" full bf16", d["full_ian"]["bf16"]["value"], d["full_ian"]["bf16"]["layer_ms"], "fp32", d["full_ian"]["fp32_split"]["value"])
        if d.get("edit"): print(" edit", d["edit"]["value"])
        if d.get("config5"): print(" config5", d["config5"]["value"])
    except Exception as e:
        print(f, "unreadable", e)
PY
