#!/bin/bash
# quick check: full-IAN tests + pair tests, then the bench line (per-layer times of both models)
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_full.py tests/test_gpu_parity.py::test_pair_kernel_equals_one_cta_kernel -x -q 2>&1 | tail -8) > gpurun_out/quick_tests.log
cat gpurun_out/quick_tests.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
tail -2 gpurun_out/bench_quick.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_quick.json").read().strip().splitlines()[-1])
print(round(d["value"]), d["ms_per_step"], d["roofline"]["frac_burst"], d["roofline"]["layer_ms"], d["roofline"]["edge_kernel_ms"])
f = d["full_ian"]
print(" full bf16", f["bf16"]["value"], f["bf16"]["frac_burst"], f["bf16"]["layer_ms"]); print(" fp32", f["fp32_split"]["value"], f["fp32_split"]["layer_ms"])
print(" bf16 vs fp32", f["bf16_vs_fp32_max_abs"], f["bf16_vs_fp32_mean_abs"], f["bf16_vs_fp32_psnr_db"])
print(" edit", d["edit"]); print(" config5", d["config5"]["value"]); print(" e2e", d["e2e"]["value"], d["e2e"]["sync_value"], d["e2e"]["pageable_value"])
PY
