#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -8) > gpurun_out/sk_tests.log
cat gpurun_out/sk_tests.log
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-config5 --no-full"
for i in 1 2; do
timeout 600 $B > gpurun_out/bench_sk_on$i.json 2>/dev/null
IAN_STREAMK=0 timeout 600 $B > gpurun_out/bench_sk_off$i.json 2>/dev/null
done
IAN_TC2=0 timeout 600 $B > gpurun_out/bench_tc2_off.json 2>/dev/null
python - <<'PY'
import json
for f in ("sk_on1", "sk_off1", "sk_on2", "sk_off2", "tc2_off"):
    try:
        d = json.loads(open("gpurun_out/bench_%s.json" % f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"], 4), round(d["roofline"]["frac_burst"], 4), d["roofline"]["layer_ms"], "edit", round(d["edit"]["value"]))
    except Exception as e:
        print(f, "unreadable", e)
PY
