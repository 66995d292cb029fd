#!/bin/bash
# final evidence of round 2 on ONE box: GPU suite with recorded parity numbers, the default bench line, ncu launch lists +
# --set full summaries, compute-sanitizer (tools/r2_profile.sh)
mkdir -p gpurun_out/rec
export IAN_TEST_RECORD=$PWD/gpurun_out/rec
(timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -25) > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log | cut -c1-400
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -2 gpurun_out/bench_final.err | cut -c1-300
python tools/bench_brief.py final gpurun_out/bench_final.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_reference.json 2> /dev/null
tail -c 600 gpurun_out/bench_final_reference.json
SKIP_FULL_NCU=1 bash tools/r2_profile.sh
