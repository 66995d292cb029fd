#!/bin/bash
# exploration call 4: default build (PDL on, seed kernel with smem weights) + pair-kernel split-K behind IAN_TC2_SPLITK=1
mkdir -p gpurun_out/rec
export IAN_TEST_RECORD=$PWD/gpurun_out/rec
(timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -25) > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log | cut -c1-400
(IAN_TC2_SPLITK=1 timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -25) > gpurun_out/pytest_gpu_sk.log
echo "== IAN_TC2_SPLITK=1"; cat gpurun_out/pytest_gpu_sk.log | cut -c1-400
bash tools/r2_ab.sh "new new+sk" 2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/x4_launches_edit_b128.csv python tools/edit_once.py > /dev/null 2>&1
IAN_TC2_SPLITK=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/x4_launches_sk.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-edit --no-full --no-config5 > /dev/null 2>&1
