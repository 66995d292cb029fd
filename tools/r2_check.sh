#!/bin/bash
# round-2 GPU check: GPU test suite (+ recorded parity numbers), default bench line
mkdir -p gpurun_out/rec
export IAN_TEST_RECORD=$PWD/gpurun_out/rec
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
(timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15) > gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
cat gpurun_out/pytest_gpu.log; tail -3 gpurun_out/bench_n1.err; cut -c1-1500 gpurun_out/bench_n1.json
