#!/bin/bash
mkdir -p gpurun_out/rec
export IAN_TEST_RECORD=$PWD/gpurun_out/rec
(timeout 1200 python -m pytest tests/test_gpu_full.py -x -q 2>&1 | tail -30) > gpurun_out/grad_tests.log
cat gpurun_out/grad_tests.log | cut -c1-400
(timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -12) > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log | cut -c1-400
cat gpurun_out/rec/flow_brush_*.json
