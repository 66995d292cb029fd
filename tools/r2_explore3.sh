#!/bin/bash
# exploration call 3: full suite on the default build and with IAN_PDL=1; same-box A/B head vs new vs new+pdl
mkdir -p gpurun_out/rec
export IAN_TEST_RECORD=$PWD/gpurun_out/rec
(timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -25) > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log | cut -c1-400
(IAN_PDL=1 timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -25) > gpurun_out/pytest_gpu_pdl.log
echo "== IAN_PDL=1"; cat gpurun_out/pytest_gpu_pdl.log | cut -c1-400
bash tools/r2_ab.sh "${AB_TAGS:-head new new+pdl}" 2
IAN_PDL=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/x3_launches_edit_b128.csv python tools/edit_once.py > /dev/null 2>&1
