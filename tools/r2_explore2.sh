#!/bin/bash
# exploration call 2: conv1 (TMA-store epilogue) + cooperative split-K finalize in the default build; PDL behind IAN_PDL=1
mkdir -p gpurun_out/rec
export IAN_TEST_RECORD=$PWD/gpurun_out/rec
(timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -25) > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log | cut -c1-400
(IAN_PDL=1 timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -25) > gpurun_out/pytest_gpu_pdl.log
echo "== IAN_PDL=1"; cat gpurun_out/pytest_gpu_pdl.log | cut -c1-400
bash tools/r2_ab.sh "head new new+pdl" 2
IAN_PDL=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/x2_launches_edit_b128.csv python tools/edit_once.py > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'conv1_tc|splitk_finalize' -c 6 -o gpurun_out/x2_conv1 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-edit --no-full --no-config5 > gpurun_out/x2_ncu.log 2>&1
ls -la gpurun_out/x2*.ncu-rep
