#!/bin/bash
# exploration call: GPU tests of the new build, same-box A/B against earlier builds, ncu of the SIMT edge kernels
mkdir -p gpurun_out/rec
export IAN_TEST_RECORD=$PWD/gpurun_out/rec
(timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -25) > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log | cut -c1-400
bash tools/r2_ab.sh "head st128 new" 2
# launch list of the edit loop with the new build + what bounds the per-pixel kernels
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/x1_launches_edit_b128.csv python tools/edit_once.py > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'brush_seed|splitk_finalize|brush_update' -c 6 -o gpurun_out/x1_edit_edge python tools/edit_once.py > gpurun_out/x1_ncu_edit.log 2>&1
FULL_PREC=bf16 FULL_IT=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'head_g_kernel|head_b_out|head_r_kernel|made_iaf' -s 4 -c 4 -o gpurun_out/x1_head_edge python tools/full_once.py > gpurun_out/x1_ncu_head.log 2>&1
ls -la gpurun_out/*.ncu-rep
