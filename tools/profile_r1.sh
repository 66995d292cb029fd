#!/bin/bash
# ncu evidence for profiles/: launch list of a bench step + full capture of the dominant kernels
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -3) > gpurun_out/pytest_gpu.log
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/launches_r1b.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-edit > gpurun_out/ncu_b.log 2>&1
timeout 800 ncu --set full --clock-control none --import-source on -k regex:'tapgemm_tc|decout_tc' -s 20 -c 10 -o gpurun_out/prof_r1b \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-edit > gpurun_out/ncu_f.log 2>&1
timeout 300 python bench.py --steps 30 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_r1_n1.json
cat gpurun_out/pytest_gpu.log; tail -2 gpurun_out/ncu_f.log | cut -c1-200; cut -c1-600 gpurun_out/bench_r1_n1.json
