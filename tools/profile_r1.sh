#!/bin/bash
# ncu evidence for profiles/: launch list of a bench step + full capture of the tensor-core kernels + bench line
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -3) > gpurun_out/pytest_gpu.log
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1c.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-edit --no-full > gpurun_out/ncu_b.log 2>&1
timeout 800 ncu --set full --clock-control none --import-source on -k regex:'tapgemm_tc|decout_tc|conv1_tc' -s 22 -c 11 -o gpurun_out/prof_r1c \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-edit --no-full > gpurun_out/ncu_f.log 2>&1
timeout 400 python bench.py --steps 30 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_r1_n1.json
cat gpurun_out/pytest_gpu.log; tail -1 gpurun_out/ncu_f.log | cut -c1-200; cut -c1-300 gpurun_out/bench_r1_n1.json
