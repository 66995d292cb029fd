#!/bin/bash
# N-GPU bench (torchrun, NCCL) + the reference arm, as the driver launches them
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 30 --warmup 5 2>&1 | grep -E '^\{|Error|error' | tail -3 > gpurun_out/bench_n$N.json
python - "$N" <<'PY'
import json, sys
n = sys.argv[1]
for line in open('gpurun_out/bench_n%s.json' % n):
    try:
        d = json.loads(line)
    except Exception:
        print(line[:300]); continue
    print("N=%s ms/step %.3f  img/s %.0f  e2e %.0f  launches %d" % (d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["value"], d["gpu_launches"]), d["clocks"])
PY
# the reference arm holds every GPU of the box idle while the CPU works: opt in with REF_ARM=1
[ "${REF_ARM:-0}" = 1 ] && timeout 300 python bench.py --impl reference --steps 4 --warmup 1 2>&1 | tail -1 | cut -c1-700
exit 0
