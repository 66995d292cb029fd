#!/bin/bash
# multi-GPU check (run under gpurun --gpus N): 2-process gather test, then bench at N with every gather mode
N=${1:-2}
mkdir -p gpurun_out/rec
export IAN_TEST_RECORD=$PWD/gpurun_out/rec
nvidia-smi -L | head -8 > gpurun_out/multi_smi.txt
(timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -15) > gpurun_out/multi_tests.log
cat gpurun_out/multi_tests.log | cut -c1-600
for mode in ${MODES:-p2p_async p2p nccl}; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 30 --warmup 5 --gather $mode > gpurun_out/bench_n${N}_$mode.json 2> gpurun_out/bench_n${N}_$mode.err
  tail -1 gpurun_out/bench_n${N}_$mode.err | cut -c1-300
done
if [ -n "$PUSH_KERNEL" ]; then
  IAN_PUSH=kernel timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 30 --warmup 5 --gather p2p_async --no-config5 > gpurun_out/bench_n${N}_p2p_async_kernel.json 2> /dev/null
fi
python - <<PY
import json
for mode in ("p2p_async", "p2p_async_kernel", "p2p", "nccl"):
    try:
        d = json.loads([l for l in open("gpurun_out/bench_n${N}_%s.json" % mode) if l.startswith("{")][-1])
        c5 = d.get("config5") or {}
        print(mode, "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "check", d["config"]["gather_check_max_abs_vs_nccl"],
              "| config5", round(c5.get("value", 0)), c5.get("gather"), c5.get("gather_check_max_abs_vs_nccl"), "| e2e", round(d["e2e"]["value"]),
              "| enc_conv2", d["roofline"]["layer_ms"]["enc_conv2"], "dec_out", d["roofline"]["edge_kernel_ms"]["dec_out"])
    except Exception as e:
        print(mode, "unreadable", e)
PY
