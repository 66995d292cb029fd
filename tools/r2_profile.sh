#!/bin/bash
# ncu evidence for profiles/ (round 2): launch list of a bench step, --set full of the tensor-core kernels, sanitizer subset
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-edit --no-full --no-config5"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/ncu_l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tapgemm_tc|decout_tc|conv1_tc' -s 22 -c 11 -o gpurun_out/r2_prof $B > gpurun_out/ncu_f.log 2>&1
tail -2 gpurun_out/ncu_f.log | cut -c1-300
# launch lists of the other two measured configurations: edit loop (batch 128), full IAN bf16 (batch 512)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_edit_b128.csv python tools/edit_once.py > /dev/null 2>&1
FULL_PREC=bf16 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_full_ian_bf16_b512.csv python tools/full_once.py > /dev/null 2>&1
# --set full of the full-IAN bf16 tensor-core kernels (second iteration): what bounds the Cout = 128 layers
[ -z "$SKIP_FULL_NCU" ] && FULL_PREC=bf16 FULL_IT=2 timeout 900 ncu --set full --clock-control none -k regex:'tapgemm_tc|head_tc|conv1_tc' -s 18 -c 18 -o gpurun_out/r2_prof_full python tools/full_once.py > gpurun_out/ncu_full.log 2>&1
# sanitizer: small batches through every kernel family (forward, brush, edit, full IAN, v1)
cat > /tmp/san.py <<'PY'
import importlib, sys, numpy as np
sys.path.insert(0, ".")
from oracle import weights as ow
pkg = importlib.import_module("neural-photo-editor_b200")
rng = np.random.default_rng(0)
m = pkg.IAN("IAN_simple.py", True, weights=ow.make_simple_weights(0))
for n in (1, 5):
    x = rng.uniform(-1, 1, (n, 3, 64, 64)).astype(np.float32)
    xh, z = m.reconstruct(x, return_z=True)
    boxes = np.tile(np.array([[8, 8, 30, 28]], np.int32), (n, 1)); rgb = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    m.grad(z, boxes, rgb); m.edit_steps(z, boxes, rgb, n_steps=2)
x = rng.uniform(-1, 1, (160, 3, 64, 64)).astype(np.float32)      # large enough for the pair kernel and stream-K
m.reconstruct(x)
x = rng.uniform(-1, 1, (256, 3, 64, 64)).astype(np.float32)      # enc_fc1 on the pair kernel's split-K
m.reconstruct(x)
z, boxes, rgb = ow.config4_inputs(128)                            # the batch-128 edit step: seed kernel, pair stream-K backward, cooperative finalize
m.edit_steps(z, boxes, rgb, n_steps=2)
m.close()
f = pkg.IAN("IAN.py", True, weights=ow.make_full_weights(0))
x = rng.uniform(-1, 1, (3, 3, 64, 64)).astype(np.float32)
f.reconstruct(x); f.set_precision("bf16"); f.reconstruct(x); f.close()
print("sanitizer workload done")
PY
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool" > gpurun_out/r2_sanitizer_$tool.log
  IAN_GRAPHS=0 timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san.py 2>&1 | grep -v "^Loading\|^Shuffling" | tail -40 >> gpurun_out/r2_sanitizer_$tool.log
  tail -3 gpurun_out/r2_sanitizer_$tool.log
done
