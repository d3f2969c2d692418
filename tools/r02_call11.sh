#!/bin/bash
# Round 2, call 11: GPU suite (skinny-rows kernel), bench, weight-gradient Cout-tile experiment.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
LOG=gpurun_out/r02_call11.log
: > $LOG
timeout 1500 python -m pytest tests -q -m gpu -rf >> $LOG 2>&1
echo "exit $? (gpu suite)" >> $LOG
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r02k_bench_bf16x3.json 2>> $LOG
echo "== weight gradient, Cout tile: heuristic vs pinned" >> $LOG
for s in mid small; do
  timeout 120 python tools/prof_conv.py wgrad $s bf16x3 >> $LOG 2>&1
  SG2IM_WG_BN=128 timeout 120 python tools/prof_conv.py wgrad $s bf16x3 >> $LOG 2>&1
  SG2IM_WG_BN=64 timeout 120 python tools/prof_conv.py wgrad $s bf16x3 >> $LOG 2>&1
done
echo "== bench with SG2IM_WG_BN=128" >> $LOG
SG2IM_WG_BN=128 timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r02k_bench_bf16x3_wg128.json 2>> $LOG
grep -E "^exit|passed|failed|TFLOP" $LOG
for f in gpurun_out/r02k_bench_*.json; do
  python - "$f" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], d['value'], d['ms_per_step'], [ (k, round(v['ms_per_step'],3)) for k,v in d['roofline']['by_kernel'].items()])
PY
done
