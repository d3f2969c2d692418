#!/bin/bash
# Round 2, call 4: pre-split weight operands + 1/world folded into Adam on hardware; the whole GPU
# suite; bench per arithmetic; the other configurations of BASELINE.json.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
LOG=gpurun_out/r02_call4.log
: > $LOG
echo "== full gpu suite" >> $LOG
timeout 1500 python -m pytest tests -q -m gpu -rf -x >> $LOG 2>&1
echo "exit $? (gpu suite)" >> $LOG
for m in bf16x3 tf32 bf16; do
  echo "== bench --math $m" >> $LOG
  timeout 400 python bench.py --no-cpu-baseline --math $m > gpurun_out/r02d_bench_$m.json 2>> $LOG
  echo "exit $?" >> $LOG
done
echo "== conv shapes, events" >> $LOG
for w in fwd dgrad; do for s in big mid small n64; do
  timeout 120 python tools/prof_conv.py $w $s bf16x3 >> $LOG 2>&1
done; done
for wl in coco64 vg256 dense128; do
  echo "== bench --workload $wl (bf16x3)" >> $LOG
  timeout 400 python bench.py --no-cpu-baseline --workload $wl --steps 20 --warmup 5 > gpurun_out/r02d_bench_wl_$wl.json 2>> $LOG
  echo "exit $?" >> $LOG
done
echo "== bench --workload vg256 --math bf16 (BASELINE.json configs[3]: bf16 operands)" >> $LOG
timeout 400 python bench.py --no-cpu-baseline --workload vg256 --math bf16 --steps 20 --warmup 5 > gpurun_out/r02d_bench_wl_vg256_bf16.json 2>> $LOG
grep -E "^exit|passed|failed|TFLOP" $LOG
for f in gpurun_out/r02d_bench_*.json; do
  python - "$f" <<'PY'
import json, sys
try:
  d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
  print(sys.argv[1], d['value'], d['unit'], d['ms_per_step'], 'ms', d.get('roofline', {}).get('frac'))
except Exception as e:
  print(sys.argv[1], 'unreadable:', e)
PY
done
