#!/bin/bash
# Round 2, call 6: full GPU suite after the ADVICE fixes, bench with per-shape rows measured in the
# step's real configuration (pre-split weights), measured parity field, variable-shape runs.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
LOG=gpurun_out/r02_call6.log
: > $LOG
echo "== full gpu suite" >> $LOG
timeout 1500 python -m pytest tests -q -m gpu -rf -x >> $LOG 2>&1
echo "exit $? (gpu suite)" >> $LOG
echo "== bench bf16x3 (with cpu baseline)" >> $LOG
timeout 600 python bench.py --shapes-out gpurun_out/r02f_shapes_bf16x3.json > gpurun_out/r02f_bench_bf16x3.json 2>> $LOG
echo "exit $?" >> $LOG
for m in tf32 bf16; do
  timeout 400 python bench.py --no-cpu-baseline --math $m > gpurun_out/r02f_bench_$m.json 2>> $LOG
done
echo "== variable batch shapes: 3 recurring signatures (replayed), 40 signatures (eager fallback)" >> $LOG
timeout 400 python bench.py --no-cpu-baseline --shape-jitter 3 > gpurun_out/r02f_bench_jitter3.json 2>> $LOG
timeout 400 python bench.py --no-cpu-baseline --shape-jitter 40 --steps 40 > gpurun_out/r02f_bench_jitter40.json 2>> $LOG
echo "== reference arm" >> $LOG
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02f_bench_reference.json 2>> $LOG
echo "== conv shapes, events, pre-split" >> $LOG
for w in fwd dgrad; do for s in big mid small n64; do
  timeout 120 python tools/prof_conv.py $w $s bf16x3 >> $LOG 2>&1
done; done
grep -E "^exit|passed|failed|TFLOP" $LOG
for f in gpurun_out/r02f_bench_*.json; do
  python - "$f" <<'PY'
import json, sys
try:
  d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
  print(sys.argv[1], d['value'], d['unit'], d['ms_per_step'], 'ms', (d.get('roofline') or {}).get('frac'), d.get('parity'))
except Exception as e:
  print(sys.argv[1], 'unreadable:', e)
PY
done
