#!/bin/bash
# Round 2, call 25: discriminators' real-image half beside the generator forward (deferred BatchNorm
# running statistics): GPU suite, A/B against SG2IM_EARLY_DREAL=0 on one box.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
LOG=gpurun_out/r02_call25.log
: > $LOG
timeout 1500 python -m pytest tests -q -m gpu -rf >> $LOG 2>&1
echo "exit $? (gpu suite)" >> $LOG
for i in 1 2; do
  timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02y_bench_early_$i.json 2>> $LOG
  SG2IM_EARLY_DREAL=0 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02y_bench_late_$i.json 2>> $LOG
done
grep -E "^exit|passed|failed" $LOG
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02y_bench_*.json')):
  try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d['value'], d['ms_per_step'], d['e2e']['value'])
  except Exception as e:
    print(f, 'ERR', e)
PY
tail -5 $LOG
