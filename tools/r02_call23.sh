#!/bin/bash
# Round 2, call 23: weight gradients on their own stream: GPU suite, A/B against SG2IM_WGRAD_STREAM=0
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
LOG=gpurun_out/r02_call23.log
: > $LOG
timeout 1500 python -m pytest tests -q -m gpu -rf >> $LOG 2>&1
echo "exit $? (gpu suite)" >> $LOG
for i in 1 2; do
  timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02w_bench_wgstream_$i.json 2>> $LOG
  SG2IM_WGRAD_STREAM=0 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02w_bench_nowgstream_$i.json 2>> $LOG
done
timeout 300 python bench.py --no-cpu-baseline --no-graph > gpurun_out/r02w_bench_eager.json 2>> $LOG
grep -E "^exit|passed|failed" $LOG
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02w_bench_*.json')):
  try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d.get('parity', {}).get('rel_err', {}).get('image'))
  except Exception as e:
    print(f, 'ERR', e)
PY
tail -5 $LOG
