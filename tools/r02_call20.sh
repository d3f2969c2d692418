#!/bin/bash
# Round 2, call 20: discriminator iteration on a second stream (fork / join inside the captured graph):
# GPU suite, then A/B against SG2IM_OVERLAP_DSTEP=0 on the same box.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
LOG=gpurun_out/r02_call20.log
: > $LOG
timeout 1500 python -m pytest tests -q -m gpu -rf >> $LOG 2>&1
echo "exit $? (gpu suite)" >> $LOG
for i in 1 2; do
  timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02t_bench_overlap_$i.json 2>> $LOG
  SG2IM_OVERLAP_DSTEP=0 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02t_bench_serial_$i.json 2>> $LOG
done
grep -E "^exit|passed|failed" $LOG
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02t_bench_*.json')):
  try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['last_losses'] if 'last_losses' in d else '')
  except Exception as e:
    print(f, 'ERR', e)
PY
tail -5 $LOG
