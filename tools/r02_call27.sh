#!/bin/bash
# Round 2, call 27: final state — smoke(), full GPU suite, the default bench line (with the CPU arm),
# the reference arm.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
LOG=gpurun_out/r02_call27.log
: > $LOG
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" >> $LOG 2>&1
echo "exit $? (smoke)" >> $LOG
timeout 1500 python -m pytest tests -q -m gpu -rf >> $LOG 2>&1
echo "exit $? (gpu suite)" >> $LOG
timeout 900 python bench.py > gpurun_out/r02_final_bench.json 2>> $LOG
echo "exit $? (bench)" >> $LOG
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_final_bench_reference.json 2>> $LOG
echo "exit $? (reference arm)" >> $LOG
grep -E "^exit|passed|failed|rel err" $LOG
python - <<'PY'
import json
for f in ('gpurun_out/r02_final_bench.json', 'gpurun_out/r02_final_bench_reference.json'):
  d = json.loads(open(f).read().strip().splitlines()[-1])
  print(f, d.get('value'), d.get('ms_per_step'), d.get('e2e'), d.get('cpu_baseline'), d.get('gpu_launches'))
  if 'roofline' in d:
    r = d['roofline']
    print({k: r[k] for k in ('achieved', 'peak', 'frac', 'traffic', 'share_of_step', 'frac_of_arithmetic_ceiling') if k in r})
PY
