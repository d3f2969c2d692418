#!/bin/bash
# Round 2, call 2: first hardware run of the bf16x3 / bf16 tensor-core arithmetic (in-kernel
# operand split), then the whole GPU suite and the bench line in each arithmetic.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
LOG=gpurun_out/r02_call2.log
: > $LOG
echo "== bf16x3 kernels" >> $LOG
timeout 900 python -m pytest tests/test_gpu_bf16x3.py -q -m gpu -rf -s >> $LOG 2>&1
echo "exit $? (bf16x3 tests)" >> $LOG
echo "== smoke" >> $LOG
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" >> $LOG 2>&1
echo "exit $? (smoke)" >> $LOG
echo "== full gpu suite" >> $LOG
timeout 1500 python -m pytest tests -q -m gpu -rf >> $LOG 2>&1
echo "exit $? (gpu suite)" >> $LOG
for m in bf16x3 tf32 bf16; do
  echo "== bench --math $m" >> $LOG
  timeout 400 python bench.py --no-cpu-baseline --math $m > gpurun_out/r02b_bench_$m.json 2>> $LOG
  echo "exit $?" >> $LOG
done
echo "== bench bf16x3 packed weights / torch adam (the reference's layout)" >> $LOG
timeout 400 python bench.py --no-cpu-baseline --weights oihw --adam torch > gpurun_out/r02b_bench_bf16x3_oihw.json 2>> $LOG
echo "== launch list of the default step" >> $LOG
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
  --log-file gpurun_out/r02b_launches.csv python bench.py --no-cpu-baseline --no-e2e --steps 2 --warmup 1 --no-graph \
  > gpurun_out/r02b_ncu_bench.log 2>&1
echo "exit $? (ncu)" >> $LOG
grep -E "^exit|passed|failed" $LOG
for f in gpurun_out/r02b_bench_*.json; do
  python - "$f" <<'PY'
import json, sys
try:
  d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
  print(sys.argv[1], d['value'], d['unit'], d['ms_per_step'], 'ms', d.get('roofline', {}).get('frac'))
except Exception as e:
  print(sys.argv[1], 'unreadable:', e)
PY
done
