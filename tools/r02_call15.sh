#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
LOG=gpurun_out/r02_call15.log
: > $LOG
timeout 900 python -m pytest tests/test_gpu_bf16x3.py tests/test_gpu_next_rows.py tests/test_gpu_configs.py -q -m gpu -rf >> $LOG 2>&1
echo "exit $? (gpu tests)" >> $LOG
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r02o_bench_bf16x3.json 2>> $LOG
timeout 300 python tools/graph_gaps.py > gpurun_out/r02o_graph_gaps.txt 2>> $LOG
echo "exit $? (graph gaps)" >> $LOG
grep -E "^exit|passed|failed" $LOG
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02o_bench_bf16x3.json').read().strip().splitlines()[-1])
print('bench', d['value'], d['ms_per_step'], d['gpu_launches'])
for r in d['roofline']['by_shape'][:12]:
  print(r)
PY
cat gpurun_out/r02o_graph_gaps.txt
tail -5 $LOG
