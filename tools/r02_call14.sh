#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
LOG=gpurun_out/r02_call14.log
: > $LOG
timeout 1500 python -m pytest tests -q -m gpu -rf >> $LOG 2>&1
echo "exit $? (gpu suite)" >> $LOG
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r02n_bench_bf16x3.json 2>> $LOG
timeout 300 python tools/aten_sites.py > gpurun_out/r02n_aten_sites.txt 2>> $LOG
echo "exit $? (aten sites)" >> $LOG
grep -E "^exit|passed|failed" $LOG
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02n_bench_bf16x3.json').read().strip().splitlines()[-1])
print('bench', d['value'], d['ms_per_step'], d['gpu_launches'])
PY
head -60 gpurun_out/r02n_aten_sites.txt
