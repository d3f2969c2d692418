#!/bin/bash
# A/B: maximum shared-memory carve-out on every kernel (SG2IM_CARVEOUT=1) vs the default
set -u
mkdir -p gpurun_out
for i in 1 2; do
  timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02s_bench_default_$i.json 2>> gpurun_out/r02_call19.log
  SG2IM_CARVEOUT=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02s_bench_carveout_$i.json 2>> gpurun_out/r02_call19.log
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02s_bench_*.json')):
  d = json.loads(open(f).read().strip().splitlines()[-1])
  print(f, d['value'], d['ms_per_step'])
PY
tail -3 gpurun_out/r02_call19.log
