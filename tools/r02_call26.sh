#!/bin/bash
# Round 2, call 26 (8 GPUs): the multi-stream step at the scaling configuration the driver runs at
# round end: bench.py under torchrun at N = 8 and N = 1, replica check.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
LOG=gpurun_out/r02_call26.log
: > $LOG
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
  --master-port 29611 bench.py --gpus 8 --no-cpu-baseline > gpurun_out/r02z_bench_8gpu.json 2>> $LOG
echo "exit $? (8 gpus)" >> $LOG
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 \
  tools/ddp_replica_check.py >> $LOG 2>&1
echo "exit $? (replica check)" >> $LOG
CUDA_VISIBLE_DEVICES=0 timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r02z_bench_1gpu.json 2>> $LOG
echo "exit $? (1 gpu)" >> $LOG
grep -E "^exit|replicas" $LOG
for f in gpurun_out/r02z_bench_*.json; do
  python - "$f" <<'PY'
import json, sys
try:
  d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
  print(sys.argv[1], d['n_gpus'], d['value'], d['unit'], d['ms_per_step'], 'ms', d['e2e']['value'], d['clocks'])
except Exception as e:
  print(sys.argv[1], 'unreadable:', e)
PY
done
tail -3 $LOG
