#!/bin/bash
# Round 2, call 24 (2 GPUs): the two-stream step under NCCL: bench N=1 / N=2, replica check.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
LOG=gpurun_out/r02_call24.log
: > $LOG
CUDA_VISIBLE_DEVICES=0 timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r02x_bench_1gpu.json 2>> $LOG
echo "exit $? (1 gpu)" >> $LOG
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29527 \
  bench.py --gpus 2 --no-cpu-baseline > gpurun_out/r02x_bench_2gpu.json 2>> $LOG
echo "exit $? (2 gpus)" >> $LOG
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29528 \
  tools/ddp_replica_check.py >> $LOG 2>&1
echo "exit $? (replica check)" >> $LOG
grep -E "^exit|replicas" $LOG
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02x_bench_*.json')):
  try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['n_gpus'])
  except Exception as e:
    print(f, 'ERR', e)
PY
tail -4 $LOG
