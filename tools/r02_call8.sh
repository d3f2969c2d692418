#!/bin/bash
# Round 2, call 8 (2 GPUs): config-count parity tests, overlapped generator all-reduce,
# then the data-parallel path: replica sync at construction, SUM all-reduce with 1/world in Adam.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
LOG=gpurun_out/r02_call8.log
: > $LOG
echo "== gpu suite" >> $LOG
CUDA_VISIBLE_DEVICES=0 timeout 1500 python -m pytest tests -q -m gpu -rf -x >> $LOG 2>&1
echo "exit $? (gpu suite)" >> $LOG
echo "== bench N=1" >> $LOG
CUDA_VISIBLE_DEVICES=0 timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r02h_bench_1gpu.json 2>> $LOG
echo "exit $?" >> $LOG
echo "== bench N=1, 40 signatures cycling (capture cooldown -> mostly eager)" >> $LOG
CUDA_VISIBLE_DEVICES=0 timeout 400 python bench.py --no-cpu-baseline --shape-jitter 40 --steps 40 > gpurun_out/r02h_bench_jitter40.json 2>> $LOG
echo "== bench N=1 eager (--no-graph)" >> $LOG
CUDA_VISIBLE_DEVICES=0 timeout 400 python bench.py --no-cpu-baseline --no-graph --steps 20 > gpurun_out/r02h_bench_eager.json 2>> $LOG
echo "== bench N=2 (torchrun, NCCL)" >> $LOG
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29527 \
  bench.py --gpus 2 --no-cpu-baseline > gpurun_out/r02h_bench_2gpu.json 2>> $LOG
echo "exit $? (2 gpus)" >> $LOG
echo "== 2-rank replica check (different seeds per rank -> identical after TrainStep)" >> $LOG
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29528 \
  tools/ddp_replica_check.py >> $LOG 2>&1
echo "exit $? (replica check)" >> $LOG
grep -E "^exit|passed|failed|replicas" $LOG
for f in gpurun_out/r02h_bench_*.json; do
  python - "$f" <<'PY'
import json, sys
try:
  d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
  print(sys.argv[1], d['value'], d['unit'], d['ms_per_step'], 'ms', d['config'].get('graphs_cached'), d['config'].get('graph_evictions'), d['config'].get('graph_replays'))
except Exception as e:
  print(sys.argv[1], 'unreadable:', e)
PY
done
