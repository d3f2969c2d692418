#!/bin/bash
# Round 2, call 12 (8 GPUs): the scaling configuration the driver runs at round end — bench.py under
# torchrun at N = 8 and N = 4, with the overlapped generator all-reduce and without it.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
LOG=gpurun_out/r02_call12.log
: > $LOG
run() {   # name, nproc, env...
  local name=$1 n=$2; shift 2
  echo "== $name" >> $LOG
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
    --master-port $((29600 + RANDOM % 300)) bench.py --gpus $n --no-cpu-baseline > gpurun_out/r02l_bench_$name.json 2>> $LOG
  echo "exit $? ($name)" >> $LOG
}
run 8gpu 8 SG2IM_OVERLAP_ALLREDUCE=1
run 8gpu_nooverlap 8 SG2IM_OVERLAP_ALLREDUCE=0
run 4gpu 4 SG2IM_OVERLAP_ALLREDUCE=1
echo "== replica check, 8 ranks" >> $LOG
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 \
  tools/ddp_replica_check.py >> $LOG 2>&1
echo "exit $? (replica check)" >> $LOG
CUDA_VISIBLE_DEVICES=0 timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r02l_bench_1gpu.json 2>> $LOG
grep -E "^exit|replicas" $LOG
for f in gpurun_out/r02l_bench_*.json; do
  python - "$f" <<'PY'
import json, sys
try:
  d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
  print(sys.argv[1], d['n_gpus'], d['value'], d['unit'], d['ms_per_step'], 'ms', d['clocks'])
except Exception as e:
  print(sys.argv[1], 'unreadable:', e)
PY
done
