#!/bin/bash
# Round 2, call 28: ncu launch list of the final step (eager, one stream) and an `ncu --set full`
# capture of one small per-tap launch (graph-convolution Linear 448 x 384 -> 512).
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
LOG=gpurun_out/r02_call28.log
: > $LOG
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv \
  --log-file gpurun_out/r02_final4_launches.csv python bench.py --no-cpu-baseline --no-e2e --steps 2 --warmup 1 --no-graph \
  > gpurun_out/r02_final4_ncu_bench.log 2>&1
echo "exit $? (ncu launch list)" >> $LOG
cat > /tmp/one_linear.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from sg2im_b200 import ops
ops.set_conv_math('bf16x3')
dev = torch.device('cuda:0')
x = torch.randn(448, 1, 1, 384, device=dev)
w = torch.randn(512, 384, 1, 1, device=dev) * 0.05
b = torch.randn(512, device=dev)
wk = w.permute(2, 3, 1, 0).contiguous().permute(3, 2, 0, 1)
sh = ops.SplitShadows([wk]); sh.refresh()
for _ in range(8):
  y = ops.conv_tc_presplit(x, wk._split_fwd, 512, b, 1, 1, 0, 512, 1, 0.0)
torch.cuda.synchronize()
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 5 -c 1 \
  -f -o gpurun_out/r02_final4_linear_448_384_512 python /tmp/one_linear.py > gpurun_out/r02_final4_linear.log 2>&1
echo "exit $? (ncu linear)" >> $LOG
ncu -i gpurun_out/r02_final4_linear_448_384_512.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_raw_extract.py > gpurun_out/r02_final4_linear.txt
cat gpurun_out/r02_final4_linear.txt >> $LOG
grep -E "^exit" $LOG; cat gpurun_out/r02_final4_linear.txt | head -30; wc -l gpurun_out/r02_final4_launches.csv
