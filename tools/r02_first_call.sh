#!/bin/bash
# First GPU call of round 2 (run under gpurun from the repo root):
#   gpurun --timeout 2400 -- 'bash tools/r02_first_call.sh'      (about 25 GPU-minutes)
# 1. hardware validation of everything written after round 1's GPU budget ran out
# 2. A/B of the opt-in kernels on the default bench
# 3. tile sweep of the conv kernel over the step's shapes
# Every step is bounded by its own timeout and writes under gpurun_out/.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== unverified GPU tests" | tee gpurun_out/r02_first.log
SG2IM_RUN_UNVERIFIED=1 timeout 600 python -m pytest tests/test_gpu_next_rows.py -q -m gpu -rf \
    >> gpurun_out/r02_first.log 2>&1
echo "exit $?" >> gpurun_out/r02_first.log
echo "== bench default" >> gpurun_out/r02_first.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_default.json 2>> gpurun_out/r02_first.log
echo "== bench BN backward v2" >> gpurun_out/r02_first.log
SG2IM_BNBWD_V2=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_bnv2.json 2>> gpurun_out/r02_first.log
echo "== bench BN forward v2" >> gpurun_out/r02_first.log
SG2IM_BNFWD_V2=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_bnfwdv2.json 2>> gpurun_out/r02_first.log
echo "== bench layout v2 (forward + backward)" >> gpurun_out/r02_first.log
SG2IM_LAYOUT_V2=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_layoutv2.json 2>> gpurun_out/r02_first.log
echo "== bench flat Adam" >> gpurun_out/r02_first.log
timeout 300 python bench.py --no-cpu-baseline --adam flat > gpurun_out/r02_bench_flatadam.json 2>> gpurun_out/r02_first.log
echo "== bench wgrad cluster multicast" >> gpurun_out/r02_first.log
SG2IM_WGRAD_MC=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_wgradmc.json 2>> gpurun_out/r02_first.log
echo "== bench forward cluster multicast" >> gpurun_out/r02_first.log
SG2IM_CONV_MC=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_convmc.json 2>> gpurun_out/r02_first.log
echo "== bench small-image halo kernel (8-row maps)" >> gpurun_out/r02_first.log
SG2IM_HALO_SMALL=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_halosmall.json 2>> gpurun_out/r02_first.log
echo "== bench pack-both" >> gpurun_out/r02_first.log
SG2IM_PACK_BOTH=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_packboth.json 2>> gpurun_out/r02_first.log
echo "== bench weights in the gradient layout (no pack / unpack)" >> gpurun_out/r02_first.log
timeout 300 python bench.py --no-cpu-baseline --weights kcc > gpurun_out/r02_bench_kcc.json 2>> gpurun_out/r02_first.log
timeout 300 python bench.py --no-cpu-baseline --weights kcc --adam flat > gpurun_out/r02_bench_kcc_flatadam.json 2>> gpurun_out/r02_first.log
echo "== bench fused activation backward + bias gradient (kcc, flat Adam)" >> gpurun_out/r02_first.log
SG2IM_ACTBWD_FUSED=1 timeout 300 python bench.py --no-cpu-baseline --weights kcc --adam flat > gpurun_out/r02_bench_kcc_actbwd.json 2>> gpurun_out/r02_first.log
echo "== bench all opt-ins" >> gpurun_out/r02_first.log
SG2IM_BNBWD_V2=1 SG2IM_BNFWD_V2=1 SG2IM_LAYOUT_V2=1 SG2IM_PACK_BOTH=1 SG2IM_COLSUM_V2=1 SG2IM_ACTBWD_FUSED=1 SG2IM_WGRAD_MC=1 timeout 300 python bench.py --no-cpu-baseline --adam flat --weights kcc > gpurun_out/r02_bench_all.json 2>> gpurun_out/r02_first.log
echo "== conv tile sweep" >> gpurun_out/r02_first.log
timeout 300 python tools/sweep_conv.py --out gpurun_out/r02_sweep_conv.json >> gpurun_out/r02_first.log 2>&1
echo "== kernel table (BN v2)" >> gpurun_out/r02_first.log
SG2IM_BNBWD_V2=1 timeout 300 python tools/kernel_table.py > gpurun_out/r02_kernel_table_bnv2.txt 2>> gpurun_out/r02_first.log
tail -5 gpurun_out/r02_first.log
for f in gpurun_out/r02_bench_*.json; do
  python - "$f" <<'PY'
import json, sys
try:
  d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
  print(sys.argv[1], d['value'], d['unit'], d['ms_per_step'], 'ms')
except Exception as e:
  print(sys.argv[1], 'unreadable:', e)
PY
done
