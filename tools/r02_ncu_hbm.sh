#!/bin/bash
# ncu --set full captures of the HBM-bound kernels at their benchmark sizes (north star: "graph
# scatter and bilinear warp ... evidenced by committed ncu captures reporting achieved HBM GB/s").
#   gpurun --timeout 1200 -- 'bash tools/r02_ncu_hbm.sh'            (one GPU, ~6-8 GPU-minutes)
# Per kernel: one .ncu-rep (bring back, read here with `ncu -i ... --page source --csv`) and a
# raw-page extract of the numbers that go into profiles/ (duration, DRAM bytes, DRAM throughput).
# The events-based GB/s of the same launches (no profiler attached) is printed first.
set -u
mkdir -p gpurun_out
OUT=gpurun_out/r02_hbm_kernels.txt
: > $OUT
echo "== CUDA-event timing, no profiler (tools/prof_hbm.py all)" >> $OUT
timeout 300 python tools/prof_hbm.py all >> $OUT 2>&1
# the first-generation kernels for comparison (the v2 kernels are the defaults now)
echo "== same, first-generation kernels (SG2IM_LAYOUT_V2=0 SG2IM_BNBWD_V2=0 SG2IM_BNFWD_V2=0)" >> $OUT
SG2IM_LAYOUT_V2=0 SG2IM_BNBWD_V2=0 SG2IM_BNFWD_V2=0 timeout 300 python tools/prof_hbm.py all >> $OUT 2>&1
cap() {   # name, workload, kernel regex, launches to skip (warm-up + earlier repetitions)
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$3" -s "$4" -c 1 \
    -f -o "gpurun_out/r02_prof_$1" python tools/prof_hbm.py "$2" > "gpurun_out/r02_prof_$1.log" 2>&1
  echo "== $1 (kernel regex $3)" >> $OUT
  ncu -i "gpurun_out/r02_prof_$1.ncu-rep" --page raw --csv 2>/dev/null \
    | python tools/ncu_raw_extract.py >> $OUT
}
# 6 repetitions of each workload (1 warm-up + 5): skip the first 3 launches of the kernel
cap layout_fwd layout layout_fwd 3
cap layout_bwd layout layout_bwd 3
cap triple_gather graph triple_gather 6
cap segment_sum graph segment_sum 6
cap crop_fwd crop crop_fwd 3
cap crop_bwd crop crop_bwd 3
cap scale_act_fwd bn scale_act_fwd 3
cap scale_act_bwd_apply bn "bwd_apply" 3
cap bn_bwd_reduce bn "bwd_reduce" 3
tail -40 $OUT
