#!/bin/bash
# HISTORIC: this is the script that ran as round 2's first GPU call (profiles/r02_call1_*), against
# the tree of commit 95c8e88.  The variants it A/B-tested that lost or failed (cluster multicast, CTA
# pair, two-image halo, pack-both, tf32x3) have since been deleted, so it no longer runs as is.
# First GPU call of round 2 (run under gpurun from the repo root):
#   gpurun --timeout 3000 -- 'bash tools/r02_first_call.sh'      (about 35-40 GPU-minutes)
# 1. hardware validation of everything written after round 1's GPU budget ran out
# 2. A/B of the opt-in kernels on the default bench
# 3. tile sweep of the conv kernel over the step's shapes
# Every step is bounded by its own timeout and writes under gpurun_out/.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
LOG=gpurun_out/r02_first.log
: > $LOG
# each group of not-yet-run kernels in its own process and under its own timeout: a kernel that
# hangs or faults on hardware takes down only its group, and its A/B bench below is skipped
group() {   # name, timeout, -k expression
  echo "== tests: $1" >> $LOG
  SG2IM_RUN_UNVERIFIED=1 timeout "$2" python -m pytest tests/test_gpu_next_rows.py -q -m gpu -rf -k "$3" >> $LOG 2>&1
  local rc=$?
  echo "exit $rc ($1)" >> $LOG
  eval "RC_$1=$rc"
}
group rows 600 "deprocess or check_model or staged_batch or align_corners or eval_bn or coco_relations"
group simt 600 "bn_backward_v2 or layout_backward_v2 or layout_forward_v2 or scale_act_forward_v2 or colsum or pack_both or flat_adam_kernel or fused_activation"
group kcc 300 "conv_from_weight_gradient_layout"
group wgradmc 300 "wgrad_cluster_multicast"
group convmc 300 "conv_cluster_multicast"
group halosmall 300 "small_image_halo"
group halopair 300 "cta_pair_halo"
group steps 600 "train_step_with"
group factory 300 "pool2d or instance_norm or build_cnn_residual"
group x3 600 "split_tf32 or tf32x3"
group boxes 300 "wrt_boxes or predicted_boxes"
grep -E "^exit|passed|failed" $LOG
echo "== tcgen05 issue-rate probe" >> $LOG
(timeout 120 nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -I sg2im_b200/csrc \
   -o /tmp/umma_rate_probe tools/umma_rate_probe.cu -lcuda && timeout 60 /tmp/umma_rate_probe) \
   > gpurun_out/r02_umma_rate.txt 2>&1
echo "exit $?" >> $LOG
echo "== CTA-pair (cta_group::2) semantics + rate probe" >> $LOG
(timeout 120 nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -I sg2im_b200/csrc \
   -o /tmp/umma_2cta_probe tools/umma_2cta_probe.cu -lcuda && timeout 60 /tmp/umma_2cta_probe) \
   > gpurun_out/r02_umma_2cta.txt 2>&1
echo "exit $?" >> $LOG
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
[ "$RC_wgradmc" = 0 ] && SG2IM_WGRAD_MC=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_wgradmc.json 2>> gpurun_out/r02_first.log
echo "== bench forward cluster multicast" >> gpurun_out/r02_first.log
[ "$RC_convmc" = 0 ] && SG2IM_CONV_MC=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_convmc.json 2>> gpurun_out/r02_first.log
echo "== bench small-image halo kernel (8-row maps)" >> gpurun_out/r02_first.log
[ "$RC_halosmall" = 0 ] && SG2IM_HALO_SMALL=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_halosmall.json 2>> gpurun_out/r02_first.log
echo "== bench CTA-pair halo kernel (cta_group::2)" >> gpurun_out/r02_first.log
[ "$RC_halopair" = 0 ] && SG2IM_HALO_PAIR=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_halopair.json 2>> gpurun_out/r02_first.log
echo "== bench error-compensated tensor-core mode (tf32x3: the 1e-3 parity bar on tcgen05)" >> gpurun_out/r02_first.log
[ "$RC_x3" = 0 ] && timeout 300 python bench.py --no-cpu-baseline --math tf32x3 > gpurun_out/r02_bench_tf32x3.json 2>> gpurun_out/r02_first.log
echo "== bench pack-both" >> gpurun_out/r02_first.log
SG2IM_PACK_BOTH=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_packboth.json 2>> gpurun_out/r02_first.log
echo "== bench weights in the gradient layout (no pack / unpack)" >> gpurun_out/r02_first.log
[ "$RC_kcc" = 0 ] && timeout 300 python bench.py --no-cpu-baseline --weights kcc > gpurun_out/r02_bench_kcc.json 2>> gpurun_out/r02_first.log
[ "$RC_kcc" = 0 ] && timeout 300 python bench.py --no-cpu-baseline --weights kcc --adam flat > gpurun_out/r02_bench_kcc_flatadam.json 2>> gpurun_out/r02_first.log
echo "== bench fused activation backward + bias gradient (kcc, flat Adam)" >> gpurun_out/r02_first.log
[ "$RC_kcc" = 0 ] && SG2IM_ACTBWD_FUSED=1 timeout 300 python bench.py --no-cpu-baseline --weights kcc --adam flat > gpurun_out/r02_bench_kcc_actbwd.json 2>> gpurun_out/r02_first.log
echo "== bench all opt-ins" >> gpurun_out/r02_first.log
ALL="SG2IM_PACK_BOTH=1"
[ "$RC_simt" = 0 ] && ALL="$ALL SG2IM_BNBWD_V2=1 SG2IM_BNFWD_V2=1 SG2IM_LAYOUT_V2=1 SG2IM_COLSUM_V2=1 SG2IM_ACTBWD_FUSED=1"
[ "$RC_wgradmc" = 0 ] && ALL="$ALL SG2IM_WGRAD_MC=1"
[ "$RC_halosmall" = 0 ] && ALL="$ALL SG2IM_HALO_SMALL=1"
[ "$RC_halopair" = 0 ] && ALL="$ALL SG2IM_HALO_PAIR=1"
ARGS="--adam flat"
[ "$RC_kcc" = 0 ] && [ "$RC_steps" = 0 ] && ARGS="$ARGS --weights kcc"
echo "all: $ALL $ARGS" >> gpurun_out/r02_first.log
env $ALL timeout 300 python bench.py --no-cpu-baseline $ARGS > gpurun_out/r02_bench_all.json 2>> gpurun_out/r02_first.log
echo "== the other configurations of BASELINE.json (C2 COCO-64, C4 VG-256 six-stage CRN, C5 dense graphs)" >> gpurun_out/r02_first.log
for wl in coco64 vg256 dense128; do
  timeout 300 python bench.py --no-cpu-baseline --workload $wl --steps 20 --warmup 5 > gpurun_out/r02_bench_wl_$wl.json 2>> gpurun_out/r02_first.log
done
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
