#!/bin/bash
# Round 2, call 9: full GPU suite, final ncu captures (pre-split halo kernels), launch list of the
# final step, per-tap N-tile experiment on 8x8 maps, racecheck over the SIMT kernels' unit tests.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
LOG=gpurun_out/r02_call9.log
: > $LOG
echo "== full gpu suite" >> $LOG
timeout 1500 python -m pytest tests -q -m gpu -rf >> $LOG 2>&1
echo "exit $? (gpu suite)" >> $LOG
echo "== bench bf16x3" >> $LOG
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r02i_bench_bf16x3.json 2>> $LOG
echo "== per-tap N tile on 8x8 maps (1024 -> 1024): heuristic (64) vs 128" >> $LOG
timeout 120 python tools/prof_conv.py fwd small bf16x3 >> $LOG 2>&1
SG2IM_TC_BN=128 timeout 120 python tools/prof_conv.py fwd small bf16x3 >> $LOG 2>&1
SG2IM_TC_BN=256 timeout 120 python tools/prof_conv.py fwd small bf16x3 >> $LOG 2>&1
cap() {   # name, what, shape, kernel regex
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$4" -s 6 -c 1 \
    -f -o "gpurun_out/r02_final_$1" python tools/prof_conv.py "$2" "$3" bf16x3 > "gpurun_out/r02_final_$1.log" 2>&1
  echo "== $1" >> gpurun_out/r02_final_conv_kernels.txt
  ncu -i "gpurun_out/r02_final_$1.ncu-rep" --page raw --csv 2>/dev/null | python tools/ncu_raw_extract.py >> gpurun_out/r02_final_conv_kernels.txt
}
: > gpurun_out/r02_final_conv_kernels.txt
cap halo64_fwd_stage4_conv1 fwd big conv_tc_halo_kernel
cap halo128_dgrad_stage4_conv1 dgrad big conv_tc_halo_kernel
cap halo128_fwd_672_256 fwd mid conv_tc_halo_kernel
cap pertap_fwd_1024_8x8 fwd small conv_tc_kernel
cap wgrad_stage4_conv1 wgrad big conv_wgrad_tc_kernel
cat gpurun_out/r02_final_conv_kernels.txt >> $LOG
echo "== launch list of the final step (eager, 2 steps)" >> $LOG
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
  --log-file gpurun_out/r02i_launches.csv python bench.py --no-cpu-baseline --no-e2e --steps 2 --warmup 1 --no-graph \
  > gpurun_out/r02i_ncu_bench.log 2>&1
echo "exit $? (ncu launch list)" >> $LOG
echo "== compute-sanitizer racecheck: graph pooling / layout / crop / BN unit tests" >> $LOG
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 --log-file gpurun_out/r02_racecheck_ops.txt \
  python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "graph_pool_empty or gconv_layer or layout_golden or crop_golden or upsample_then_bn" >> $LOG 2>&1
echo "exit $? (racecheck ops)" >> $LOG
tail -3 gpurun_out/r02_racecheck_ops.txt >> $LOG
grep -E "^exit|passed|failed|TFLOP|RACECHECK SUMMARY" $LOG
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02i_bench_bf16x3.json').read().strip().splitlines()[-1])
print('bench', d['value'], d['ms_per_step'])
PY
