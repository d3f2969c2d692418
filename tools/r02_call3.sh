#!/bin/bash
# Round 2, call 3: converter-warp restructure on hardware, bench per arithmetic, ncu --set full
# captures of the tensor-core kernels (bf16x3) and of the HBM-bound kernels.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
LOG=gpurun_out/r02_call3.log
: > $LOG
echo "== bf16x3 kernels" >> $LOG
timeout 900 python -m pytest tests/test_gpu_bf16x3.py -q -m gpu -rf -x >> $LOG 2>&1
echo "exit $? (bf16x3 tests)" >> $LOG
for m in bf16x3 tf32 bf16; do
  echo "== bench --math $m" >> $LOG
  timeout 400 python bench.py --no-cpu-baseline --math $m > gpurun_out/r02c_bench_$m.json 2>> $LOG
  echo "exit $?" >> $LOG
done
echo "== conv shapes, events" >> $LOG
for w in fwd dgrad wgrad; do for s in big mid small n64; do
  timeout 120 python tools/prof_conv.py $w $s bf16x3 >> $LOG 2>&1
done; done
cap() {   # name, what, shape, kernel regex
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$4" -s 3 -c 1 \
    -f -o "gpurun_out/r02_prof_$1" python tools/prof_conv.py "$2" "$3" bf16x3 > "gpurun_out/r02_prof_$1.log" 2>&1
  echo "== $1" >> gpurun_out/r02_conv_kernels.txt
  ncu -i "gpurun_out/r02_prof_$1.ncu-rep" --page raw --csv 2>/dev/null | python tools/ncu_raw_extract.py >> gpurun_out/r02_conv_kernels.txt
}
: > gpurun_out/r02_conv_kernels.txt
cap halo_fwd_big fwd big conv_tc_halo_kernel
cap halo_dgrad_big dgrad big conv_tc_halo_kernel
cap wgrad_big wgrad big conv_wgrad_tc_kernel
cap pertap_fwd_small fwd small conv_tc_kernel
cap halo_fwd_n64 fwd n64 conv_tc_halo_kernel
cat gpurun_out/r02_conv_kernels.txt >> $LOG
echo "== HBM-bound kernels (ncu --set full + events)" >> $LOG
timeout 1500 bash tools/r02_ncu_hbm.sh >> $LOG 2>&1
grep -E "^exit|passed|failed" $LOG
for f in gpurun_out/r02c_bench_*.json; do
  python - "$f" <<'PY'
import json, sys
try:
  d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
  print(sys.argv[1], d['value'], d['unit'], d['ms_per_step'], 'ms', d.get('roofline', {}).get('frac'))
except Exception as e:
  print(sys.argv[1], 'unreadable:', e)
PY
done
