#!/bin/bash
# Round 2, call 10: explicit LDS/STS in the converters, relaxed mbarrier waits, wave-aware N tile of
# the per-tap kernel: full GPU suite, bench, per-shape events, ncu of the 8x8 per-tap and wgrad kernels.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
LOG=gpurun_out/r02_call10.log
: > $LOG
echo "== full gpu suite" >> $LOG
timeout 1500 python -m pytest tests -q -m gpu -rf >> $LOG 2>&1
echo "exit $? (gpu suite)" >> $LOG
for m in bf16x3 tf32 bf16; do
  timeout 400 python bench.py --no-cpu-baseline --math $m --shapes-out gpurun_out/r02j_shapes_$m.json > gpurun_out/r02j_bench_$m.json 2>> $LOG
done
echo "== conv shapes, events (pre-split weights)" >> $LOG
for w in fwd dgrad wgrad; do for s in big mid small n64; do
  timeout 120 python tools/prof_conv.py $w $s bf16x3 >> $LOG 2>&1
done; done
cap() {   # name, what, shape, kernel regex
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$4" -s 6 -c 1 \
    -f -o "gpurun_out/r02_final2_$1" python tools/prof_conv.py "$2" "$3" bf16x3 > "gpurun_out/r02_final2_$1.log" 2>&1
  echo "== $1" >> gpurun_out/r02_final2_conv_kernels.txt
  ncu -i "gpurun_out/r02_final2_$1.ncu-rep" --page raw --csv 2>/dev/null | python tools/ncu_raw_extract.py >> gpurun_out/r02_final2_conv_kernels.txt
}
: > gpurun_out/r02_final2_conv_kernels.txt
cap halo64_fwd_stage4_conv1 fwd big conv_tc_halo_kernel
cap pertap_fwd_1024_8x8 fwd small conv_tc_kernel
cap wgrad_stage4_conv1 wgrad big conv_wgrad_tc_kernel
cat gpurun_out/r02_final2_conv_kernels.txt >> $LOG
grep -E "^exit|passed|failed|TFLOP" $LOG
for f in gpurun_out/r02j_bench_*.json; do
  python - "$f" <<'PY'
import json, sys
try:
  d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
  print(sys.argv[1], d['value'], d['unit'], d['ms_per_step'], 'ms', (d.get('roofline') or {}).get('frac'))
except Exception as e:
  print(sys.argv[1], 'unreadable:', e)
PY
done
