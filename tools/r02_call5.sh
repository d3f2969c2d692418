#!/bin/bash
# Round 2, call 5: 128-wide Cout tile of the halo kernel on hardware (A/B against the 64-wide tile),
# then compute-sanitizer memcheck over smoke() (one small G+D iteration per arithmetic).
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
LOG=gpurun_out/r02_call5.log
: > $LOG
echo "== bf16x3 kernels" >> $LOG
timeout 900 python -m pytest tests/test_gpu_bf16x3.py -q -m gpu -rf -x >> $LOG 2>&1
echo "exit $? (bf16x3 tests)" >> $LOG
echo "== bench bf16x3 (halo Cout tile 128 where Cout >= 128)" >> $LOG
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r02e_bench_bf16x3.json 2>> $LOG
echo "== bench bf16x3, SG2IM_HALO_BN=64" >> $LOG
SG2IM_HALO_BN=64 timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r02e_bench_bf16x3_halo64.json 2>> $LOG
echo "== bench bf16" >> $LOG
timeout 400 python bench.py --no-cpu-baseline --math bf16 > gpurun_out/r02e_bench_bf16.json 2>> $LOG
echo "== conv shapes, events" >> $LOG
for w in fwd dgrad; do for s in mid; do
  timeout 120 python tools/prof_conv.py $w $s bf16x3 >> $LOG 2>&1
  SG2IM_HALO_BN=64 timeout 120 python tools/prof_conv.py $w $s bf16x3 >> $LOG 2>&1
done; done
echo "== ncu: halo kernel, Cout tile 128 (672 -> 256 at 32x32)" >> $LOG
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:conv_tc_halo_kernel" -s 3 -c 1 \
  -f -o gpurun_out/r02_prof_halo128_fwd_mid python tools/prof_conv.py fwd mid bf16x3 > gpurun_out/r02_prof_halo128_fwd_mid.log 2>&1
ncu -i gpurun_out/r02_prof_halo128_fwd_mid.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_raw_extract.py >> $LOG
echo "== compute-sanitizer memcheck: smoke()" >> $LOG
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/r02_memcheck_smoke.txt \
  python -c "import __graft_entry__ as g; g.smoke()" >> $LOG 2>&1
echo "exit $? (memcheck smoke)" >> $LOG
tail -5 gpurun_out/r02_memcheck_smoke.txt >> $LOG
grep -E "^exit|passed|failed|TFLOP|ERROR SUMMARY" $LOG
for f in gpurun_out/r02e_bench_*.json; do
  python - "$f" <<'PY'
import json, sys
try:
  d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
  print(sys.argv[1], d['value'], d['unit'], d['ms_per_step'], 'ms', d.get('roofline', {}).get('frac'))
except Exception as e:
  print(sys.argv[1], 'unreadable:', e)
PY
done
