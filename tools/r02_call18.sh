#!/bin/bash
# Round 2, call 18: full GPU suite on the final kernels, bench in the three arithmetics and the other
# workloads, BN-kernel GB/s after the 4-CTA reduce, ncu captures of the stacked weight gradient.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
LOG=gpurun_out/r02_call18.log
: > $LOG
echo "== full gpu suite" >> $LOG
timeout 1500 python -m pytest tests -q -m gpu -rf >> $LOG 2>&1
echo "exit $? (gpu suite)" >> $LOG
for m in bf16x3 tf32 bf16; do
  timeout 400 python bench.py --no-cpu-baseline --math $m > gpurun_out/r02r_bench_$m.json 2>> $LOG
done
for w in coco64 vg256 dense128; do
  timeout 400 python bench.py --no-cpu-baseline --workload $w > gpurun_out/r02r_bench_wl_$w.json 2>> $LOG
done
echo "== HBM kernels (events)" >> $LOG
timeout 200 python tools/prof_hbm.py bn >> $LOG 2>&1
cap() {   # name, what, shape, kernel regex
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$4" -s 6 -c 1 \
    -f -o "gpurun_out/r02_final3_$1" python tools/prof_conv.py "$2" "$3" bf16x3 > "gpurun_out/r02_final3_$1.log" 2>&1
  echo "== $1" >> gpurun_out/r02_final3_conv_kernels.txt
  ncu -i "gpurun_out/r02_final3_$1.ncu-rep" --page raw --csv 2>/dev/null | python tools/ncu_raw_extract.py >> gpurun_out/r02_final3_conv_kernels.txt
}
: > gpurun_out/r02_final3_conv_kernels.txt
cap wgrad_stacked_64_64 wgrad n64 conv_wgrad_tc_kernel
cap wgrad_stage4_conv1 wgrad big conv_wgrad_tc_kernel
cap halo64_fwd_stage4_conv1 fwd big conv_tc_halo_kernel
cat gpurun_out/r02_final3_conv_kernels.txt >> $LOG
grep -E "^exit|passed|failed|GB/s" $LOG | head -40
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02r_bench_*.json')):
  try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d['value'], d['ms_per_step'], d['config'].get('workload'), d['dtype'], d.get('parity'))
  except Exception as e:
    print(f, 'ERR', e)
PY
grep -E "^==|tensor|duration|dram__bytes" gpurun_out/r02_final3_conv_kernels.txt | head -40
