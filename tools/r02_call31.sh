#!/bin/bash
# Round 2, call 31: compute-sanitizer memcheck over smoke() on the final build (fused BCE, in-place
# stride-2 filters, stacked weight-gradient tiles, two-stream step are newer than the last memcheck).
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
LOG=gpurun_out/r02_call31.log
: > $LOG
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/r02_final_memcheck_smoke.txt \
  python -c "import __graft_entry__ as g; g.smoke()" >> $LOG 2>&1
echo "exit $? (memcheck smoke)" >> $LOG
tail -4 gpurun_out/r02_final_memcheck_smoke.txt >> $LOG
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/r02_final_memcheck_tests.txt \
  python -m pytest tests/test_gpu_bf16x3.py tests/test_gpu_next_rows.py -q -m gpu -x -k "stride2 or bce or in_place" >> $LOG 2>&1
echo "exit $? (memcheck new-kernel tests)" >> $LOG
tail -4 gpurun_out/r02_final_memcheck_tests.txt >> $LOG
grep -E "^exit|ERROR SUMMARY|passed|failed" $LOG
