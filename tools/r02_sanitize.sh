#!/bin/bash
# compute-sanitizer passes over the CUDA path (race detection / sanitizers), separate from
# the first call so that one stays ~30 minutes:
#   gpurun --timeout 1500 -- 'bash tools/r02_sanitize.sh'          (one GPU, bounded at ~20 GPU-minutes)
set -u
mkdir -p gpurun_out
: > gpurun_out/r02_sanitize.log
echo "== compute-sanitizer memcheck: smoke() (one small G+D iteration on the fp32 and the tcgen05 paths)" >> gpurun_out/r02_sanitize.log
timeout 540 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/r02_memcheck_smoke.txt \
  python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/r02_sanitize.log 2>&1
echo "exit $? (memcheck smoke)" >> gpurun_out/r02_sanitize.log
echo "== compute-sanitizer racecheck: graph pooling / layout / crop / BN unit tests" >> gpurun_out/r02_sanitize.log
timeout 540 compute-sanitizer --tool racecheck --error-exitcode 7 --log-file gpurun_out/r02_racecheck_ops.txt \
  python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "graph_pool_empty or gconv_layer or layout_golden or crop_golden or upsample_then_bn" >> gpurun_out/r02_sanitize.log 2>&1
echo "exit $? (racecheck ops)" >> gpurun_out/r02_sanitize.log
tail -20 gpurun_out/r02_sanitize.log
