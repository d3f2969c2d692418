#!/bin/bash
# Memcheck / racecheck of the non-tensor-core kernel SOURCES without a GPU: the emulated-kernel
# tests (tests/test_kernels_emulated_cpu.py) rebuilt with the GCC sanitizers.
#   address + alignment + bounds : out-of-bounds global / shared accesses, misaligned float4 accesses
#   thread                       : shared-memory data races (missing __syncthreads / __syncwarp)
# Run from the repo root; ~7 minutes on 8 cores.
set -u
cd "$(dirname "$0")/.."
ASAN=$(g++ -print-file-name=libasan.so); TSAN=$(g++ -print-file-name=libtsan.so)
echo "== address,alignment,bounds"
SG2IM_EMUL_CXXFLAGS='-g -DSG2IM_EMUL_THREADS -fsanitize=address,alignment,bounds -fno-sanitize-recover=alignment,bounds' \
  LD_PRELOAD=$ASAN ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 \
  python -m pytest tests/test_kernels_emulated_cpu.py tests/test_relations_cpu.py -q -x -p no:cacheprovider | tail -2
echo "== thread"
rm -f /tmp/sg2im_tsan.*
SG2IM_EMUL_CXXFLAGS='-g -DSG2IM_EMUL_THREADS -fsanitize=thread' LD_PRELOAD=$TSAN OMP_NUM_THREADS=1 \
  TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 log_path=/tmp/sg2im_tsan" \
  python -m pytest tests/test_kernels_emulated_cpu.py tests/test_relations_cpu.py -q -p no:cacheprovider | tail -2
echo "race reports: $(cat /tmp/sg2im_tsan.* 2>/dev/null | grep -c 'WARNING: ThreadSanitizer')"
echo "== tensor-core kernels (functional model, fibers): alignment,bounds"
SG2IM_EMUL_CXXFLAGS='-g -fsanitize=alignment,bounds -fno-sanitize-recover=alignment,bounds' \
  LD_PRELOAD=$(g++ -print-file-name=libubsan.so) \
  python -m pytest tests/test_tc_kernels_emulated_cpu.py -q -x -p no:cacheprovider | tail -2
