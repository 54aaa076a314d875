#!/bin/bash
# compute-sanitizer targets for the device library (memcheck / racecheck / initcheck / synccheck).
#
#   gpurun --timeout 1500 -- 'bash tools/sanitize_gpu.sh [memcheck|racecheck|initcheck|synccheck] [pytest -k expression]'
#
# Runs the small-shape GPU parity tests (every kernel family: SIMT and tcgen05 convolutions, GEMM, pooling, BatchNorm,
# element-wise, loss, optimizer) under the chosen tool and writes gpurun_out/sanitize_<tool>.log; the exit code is the
# sanitizer's (non-zero on any report: --error-exitcode 66).  The full-size tests are left out (the tools slow kernels
# down 10-100x).  NOT RUN in round 2 (the GPU budget was spent on measurements): the CPU counterparts of memcheck and
# racecheck on the kernel sources are tools/kernel_memcheck.py, tools/kernel_race_check.py and tools/kernel_ubcheck.py
# (profiles/r2_memcheck.txt, r2_race_check.txt, r2_ubcheck.txt); the tcgen05 / TMA kernels are only reachable by this
# script.
set -u
TOOL="${1:-memcheck}"
KEXPR="${2:-cuda}"
mkdir -p gpurun_out
LOG="gpurun_out/sanitize_${TOOL}.log"
EXTRA=""
case "$TOOL" in
  memcheck)  EXTRA="--leak-check no --report-api-errors no" ;;
  racecheck) EXTRA="--racecheck-report all" ;;
  initcheck) EXTRA="--track-unused-memory no" ;;
  synccheck) EXTRA="" ;;
  *) echo "unknown tool $TOOL" >&2; exit 2 ;;
esac
compute-sanitizer --tool "$TOOL" $EXTRA --error-exitcode 66 --target-processes all --log-file "$LOG" \
  python -m pytest -q -x -m gpu -p no:cacheprovider -k "$KEXPR" \
    tests/test_device_parity.py tests/test_tensor_core_conv.py tests/test_first_layer_conv.py tests/test_narrow_conv.py \
    tests/test_lazy_batchnorm.py tests/test_fused_layers.py
RC=$?
echo "sanitize_gpu: tool=$TOOL pytest/sanitizer rc=$RC, $(grep -c 'ERROR SUMMARY' "$LOG" 2>/dev/null) summary line(s) in $LOG"
grep "ERROR SUMMARY" "$LOG" 2>/dev/null | tail -3
exit $RC
