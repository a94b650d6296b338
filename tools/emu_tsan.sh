#!/bin/bash
# Runs the CPU emulation tests of the device kernels with the emulation library built under
# ThreadSanitizer. The emulation maps every CUDA thread to a pthread and __syncthreads /
# __syncwarp / shuffles to pthread barriers, so a shared- or global-memory hand-off between
# threads of a block that lacks a barrier is reported as a data race (the host stand-in for
# compute-sanitizer racecheck). Usage: tools/emu_tsan.sh [pytest args]
set -e
cd "$(dirname "$0")/.."
export HYP_EMU_TSAN=1
export TSAN_OPTIONS="halt_on_error=0:report_signal_unsafe=0:history_size=4:log_path=${TSAN_LOG:-/tmp/emu_tsan}"
# NumPy's OpenBLAS worker pool is not instrumented and would be reported; keep BLAS on the calling thread
export OPENBLAS_NUM_THREADS=1 OMP_NUM_THREADS=1
export LD_PRELOAD="$(gcc -print-file-name=libtsan.so)"
if [ $# -eq 0 ]; then set -- tests/test_emu_*.py; fi
python -m pytest -x -q -p no:cacheprovider "$@"
echo "ThreadSanitizer reports: $(ls ${TSAN_LOG:-/tmp/emu_tsan}.* 2>/dev/null | wc -l) file(s) under ${TSAN_LOG:-/tmp/emu_tsan}.*"
