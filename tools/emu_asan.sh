#!/bin/bash
# Runs the CPU emulation tests of the device kernels (tests/test_emu_*.py) with the emulation
# library built under AddressSanitizer + UBSan: out-of-bounds global / shared-memory accesses in
# the kernel headers show up here the way compute-sanitizer memcheck reports them on a GPU.
# Usage: tools/emu_asan.sh [pytest args]   (default: all tests/test_emu_*.py)
set -e
cd "$(dirname "$0")/.."
export HYP_EMU_ASAN=1
export ASAN_OPTIONS=detect_leaks=0:halt_on_error=1
export UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1
export LD_PRELOAD="$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so)"
if [ $# -eq 0 ]; then set -- tests/test_emu_*.py; fi
exec python -m pytest -x -q -p no:cacheprovider "$@"
