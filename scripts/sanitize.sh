#!/bin/bash
# Sanitizer runs of the host engine + emulated kernels (CPU only, not part of pytest):
#   scripts/sanitize.sh asan|tsan [minutes]   |   scripts/sanitize.sh leaks
# builds tests/emul/<kind>/libedlib_emul.so and runs the stress plan (scripts/stress.py) on it with the filter paths
# forced on for small targets.  Reports go to stderr; the last line is the stress summary.
set -e
KIND=${1:-asan}; MIN=${2:-5}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
if [ "$KIND" = leaks ]; then  # C++ driver over the ASan build: batches of every mode / task, staged API, handles, error paths; LeakSanitizer at exit
  make -s -C "$ROOT/tests/emul" CXX=g++ asan/libedlib_emul.so
  g++ -O1 -g -std=c++17 -fsanitize=address -I"$ROOT/include" "$ROOT/tests/emul/leak_driver.cpp" -o "$ROOT/tests/emul/asan/leak_driver" \
      -L"$ROOT/tests/emul/asan" -Wl,-rpath,"$ROOT/tests/emul/asan" -ledlib_emul
  ASAN_OPTIONS=detect_leaks=1 exec "$ROOT/tests/emul/asan/leak_driver"
fi
make -s -C "$ROOT/tests/emul" CXX=g++ $KIND/libedlib_emul.so  # the PATH compiler, whose sanitizer runtimes gcc -print-file-name finds
if [ "$KIND" = asan ]; then
  PRE="$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so)"
  export ASAN_OPTIONS=detect_leaks=0:halt_on_error=0 UBSAN_OPTIONS=print_stacktrace=1
else
  PRE="$(gcc -print-file-name=libtsan.so)"
  export TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0"
fi
export EDLIB_B200_FILTER_MIN_TARGET=128 EDLIB_B200_FILTER_MIN_LEVEL_READS=0 EDLIB_B200_K1_MIN_GROUP=4 \
       EDLIB_B200_STREAM_MIN_PAIRS=64 EDLIB_B200_LONG_HW_MIN_TARGET=2000
EDLIB_STRESS_LIB="$ROOT/tests/emul/$KIND/libedlib_emul.so" LD_PRELOAD="$PRE" python "$ROOT/scripts/stress.py" "$MIN"
