#!/bin/bash
# The Arrow-level host code (arrow_bridge.cpp, compiled against the stub CUDA header of the CPU harness) under
# AddressSanitizer + UndefinedBehaviorSanitizer (default) or ThreadSanitizer (`bridge_sanitize.sh tsan`): runs
# tests/test_bridge_host.py with the instrumented library.  Python itself is not instrumented, so the sanitizer
# runtime is preloaded and leak detection is off.
GCC=/usr/bin/gcc; [ -x $GCC ] || GCC=gcc
if [ "$1" = "tsan" ]; then
  shift
  export LD_PRELOAD="$($GCC -print-file-name=libtsan.so)"
  export TSAN_OPTIONS="halt_on_error=1:report_signal_unsafe=0"
  export PB_BRIDGE_CXXFLAGS="-g -fno-omit-frame-pointer -fsanitize=thread"
else
  export LD_PRELOAD="$($GCC -print-file-name=libasan.so) $($GCC -print-file-name=libubsan.so)"
  export ASAN_OPTIONS=detect_leaks=0:abort_on_error=1:halt_on_error=1
  export UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1
  export PB_BRIDGE_CXXFLAGS="-g -fno-omit-frame-pointer -fsanitize=address,undefined -fno-sanitize-recover=undefined"
fi
exec python -m pytest tests/test_bridge_host.py -x -q "$@"
