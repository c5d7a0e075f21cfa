#!/bin/bash
# Sanitizer passes over the HOST native code (csrc/host/adl_pollux.cpp: the
# C++ core of the Pollux search, a thread pool over candidate allocations).
# Runs anywhere g++ does -- no GPU:
#
#   bash tools/sanitize_host.sh            # logs under profiles/r2_sanitize/
#
#   asan+ubsan  heap / stack out-of-bounds, use-after-free, signed overflow,
#               misaligned and null accesses, invalid shifts
#   tsan        data races between the pool's worker threads
#
# Each variant is built next to the logs (never over the in-tree library) and
# loaded through ADAPTDL_B200_HOST_LIB; the sanitizer runtime has to be in the
# process before Python, hence LD_PRELOAD. The workload is the policy
# test-suite (C++ repair / mutate rules against the numpy rules entry by
# entry, whole optimisation cycles, non-preemptible jobs) plus two full
# search cycles of tools/policy_bench.py.
set -u
cd "$(dirname "$0")/.."
OUT=${SANITIZE_OUT:-profiles/r2_sanitize}
mkdir -p "$OUT"
SRC=csrc/host/adl_pollux.cpp
# not $CXX: a toolchain without the sanitizer runtimes may be configured there
CXX=${SANITIZE_CXX:-/usr/bin/g++}
STATUS=0
run() {   # name, sanitizer flags, runtime library, runtime options
  local name=$1 flags=$2 runtime=$3 opts=$4
  local lib="$OUT/libadl_host_$name.so"
  echo "== $name"
  $CXX -O1 -g -fno-omit-frame-pointer $flags -std=c++17 -fPIC -shared -pthread \
      -o "$lib" "$SRC" || { STATUS=1; return; }
  local pre
  pre=$($CXX -print-file-name=$runtime)
  env LD_PRELOAD="$pre" $opts ADAPTDL_B200_HOST_LIB="$PWD/$lib" \
      python -m pytest tests/test_policy_native.py tests/test_policy.py -q \
      -p no:cacheprovider > "$OUT/$name.pytest.log" 2> "$OUT/$name.stderr.log"
  local rc1=$?
  env LD_PRELOAD="$pre" $opts ADAPTDL_B200_HOST_LIB="$PWD/$lib" \
      python tools/policy_bench.py --sizes 60x16 --gpus-per-node 8 \
      --cycles 2 --search native > "$OUT/$name.policy_bench.log" \
      2>> "$OUT/$name.stderr.log"
  local rc2=$?
  local reports
  reports=$(grep -c "ERROR: AddressSanitizer\|runtime error:\|WARNING: ThreadSanitizer" \
      "$OUT/$name.stderr.log" "$OUT/$name.pytest.log" "$OUT/$name.policy_bench.log" \
      | awk -F: '{s+=$2} END {print s+0}')
  echo "$name: pytest exit $rc1 ($(tail -1 "$OUT/$name.pytest.log")), policy_bench exit $rc2, $reports sanitizer reports" \
      | tee "$OUT/$name.summary.txt"
  rm -f "$lib"
  [ "$rc1" -ne 0 ] || [ "$rc2" -ne 0 ] || [ "$reports" -ne 0 ] && STATUS=1
}
run asan_ubsan "-fsanitize=address,undefined -fno-sanitize-recover=undefined" libasan.so \
    "ASAN_OPTIONS=detect_leaks=0:abort_on_error=0 UBSAN_OPTIONS=print_stacktrace=1"
run tsan "-fsanitize=thread" libtsan.so \
    "TSAN_OPTIONS=report_signal_unsafe=0:exitcode=0"
exit $STATUS
