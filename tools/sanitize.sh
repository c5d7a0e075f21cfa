#!/bin/bash
# compute-sanitizer passes over the single-GPU kernel tests (run on a B200:
#   gpurun --timeout 1500 -- 'bash tools/sanitize.sh'
# results land in gpurun_out/sanitize/; copy the summaries into profiles/).
#
#   memcheck   out-of-bounds / misaligned global, shared and local accesses
#   racecheck  shared-memory hazards between warps of a CTA
#   synccheck  illegal barrier use (divergent __syncthreads, bad mbarrier ops)
#   initcheck  reads of uninitialised global memory
#
# The sanitizer serialises kernels and is 10-100x slower, so only the small-shape
# numerics tests run under it, each tool with its own timeout. The multi-GPU
# kernels spin on peer flags: they are covered by tools/allreduce_stress.py, not here.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/sanitize
mkdir -p "$OUT"
SELECT=${SANITIZE_TESTS:-"primitives_match or preconditioned_statistics or fused_optimizers or fused_bn_act_matches or tcgen05_linear_gelu_forward"}
STATUS=0
for tool in ${SANITIZE_TOOLS:-memcheck racecheck synccheck initcheck}; do
  echo "== $tool"
  timeout "${SANITIZE_TIMEOUT:-900}" compute-sanitizer --tool "$tool" --error-exitcode 9 \
      --launch-timeout 120 --log-file "$OUT/$tool.log" \
      python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "$SELECT" \
      -p no:cacheprovider > "$OUT/$tool.pytest.log" 2>&1
  rc=$?
  errors=$(grep -c "========= .*error\|========= Invalid\|========= Race\|========= Uninitialized" "$OUT/$tool.log" 2>/dev/null || true)
  echo "$tool: exit $rc, ${errors:-0} reports ($(tail -1 "$OUT/$tool.pytest.log" | cut -c1-100))"
  grep "ERROR SUMMARY\|RACECHECK SUMMARY" "$OUT/$tool.log" | tail -1
  [ "$rc" -ne 0 ] && STATUS=1
done
exit $STATUS
