#!/usr/bin/env bash
# Round 2, first GPU call (1 GPU): validate what round 1 left unvalidated.
set -u
export OMP_NUM_THREADS=1
OUT=gpurun_out/pending
mkdir -p "$OUT"
echo "== 1. experimental numerics tests"
ADAPTDL_B200_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu \
    -k "layer_norm or bitmask or phase_dgrad" > "$OUT/experimental_tests.log" 2>&1
tail -3 "$OUT/experimental_tests.log"
echo "== 2. full GPU suite (defaults)"
timeout 900 python -m pytest tests -q -m gpu > "$OUT/pytest_gpu.log" 2>&1
tail -3 "$OUT/pytest_gpu.log"
echo "== 3. headline bench variants"
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 100 --warmup 8 > "$OUT/bench_n1_$name.log" 2>&1;
  python - "$OUT/bench_n1_$name.log" <<'PY'
import json, sys
for line in open(sys.argv[1]):
    if line.startswith("{"):
        d = json.loads(line)
        print(sys.argv[1].split("/")[-1], round(d["value"]), d["unit"], round(d["ms_per_step"], 4), "ms/step e2e", round(d["e2e"]["ms_per_step"],4))
PY
}
run default A=1
run padstem ADAPTDL_B200_PAD_STEM=1
run bitmask ADAPTDL_B200_BN_BITMASK=1
run phasedgrad ADAPTDL_B200_PHASE_DGRAD=1
run all3 ADAPTDL_B200_PAD_STEM=1 ADAPTDL_B200_BN_BITMASK=1 ADAPTDL_B200_PHASE_DGRAD=1
timeout 300 python tools/conv_bench.py --out "$OUT/conv_bench.json" > "$OUT/conv_bench.log" 2>&1
tail -1 "$OUT/conv_bench.log"
echo "== 4. BERT with and without the fused LayerNorm op"
for f in default fused_ln; do
  if [ $f = fused_ln ]; then export ADAPTDL_B200_FUSED_LN=1; fi
  timeout 300 python bench.py --workload bert --steps 30 --warmup 5 > "$OUT/bench_bert_$f.log" 2>&1
  python - "$OUT/bench_bert_$f.log" <<'PY'
import json, sys
for line in open(sys.argv[1]):
    if line.startswith("{"):
        d = json.loads(line)
        print(sys.argv[1].split("/")[-1], round(d["value"]), d["unit"], round(d["ms_per_step"], 3), "ms/step")
PY
done
unset ADAPTDL_B200_FUSED_LN
echo "== 5. BN micro-benchmark"
timeout 200 python tools/bn_bench.py --out "$OUT/bn_bench_default.json" > "$OUT/bn_bench_default.log" 2>&1
ADAPTDL_B200_BN_BITMASK=1 timeout 200 python tools/bn_bench.py --out "$OUT/bn_bench_bitmask.json" > "$OUT/bn_bench_bitmask.log" 2>&1
grep -h fused_fwd_bwd_us "$OUT"/bn_bench_*.log | cut -c1-200
echo "== 6. step profile (CUPTI) resnet bf16 params"
timeout 200 python tools/step_profile.py --model resnet18 --bf16-params --out "$OUT/step_profile_resnet_bf16.json" > "$OUT/step_profile_resnet_bf16.log" 2>&1
head -30 "$OUT/step_profile_resnet_bf16.log"
timeout 200 python tools/step_profile.py --model bert --out "$OUT/step_profile_bert.json" > "$OUT/step_profile_bert.log" 2>&1
head -30 "$OUT/step_profile_bert.log"
echo "done: $OUT"
