#!/usr/bin/env bash
# First GPU call of round 2: validate everything that was written after round
# 1's GPU budget ran out (ROUND2_PLAN.md), in one go. Needs >= 2 GPUs for the
# last two steps (they are skipped on a 1-GPU box).
#
#   gpurun --gpus 2 --timeout 1500 -- 'bash tools/validate_pending.sh'
#
# Everything lands in gpurun_out/pending/.
set -u
export OMP_NUM_THREADS=1
OUT=gpurun_out/pending
mkdir -p "$OUT"
NGPU=$(python -c "import torch; print(torch.cuda.device_count())")

echo "== 1. experimental numerics tests (LayerNorm op, BN bit-mask)"
ADAPTDL_B200_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu \
    -k "layer_norm or bitmask or phase_dgrad" > "$OUT/experimental_tests.log" 2>&1
tail -3 "$OUT/experimental_tests.log"

echo "== 2. full GPU suite (defaults)"
timeout 900 python -m pytest tests -q -m gpu > "$OUT/pytest_gpu.log" 2>&1
tail -3 "$OUT/pytest_gpu.log"

echo "== 3. headline bench: defaults vs padded stem vs BN bit-mask"
timeout 300 python bench.py --steps 40 --warmup 8 > "$OUT/bench_n1_default.log" 2>&1
ADAPTDL_B200_PAD_STEM=1 timeout 300 python bench.py --steps 40 --warmup 8 > "$OUT/bench_n1_padstem.log" 2>&1
ADAPTDL_B200_BN_BITMASK=1 timeout 300 python bench.py --steps 40 --warmup 8 > "$OUT/bench_n1_bitmask.log" 2>&1
ADAPTDL_B200_PHASE_DGRAD=1 timeout 300 python bench.py --steps 40 --warmup 8 > "$OUT/bench_n1_phasedgrad.log" 2>&1
for f in default padstem bitmask phasedgrad; do
  python - "$OUT/bench_n1_$f.log" <<'PY'
import json, sys
for line in open(sys.argv[1]):
    if line.startswith("{"):
        d = json.loads(line)
        print(sys.argv[1].split("/")[-1], round(d["value"]), d["unit"], round(d["ms_per_step"], 3), "ms/step")
PY
done

timeout 300 python tools/conv_bench.py --out "$OUT/conv_bench.json" > "$OUT/conv_bench.log" 2>&1
tail -1 "$OUT/conv_bench.log"

echo "== 4. BERT with and without the fused LayerNorm op"
timeout 300 python bench.py --workload bert --steps 20 --warmup 5 > "$OUT/bench_bert_default.log" 2>&1
ADAPTDL_B200_FUSED_LN=1 timeout 300 python bench.py --workload bert --steps 20 --warmup 5 > "$OUT/bench_bert_fused_ln.log" 2>&1
for f in default fused_ln; do
  python - "$OUT/bench_bert_$f.log" <<'PY'
import json, sys
for line in open(sys.argv[1]):
    if line.startswith("{"):
        d = json.loads(line)
        print(sys.argv[1].split("/")[-1], round(d["value"]), d["unit"], round(d["ms_per_step"], 3), "ms/step")
PY
done

echo "== 5. BN micro-benchmark with and without the bit-mask"
timeout 200 python tools/bn_bench.py --out "$OUT/bn_bench_default.json" > "$OUT/bn_bench_default.log" 2>&1
ADAPTDL_B200_BN_BITMASK=1 timeout 200 python tools/bn_bench.py --out "$OUT/bn_bench_bitmask.json" > "$OUT/bn_bench_bitmask.log" 2>&1
grep -h fused_fwd_bwd_us "$OUT"/bn_bench_*.log | cut -c1-200

if [ "$NGPU" -ge 2 ]; then
  echo "== 6. N=2 headline: 25 MB (default) vs 4 MB buckets"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29521 bench.py --gpus 2 --steps 40 --warmup 8 > "$OUT/bench_n2_default.log" 2>&1
  tail -1 "$OUT/bench_n2_default.log" | cut -c1-160
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29524 bench.py --gpus 2 --steps 40 --warmup 8 --bucket-cap-mb 4 > "$OUT/bench_n2_cap4.log" 2>&1
  tail -1 "$OUT/bench_n2_cap4.log" | cut -c1-160
  echo "== 7. all-reduce stress (P2P and multimem flavours)"
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29522 tools/allreduce_stress.py --iters 5000 > "$OUT/stress_p2p.log" 2>&1
  tail -1 "$OUT/stress_p2p.log"
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29523 tools/allreduce_stress.py --iters 5000 --nvls > "$OUT/stress_nvls.log" 2>&1
  tail -1 "$OUT/stress_nvls.log"
fi
echo "done: $OUT"
