#!/usr/bin/env bash
# Round 2, second GPU call (2 GPUs): validate the v2 gradient path.
set -u
export OMP_NUM_THREADS=1
OUT=gpurun_out/call2
mkdir -p "$OUT"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
summ() { python - "$1" <<'PY'
import json, sys
for line in open(sys.argv[1]):
    if line.startswith("{"):
        d = json.loads(line)
        if "unavailable" in d: print(sys.argv[1], d); continue
        v = d.get("fp32_grad_variant") or {}
        print(sys.argv[1].split("/")[-1], d["impl"], "N=%d" % d["n_gpus"], round(d["value"]), d["unit"],
              "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4),
              "fp32var", round(v.get("ms_per_step", 0), 4), "buckets", d.get("buckets"),
              "oneshot", d.get("oneshot_launches"), "launches", d.get("gpu_launches"),
              "timed_profile", d.get("device_timed_profile_steps"))
PY
}
echo "== 1. GPU test suite"
timeout 1500 python -m pytest tests -q -m gpu -x > "$OUT/pytest_gpu.log" 2>&1
tail -5 "$OUT/pytest_gpu.log"
echo "== 2. stress"
timeout 400 $TR --master-port 29522 tools/allreduce_stress.py --iters 4000 > "$OUT/stress_p2p.log" 2>&1
tail -2 "$OUT/stress_p2p.log"
timeout 400 $TR --master-port 29523 tools/allreduce_stress.py --iters 3000 --nvls > "$OUT/stress_nvls.log" 2>&1
tail -2 "$OUT/stress_nvls.log"
echo "== 3. headline N=1"
timeout 300 python bench.py --steps 20 --warmup 5 > "$OUT/bench_n1.log" 2>&1; summ "$OUT/bench_n1.log"
echo "== 4. headline N=2, bucket caps"
port=29530
for cap in 25 8 4 2; do
  port=$((port+1))
  timeout 300 $TR --master-port $port bench.py --gpus 2 --steps 20 --warmup 5 --bucket-cap-mb $cap --no-fp32-variant > "$OUT/bench_n2_cap$cap.log" 2>&1; summ "$OUT/bench_n2_cap$cap.log"
done
ADAPTDL_B200_FUSE_FINALIZE=0 timeout 300 $TR --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 --bucket-cap-mb 4 --no-fp32-variant > "$OUT/bench_n2_cap4_nofuse.log" 2>&1; summ "$OUT/bench_n2_cap4_nofuse.log"
ADAPTDL_B200_EAGER_BUFFER_BCAST=1 timeout 300 $TR --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 --bucket-cap-mb 4 --no-fp32-variant > "$OUT/bench_n2_cap4_eagerbcast.log" 2>&1; summ "$OUT/bench_n2_cap4_eagerbcast.log"
ADAPTDL_B200_ONESHOT_KB=0 timeout 300 $TR --master-port 29543 bench.py --gpus 2 --steps 20 --warmup 5 --bucket-cap-mb 4 --no-fp32-variant > "$OUT/bench_n2_cap4_nooneshot.log" 2>&1; summ "$OUT/bench_n2_cap4_nooneshot.log"
ADAPTDL_B200_REDUCE_CTAS=16 timeout 300 $TR --master-port 29544 bench.py --gpus 2 --steps 20 --warmup 5 --bucket-cap-mb 4 --no-fp32-variant > "$OUT/bench_n2_cap4_ctas16.log" 2>&1; summ "$OUT/bench_n2_cap4_ctas16.log"
echo "== 5. all-reduce flavours N=2"
timeout 400 $TR --master-port 29550 tools/allreduce_bench.py --sizes-mb 0.0625,0.25,1,4,16,64 --sweep --nvls-ctas 64 --out "$OUT/allreduce_n2.json" > "$OUT/allreduce_n2.log" 2>&1
python - "$OUT/allreduce_n2.json" <<'PY'
import json, sys
try:
    for r in json.load(open(sys.argv[1])):
        print(r["MB"], "picked", r["picked"], "nccl", round(r["nccl_us"],1), {k: (round(v["isolated_us"],1), round(v["pipelined_us"],1)) for k, v in r["variants"].items()})
except Exception as e: print("no allreduce json", e)
PY
echo "== 6. BERT / NCF N=1 both arms"
timeout 400 python bench.py --workload bert --steps 10 --warmup 5 > "$OUT/bench_bert_own_n1.log" 2>&1; summ "$OUT/bench_bert_own_n1.log"
timeout 400 python bench.py --workload bert --impl reference --steps 10 --warmup 5 > "$OUT/bench_bert_ref_n1.log" 2>&1; summ "$OUT/bench_bert_ref_n1.log"
timeout 300 python bench.py --workload ncf --steps 20 --warmup 5 > "$OUT/bench_ncf_own_n1.log" 2>&1; summ "$OUT/bench_ncf_own_n1.log"
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > "$OUT/bench_ref_n1.log" 2>&1; summ "$OUT/bench_ref_n1.log"
echo "== 7. BERT N=2"
timeout 400 $TR --master-port 29560 bench.py --gpus 2 --workload bert --steps 10 --warmup 5 > "$OUT/bench_bert_own_n2.log" 2>&1; summ "$OUT/bench_bert_own_n2.log"
echo done
