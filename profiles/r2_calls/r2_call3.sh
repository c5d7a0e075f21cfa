#!/usr/bin/env bash
# Round 2, third GPU call (2 GPUs): full GPU suite + A/B of the step-time levers.
set -u
export OMP_NUM_THREADS=1
OUT=gpurun_out/call3
mkdir -p "$OUT"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
summ() { python - "$1" <<'PY'
import json, sys, statistics
for line in open(sys.argv[1]):
    if line.startswith("{"):
        d = json.loads(line)
        if "unavailable" in d: print(sys.argv[1], d); continue
        w = d["windows"]["device_ms"]; K = d["steps"]
        print(sys.argv[1].split("/")[-1], d["impl"], "N=%d" % d["n_gpus"], round(d["value"]), d["unit"],
              "ms/step %.4f (median window %.4f)" % (d["ms_per_step"], statistics.median(w) / K),
              "e2e %.4f" % d["e2e"]["ms_per_step"], "buckets", d.get("buckets"),
              "launches", d.get("gpu_launches"), "booked", d.get("device_timed_profile_steps"))
PY
}
b1() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-fp32-variant > "$OUT/n1_$name.log" 2>&1; summ "$OUT/n1_$name.log"; }
port=29600
b2() { name=$1; shift; port=$((port+1)); env "$@" timeout 300 $TR --master-port $port bench.py --gpus 2 --steps 20 --warmup 5 --no-fp32-variant > "$OUT/n2_$name.log" 2>&1; summ "$OUT/n2_$name.log"; }
echo "== 1. GPU test suite (all)"
timeout 1800 python -m pytest tests -q -m gpu > "$OUT/pytest_gpu.log" 2>&1
tail -8 "$OUT/pytest_gpu.log"
echo "== 2. N=1 A/B"
b1 default A=1
b1 bn2k_pdl ADAPTDL_B200_BN_SINGLE=0
b1 bn2k_nopdl ADAPTDL_B200_BN_SINGLE=0 ADAPTDL_B200_BN_PDL=0
b1 nofusefin ADAPTDL_B200_FUSE_FINALIZE=0
b1 nofusefin_bn2k_nopdl ADAPTDL_B200_FUSE_FINALIZE=0 ADAPTDL_B200_BN_SINGLE=0 ADAPTDL_B200_BN_PDL=0
echo "== 3. launch list, eager N=1 (serialised kernels)"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 4500 -c 700 --csv --log-file "$OUT/launches_eager_n1.csv" \
    python bench.py --steps 4 --warmup 3 --no-graph --no-fp32-variant --min-timed-ms 0 > "$OUT/ncu_bench.log" 2>&1
python tools/ncu_summary.py "$OUT/launches_eager_n1.csv" 2>/dev/null | head -40
echo "== 4. N=2 A/B"
b2 default A=1
b2 thr512 ADAPTDL_B200_REDUCE_THREADS=512
b2 ctas16 ADAPTDL_B200_REDUCE_CTAS=16
b2 ctas8 ADAPTDL_B200_REDUCE_CTAS=8
b2 bn2k ADAPTDL_B200_BN_SINGLE=0
b2 ctas16_bn2k ADAPTDL_B200_REDUCE_CTAS=16 ADAPTDL_B200_BN_SINGLE=0
echo "== 5. all-reduce flavours N=2 (graph replays)"
timeout 400 $TR --master-port 29650 tools/allreduce_bench.py --sizes-mb 0.0625,0.25,1,4,16,64 --out "$OUT/allreduce_n2.json" > "$OUT/allreduce_n2.log" 2>&1
python - "$OUT/allreduce_n2.json" <<'PY'
import json, sys
try:
    for r in json.load(open(sys.argv[1])):
        print(r["MB"], "picked", r["picked"], "nccl", round(r["nccl_us"],1), {k: (round(v["isolated_us"],1), round(v["pipelined_us"],1), round(v["isolated_frac_of_770"],2)) for k, v in r["variants"].items()})
except Exception as e: print("no allreduce json", e); print(open(sys.argv[1].replace(".json",".log")).read()[-2000:])
PY
echo "== 6. BERT N=1 / N=2"
timeout 400 python bench.py --workload bert --steps 10 --warmup 5 --no-fp32-variant > "$OUT/bert_n1.log" 2>&1; summ "$OUT/bert_n1.log"
timeout 400 $TR --master-port 29660 bench.py --gpus 2 --workload bert --steps 10 --warmup 5 --no-fp32-variant > "$OUT/bert_n2.log" 2>&1; summ "$OUT/bert_n2.log"
echo done
