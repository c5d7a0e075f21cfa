#!/usr/bin/env bash
# Round 2, 8-GPU call: numerics at N=8 (P2P, one-shot, NVLS), stress, all-reduce flavour table,
# ResNet-18 and BERT at N=8 (both arms), N=4 points.
set -u
export OMP_NUM_THREADS=1
OUT=gpurun_out/n8
mkdir -p "$OUT"
N=${N:-8}
TR() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 "$@"; }
summ() { python - "$1" <<'PY'
import json, sys, statistics
for line in open(sys.argv[1]):
    if line.startswith("{"):
        d = json.loads(line)
        if "unavailable" in d: print(sys.argv[1], d); continue
        w = d["windows"]["device_ms"]; K = d["steps"]
        print(sys.argv[1].split("/")[-1], d["impl"], "N=%d" % d["n_gpus"], round(d["value"]), d["unit"],
              "ms/step %.4f (median window %.4f)" % (d["ms_per_step"], statistics.median(w) / K),
              "e2e %.4f" % d["e2e"]["ms_per_step"], "buckets", d.get("buckets"), "nvls", d.get("nvls_launches"),
              "oneshot", d.get("oneshot_launches"), "fp32var", (d.get("fp32_grad_variant") or {}).get("ms_per_step"))
PY
}
port=29700
bench() { name=$1; n=$2; shift 2; port=$((port+1)); timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 20 --warmup 5 "$@" > "$OUT/$name.log" 2>&1; summ "$OUT/$name.log"; }
echo "== 1. numerics at N=$N (native P2P / one-shot, NVLS)"
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "multi_gpu and native" > "$OUT/pytest_multigpu.log" 2>&1; tail -4 "$OUT/pytest_multigpu.log"
echo "== 2. stress"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29722 tools/allreduce_stress.py --iters 1500 > "$OUT/stress_p2p.log" 2>&1; tail -1 "$OUT/stress_p2p.log"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29723 tools/allreduce_stress.py --iters 800 --nvls > "$OUT/stress_nvls.log" 2>&1; tail -1 "$OUT/stress_nvls.log"
echo "== 3. headline ResNet-18"
bench resnet_own_n8 $N
ADAPTDL_B200_REDUCE_CTAS=16 bench resnet_own_n8_ctas16 $N --no-fp32-variant
bench resnet_ref_n8 $N --impl reference
echo "== 4. all-reduce flavours (graph replays)"
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29750 tools/allreduce_bench.py --dtype bf16 --sizes-mb 0.0625,1,4,16,64,256 --sweep --nvls-ctas 64,96 --out "$OUT/allreduce_n8.json" > "$OUT/allreduce_n8.log" 2>&1
python - "$OUT/allreduce_n8.json" <<'PY'
import json, sys
try:
    for r in json.load(open(sys.argv[1])):
        print(r["MB"], "picked", r["picked"], "nccl", round(r["nccl_us"],1), {k: (round(v["isolated_us"],1), round(v["pipelined_us"],1), round(v["pipelined_frac_of_770"],2)) for k, v in r["variants"].items()})
except Exception as e: print("no allreduce json", e); print(open(sys.argv[1].replace(".json",".log")).read()[-3000:])
PY
echo "== 5. BERT"
bench bert_own_n8 $N --workload bert --steps 10 --no-fp32-variant
bench bert_ref_n8 $N --workload bert --steps 10 --impl reference
echo done
