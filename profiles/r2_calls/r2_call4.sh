#!/usr/bin/env bash
# Round 2, fourth GPU call (1 GPU): thin local kernels + latency-tuned finalize, BN single-launch v2.
set -u
export OMP_NUM_THREADS=1
OUT=gpurun_out/call4
mkdir -p "$OUT"
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
echo "== 1. GPU tests (kernels touched in this call)"
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "primitives or preconditioned or non_finite or trajectory or device_engine or graphed or checkpoint or mixed or step_profile or fused_bn or resnet_step" > "$OUT/pytest_gpu.log" 2>&1
tail -6 "$OUT/pytest_gpu.log"
echo "== 2. N=1 A/B"
b1 default A=1
timeout 300 python bench.py --steps 20 --warmup 5 --no-fp32-variant --bucket-cap-mb 25 > "$OUT/n1_cap25.log" 2>&1; summ "$OUT/n1_cap25.log"
b1 nofusefin ADAPTDL_B200_FUSE_FINALIZE=0
b1 bnsingle ADAPTDL_B200_BN_SINGLE=1
b1 local32 ADAPTDL_B200_LOCAL_CTAS=32
b1 local148 ADAPTDL_B200_LOCAL_CTAS=148
echo "== 3. launch list, eager N=1 bf16 params (serialised kernels)"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 600 --csv --log-file "$OUT/launches_eager_n1.csv" \
    python tools/step_profile.py --model resnet18 --bf16-params --steps 2 > "$OUT/ncu_stepprofile.log" 2>&1
echo "== 4. BERT step profile (CUPTI), bf16 params"
timeout 300 python tools/step_profile.py --model bert --bf16-params --out "$OUT/step_profile_bert_bf16.json" > "$OUT/step_profile_bert_bf16.log" 2>&1
head -45 "$OUT/step_profile_bert_bf16.log" | cut -c1-170
echo "== 5. kernel bench (single GPU gradient kernels)"
timeout 300 python tools/kernel_bench.py > "$OUT/kernel_bench.log" 2>&1; tail -20 "$OUT/kernel_bench.log" | cut -c1-200
echo done
