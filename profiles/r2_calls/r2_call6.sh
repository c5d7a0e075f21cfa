#!/usr/bin/env bash
set -u
export OMP_NUM_THREADS=1
OUT=gpurun_out/call7
mkdir -p "$OUT"
summ() { python - "$1" <<'PY'
import json, sys, statistics
for line in open(sys.argv[1]):
    if line.startswith("{"):
        d = json.loads(line)
        w = d["windows"]["device_ms"]; K = d["steps"]
        print(sys.argv[1].split("/")[-1], d["impl"], "N=%d" % d["n_gpus"], round(d["value"]), d["unit"],
              "ms/step %.4f (median window %.4f)" % (d["ms_per_step"], statistics.median(w) / K),
              "e2e %.4f" % d["e2e"]["ms_per_step"], "launches", d.get("gpu_launches"))
PY
}
echo "== 1. full GPU suite"
timeout 1200 python -m pytest tests -q -m gpu > "$OUT/pytest_gpu.log" 2>&1; tail -6 "$OUT/pytest_gpu.log"
echo "== 2. benches"
timeout 300 python bench.py --steps 20 --warmup 5 > "$OUT/resnet_n1.log" 2>&1; summ "$OUT/resnet_n1.log"
timeout 400 python bench.py --workload bert --steps 10 --warmup 5 > "$OUT/bert_n1.log" 2>&1; summ "$OUT/bert_n1.log"
ADAPTDL_B200_FUSED_TRANSFORMER=0 timeout 400 python bench.py --workload bert --steps 10 --warmup 5 --no-fp32-variant > "$OUT/bert_n1_nolayoutops.log" 2>&1; summ "$OUT/bert_n1_nolayoutops.log"
timeout 300 python bench.py --workload ncf --steps 20 --warmup 5 --no-fp32-variant > "$OUT/ncf_n1.log" 2>&1; summ "$OUT/ncf_n1.log"
echo "== 3. LayerNorm op"
timeout 200 python tools/ln_bench.py --out "$OUT/ln_bench.json" > "$OUT/ln_bench.log" 2>&1; cat "$OUT/ln_bench.log" | tr '\n' ' '; echo
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ln_ -s 12 -c 6 -o "$OUT/ncu_ln" python tools/ln_bench.py --iters 2 > "$OUT/ncu_ln.log" 2>&1; ls -la "$OUT" | grep ncu
echo "== 4. op profile"
timeout 300 python tools/op_profile.py --model bert --top 60 > "$OUT/op_profile_bert.log" 2>&1; grep -v "void \|nvjet\|cudnn_generated\|anonymous" "$OUT/op_profile_bert.log" | head -40 | cut -c1-200
echo "== 5. launch list resnet (eager, bf16 params)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 450 --csv --log-file "$OUT/launches_resnet.csv" python tools/step_profile.py --model resnet18 --bf16-params --steps 2 > "$OUT/ncu_resnet.log" 2>&1; wc -l "$OUT/launches_resnet.csv"
echo done
