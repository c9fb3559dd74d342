#!/bin/bash
# CTA-pair (cta_group::2) bring-up: numerics with the pair kernel forced wherever legal, then the
# per-layer table and the headline with it on (auto) / off.
tag=${1:-r}
mkdir -p gpurun_out
TFOS_IGEMM_2CTA=1 timeout 240 python tools/gpu_check.py gemm_nk gemm_kn gemm_stats_accumulate conv_fprop conv_dgrad dgrad_masked_accumulate dgrad_bn_reduce > gpurun_out/${tag}_checks_2cta.log 2>&1
echo "forced 2cta checks rc=$? ok=$(grep -c 'OK$' gpurun_out/${tag}_checks_2cta.log)"; grep "FAIL\|rror\|timeout\|tfos:" gpurun_out/${tag}_checks_2cta.log | head -20
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
TFOS_IGEMM_2CTA=2 timeout 300 python tools/bench_igemm.py > gpurun_out/${tag}_layers_2cta.txt 2>&1; tail -28 gpurun_out/${tag}_layers_2cta.txt
TFOS_IGEMM_2CTA=0 timeout 300 python tools/bench_igemm.py > gpurun_out/${tag}_layers_1cta.txt 2>&1; grep "l3c2\|l4c2\|l4c3\|l3c1 " gpurun_out/${tag}_layers_1cta.txt
for m in 2 0 2 0; do
TFOS_IGEMM_2CTA=$m timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --no-e2e > gpurun_out/${tag}_bench_2cta$m.json 2> gpurun_out/${tag}_bench_2cta$m.err
python - <<PY
import json
try:
  r=json.loads(open("gpurun_out/${tag}_bench_2cta$m.json").read().strip().splitlines()[-1])
  print("2cta mode=$m", round(r["value"],1), "img/s", round(r["ms_per_step"],3), "ms loss", r["final_loss"], r["clocks"]["reasons"])
except Exception as e:
  print("2cta mode=$m FAILED", e); print(open("gpurun_out/${tag}_bench_2cta$m.err").read()[-1500:])
PY
done
