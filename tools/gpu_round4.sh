#!/bin/bash
tag=${1:-r}
mkdir -p gpurun_out
timeout 400 python tools/gpu_check.py dgrad_bn_reduce conv_dgrad dgrad_masked_accumulate resnet_step resnet50_grad_parity ps_kernels > gpurun_out/${tag}_checks.log 2>&1; grep -c "OK$" gpurun_out/${tag}_checks.log; grep "FAIL\|rror" gpurun_out/${tag}_checks.log | head -20
for f in 1 0 1 0; do
TFOS_BN_FUSED_REDUCE=$f timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --no-e2e > gpurun_out/${tag}_bench_fr$f.json 2> gpurun_out/${tag}_bench_fr$f.err
python - <<PY
import json
try:
  r=json.loads(open("gpurun_out/${tag}_bench_fr$f.json").read().strip().splitlines()[-1])
  print("fused_reduce=$f", round(r["value"],1), "img/s", round(r["ms_per_step"],3), "ms launches", r["launches_per_step"], "loss", r["final_loss"], r["clocks"]["reasons"])
except Exception as e:
  print("fused_reduce=$f FAILED", e); print(open("gpurun_out/${tag}_bench_fr$f.err").read()[-1500:])
PY
done
