#!/bin/bash
# 1-GPU visit: serpentine A/B on the headline, kernel numerics that the change touches, TFModel inference.
tag=${1:-r}
mkdir -p gpurun_out
timeout 300 python tools/gpu_check.py batchnorm conv_dgrad resnet_step resnet50_grad_parity > gpurun_out/${tag}_checks.log 2>&1; grep -c "OK$" gpurun_out/${tag}_checks.log; grep "FAIL\|rror" gpurun_out/${tag}_checks.log | head
for s in 1 0 1 0; do
TFOS_L2_SERPENTINE=$s timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --no-e2e > gpurun_out/${tag}_bench_serp$s.json 2> gpurun_out/${tag}_bench_serp$s.err
python - <<PY
import json
r=json.loads(open("gpurun_out/${tag}_bench_serp$s.json").read().strip().splitlines()[-1])
print("serpentine=$s", round(r["value"],1), "img/s", round(r["ms_per_step"],3), "ms", r["clocks"])
PY
done
timeout 500 python bench.py --config infer --gpus 1 --steps 20 > gpurun_out/${tag}_infer.json 2> gpurun_out/${tag}_infer.err; tail -c 1200 gpurun_out/${tag}_infer.json; tail -3 gpurun_out/${tag}_infer.err | cut -c1-300
