#!/bin/bash
tag=${1:-r}
mkdir -p gpurun_out
timeout 300 python tools/gpu_check.py stem_fused resnet_step resnet50_grad_parity folded_inference > gpurun_out/${tag}_checks.log 2>&1; echo "checks ok=$(grep -c 'OK$' gpurun_out/${tag}_checks.log)"; grep "FAIL\|rror" gpurun_out/${tag}_checks.log | head
for f in 1 0 1 0; do
TFOS_FUSE_STEM=$f timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --no-e2e > gpurun_out/${tag}_bench_stem$f.json 2> gpurun_out/${tag}_bench_stem$f.err
python - <<PY
import json
try:
  r=json.loads(open("gpurun_out/${tag}_bench_stem$f.json").read().strip().splitlines()[-1])
  print("fuse_stem=$f", round(r["value"],1), "img/s", round(r["ms_per_step"],3), "ms launches", r["launches_per_step"], "loss", r["final_loss"])
except Exception as e:
  print("fuse_stem=$f FAILED", e); print(open("gpurun_out/${tag}_bench_stem$f.err").read()[-1500:])
PY
done
