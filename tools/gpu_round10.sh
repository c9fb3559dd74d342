#!/bin/bash
# N-GPU sweep of the fused all-reduce grid (CTAs): communication cost inside the captured graph
N=${1:-4}; tag=${2:-r}
mkdir -p gpurun_out
for g in 16 32 48; do
TFOS_AR_GRID=$g timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29800+g)) bench.py --gpus $N --steps 20 --warmup 5 --no-e2e > gpurun_out/${tag}_grid$g.json 2> gpurun_out/${tag}_grid$g.err
python - <<PY
import json
try:
  r=json.loads(open("gpurun_out/${tag}_grid$g.json").read().strip().splitlines()[-1])
  print("grid=$g", round(r["value"]), round(r["ms_per_step"],3), r.get("comm"), "exposed(eager)", r.get("exposed_allreduce_ms_per_step"))
except Exception as e:
  print("grid=$g FAILED", e); print(open("gpurun_out/${tag}_grid$g.err").read()[-800:])
PY
done
