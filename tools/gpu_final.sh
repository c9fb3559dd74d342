#!/bin/bash
# What the driver runs at round end, on a 2-GPU lease so that the multi-GPU tiers run too.
tag=${1:-fin}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/${tag}_gpu_tier.log 2>&1; tail -4 gpurun_out/${tag}_gpu_tier.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${tag}_bench1.json 2> gpurun_out/${tag}_bench1.err; tail -1 gpurun_out/${tag}_bench1.json | cut -c1-700
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29911 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/${tag}_bench2.json 2> gpurun_out/${tag}_bench2.err; tail -1 gpurun_out/${tag}_bench2.json | cut -c1-400; tail -3 gpurun_out/${tag}_bench2.err | cut -c1-300
timeout 200 python bench.py --impl reference --gpus 1 --steps 2 --warmup 3
# cross-host paths on the GPUs of this box (docs/deployment.md)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29913 tools/gpu_check_hier.py 2>&1 | grep "CHECK\|HIER" | tail -14
