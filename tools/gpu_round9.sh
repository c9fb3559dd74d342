#!/bin/bash
# 2-GPU visit: multi-GPU + fault test tiers, comparator at N=2, headline at N=2
tag=${1:-r}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_fault.py -q -p no:cacheprovider > gpurun_out/${tag}_multi_fault.log 2>&1; tail -15 gpurun_out/${tag}_multi_fault.log
run() { timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29702 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/${tag}_ours.json 2> gpurun_out/${tag}_ours.err; tail -1 gpurun_out/${tag}_ours.json | cut -c1-300
run 29705 bench.py --impl nccl-cudnn --gpus 2 --steps 20 --warmup 5 > gpurun_out/${tag}_base.json 2> gpurun_out/${tag}_base.err; tail -1 gpurun_out/${tag}_base.json | cut -c1-300
