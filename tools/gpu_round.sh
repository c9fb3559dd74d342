#!/bin/bash
# One GPU-box visit: GPU test tier, headline bench (ours + stock NCCL/cuDNN comparator).
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh [tag]
tag=${1:-r}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${tag}_gpus.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/${tag}_gpu_tier.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${tag}_gpu_tier.log
tail -5 gpurun_out/${tag}_gpu_tier.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${tag}_bench_ours.json 2> gpurun_out/${tag}_bench_ours.err
tail -1 gpurun_out/${tag}_bench_ours.json
timeout 600 python bench.py --impl nccl-cudnn --gpus 1 --steps 20 --warmup 5 > gpurun_out/${tag}_bench_base.json 2> gpurun_out/${tag}_bench_base.err
tail -1 gpurun_out/${tag}_bench_base.json; tail -3 gpurun_out/${tag}_bench_base.err
TFOS_BASELINE_GRAPH=1 timeout 600 python bench.py --impl nccl-cudnn --gpus 1 --steps 20 --warmup 5 > gpurun_out/${tag}_bench_base_graph.json 2> gpurun_out/${tag}_bench_base_graph.err
tail -1 gpurun_out/${tag}_bench_base_graph.json; tail -3 gpurun_out/${tag}_bench_base_graph.err
