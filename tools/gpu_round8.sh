#!/bin/bash
# N-GPU visit (N = 2, 4 or 8): headline (ours, with NVLS auto / forced / off), the stock NCCL+cuDNN
# comparator, the stand-alone fused all-reduce timing, and at N = 8 the secondary BASELINE configs.
N=${1:-8}; tag=${2:-r8}
mkdir -p gpurun_out
run() { timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
TFOS_FULL=1 run 29601 tools/gpu_check_multi.py > gpurun_out/${tag}_multi.log 2>&1; grep "TIMING\|SUMMARY\|FAIL" gpurun_out/${tag}_multi.log | head -12
run 29602 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${tag}_ours.json 2> gpurun_out/${tag}_ours.err; tail -1 gpurun_out/${tag}_ours.json | cut -c1-2500
TFOS_NVLS=1 run 29604 bench.py --gpus $N --steps 20 --warmup 5 --no-e2e > gpurun_out/${tag}_ours_nvls.json 2> gpurun_out/${tag}_ours_nvls.err; tail -1 gpurun_out/${tag}_ours_nvls.json | cut -c1-2500
run 29605 bench.py --impl nccl-cudnn --gpus $N --steps 20 --warmup 5 > gpurun_out/${tag}_base.json 2> gpurun_out/${tag}_base.err; tail -1 gpurun_out/${tag}_base.json | cut -c1-1200
if [ "$N" = "8" ]; then
timeout 500 python bench.py --config ps --gpus 8 --steps 20 > gpurun_out/${tag}_ps.json 2> gpurun_out/${tag}_ps.err; tail -1 gpurun_out/${tag}_ps.json | cut -c1-1500
timeout 500 python bench.py --config infer --gpus 8 --steps 24 > gpurun_out/${tag}_infer.json 2> gpurun_out/${tag}_infer.err; tail -1 gpurun_out/${tag}_infer.json | cut -c1-800
timeout 600 python bench.py --config unet --gpus 8 --steps 60 > gpurun_out/${tag}_unet.json 2> gpurun_out/${tag}_unet.err; tail -1 gpurun_out/${tag}_unet.json | cut -c1-1200
fi
