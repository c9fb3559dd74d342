#!/bin/bash
# 2-GPU visit: PS kernels, fused collectives incl. NVLS, slot-mode PS bench, 2-GPU headline with
# and without NVLS.
tag=${1:-r}
mkdir -p gpurun_out
timeout 300 python tools/gpu_check.py ps_kernels > gpurun_out/${tag}_ps_kernels.log 2>&1; grep -c "OK$" gpurun_out/${tag}_ps_kernels.log; grep "FAIL\|Error\|error" gpurun_out/${tag}_ps_kernels.log | head
TFOS_FULL=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/gpu_check_multi.py > gpurun_out/${tag}_multi.log 2>&1
grep "TIMING\|NVLS\|SUMMARY\|FAIL\|rror" gpurun_out/${tag}_multi.log | head -30
timeout 600 python bench/ps_resnet50.py --gpus 2 --steps 10 > gpurun_out/${tag}_ps_bench.log 2>&1; tail -3 gpurun_out/${tag}_ps_bench.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/${tag}_bench2_nvls.json 2> gpurun_out/${tag}_bench2_nvls.err; tail -1 gpurun_out/${tag}_bench2_nvls.json
TFOS_NVLS=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/${tag}_bench2_p2p.json 2> gpurun_out/${tag}_bench2_p2p.err; tail -1 gpurun_out/${tag}_bench2_p2p.json
