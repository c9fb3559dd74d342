#!/bin/bash
# ncu --set full captures of the round-2 kernels (one GPU; numbers under ncu are not bench values)
tag=${1:-r}
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
TFOS_IGEMM_2CTA=2 timeout 300 $NCU -k regex:igemm_fwd_kernel -s 2 -c 1 -f -o gpurun_out/${tag}_l3c2_fprop_2cta python tools/bench_igemm.py --only l3c2 > gpurun_out/${tag}_ncu1.log 2>&1; tail -2 gpurun_out/${tag}_ncu1.log
TFOS_IGEMM_2CTA=0 timeout 300 $NCU -k regex:igemm_fwd_kernel -s 2 -c 1 -f -o gpurun_out/${tag}_l3c2_fprop_1cta python tools/bench_igemm.py --only l3c2 > gpurun_out/${tag}_ncu2.log 2>&1; tail -2 gpurun_out/${tag}_ncu2.log
timeout 400 $NCU --profile-from-start off --kernel-name-base mangled -k regex:ILi256ELb1ELi10E -c 1 -f -o gpurun_out/${tag}_dgrad_acc_bnred python tools/step_profile.py > gpurun_out/${tag}_ncu3.log 2>&1; tail -2 gpurun_out/${tag}_ncu3.log
timeout 400 $NCU --profile-from-start off --kernel-name-base mangled -k regex:bn_bwd_apply -s 20 -c 1 -f -o gpurun_out/${tag}_bn_bwd_apply python tools/step_profile.py > gpurun_out/${tag}_ncu4.log 2>&1; tail -2 gpurun_out/${tag}_ncu4.log
ls -la gpurun_out/${tag}_*.ncu-rep
