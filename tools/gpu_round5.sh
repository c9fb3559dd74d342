#!/bin/bash
# per-kernel step breakdown with and without the fused BN reduction (ncu launch list; serialised,
# cold-cache: compare shares)
tag=${1:-r}
mkdir -p gpurun_out
for f in 1 0; do
TFOS_BN_FUSED_REDUCE=$f timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_step_fr$f.csv python tools/step_profile.py > gpurun_out/${tag}_step_fr$f.log 2>&1
python tools/ncu_summary.py gpurun_out/${tag}_step_fr$f.csv 60 > gpurun_out/${tag}_step_fr$f.txt 2>&1
head -32 gpurun_out/${tag}_step_fr$f.txt
done
