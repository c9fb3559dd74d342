#!/bin/sh
# compute-sanitizer over the kernel numerics checks.
#   tools/sanitize.sh memcheck  batchnorm pools_loss optimizer gemm_nk
#   tools/sanitize.sh racecheck conv_fprop          # shared-memory hazards in the epilogues
#   tools/sanitize.sh synccheck gemm_nk             # barrier / mbarrier misuse
# Every run is bounded by a timeout: the sanitizer slows kernels down 10-100x and a wedged GPU box
# costs a strike.  Output: gpurun_out/sanitize_<tool>.log (summary lines are what matter).
set -e
cd "$(dirname "$0")/.."
TOOL="${1:-memcheck}"; shift || true
CHECKS="${*:-batchnorm pools_loss optimizer}"
mkdir -p gpurun_out
LOG="gpurun_out/sanitize_${TOOL}.log"
timeout "${TFOS_SANITIZE_TIMEOUT:-300}" compute-sanitizer --tool "$TOOL" --error-exitcode 99 \
    --print-limit 20 python tools/gpu_check.py $CHECKS > "$LOG" 2>&1 || echo "exit code $?" >> "$LOG"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SYNCCHECK|CHECK .*FAIL|exit code|Error" "$LOG" | tail -20
