#!/bin/sh
# Run a TFCluster program on all GPUs of this box (counterpart of scripts/start_spark.sh +
# spark-submit in the reference): sparklite forks one executor per GPU; with a real Spark
# installation use spark-submit and the same program instead.
#   scripts/launch_local_cluster.sh examples/resnet/resnet_spark.py --model resnet50 --train_steps 100
set -e
cd "$(dirname "$0")/.."
NGPU=$(python -c "import torch; print(max(1, torch.cuda.device_count()))")
PROG="$1"; shift
exec python "$PROG" --cluster_size "$NGPU" "$@"
