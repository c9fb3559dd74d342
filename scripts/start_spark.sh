#!/bin/bash
# Start a Spark Standalone master + one worker per GPU on this host (counterpart of the
# reference's scripts/start_spark.sh, which starts 2 CPU workers for its test-suite).
set -e
: "${SPARK_HOME:?set SPARK_HOME (scripts/install_spark.sh)}"
export SPARK_LOCAL_IP=${SPARK_LOCAL_IP:-127.0.0.1}
export MASTER=${MASTER:-spark://$(hostname):7077}
NGPU=$(python -c "import torch; print(max(1, torch.cuda.device_count()))" 2>/dev/null || echo 2)
export SPARK_WORKER_INSTANCES=${SPARK_WORKER_INSTANCES:-$NGPU}
export CORES_PER_WORKER=${CORES_PER_WORKER:-1}
# one GPU address per worker: Spark's resource API hands it to TFSparkNode._get_gpus
cat > "$SPARK_HOME/conf/tfos-gpu-discovery.sh" <<'EOD'
#!/bin/bash
idx=$(( ${SPARK_WORKER_INSTANCE_ID:-1} - 1 ))
echo "{\"name\": \"gpu\", \"addresses\": [\"$idx\"]}"
EOD
chmod +x "$SPARK_HOME/conf/tfos-gpu-discovery.sh"
"$SPARK_HOME/sbin/start-master.sh"
SPARK_WORKER_OPTS="-Dspark.worker.resource.gpu.amount=1 -Dspark.worker.resource.gpu.discoveryScript=$SPARK_HOME/conf/tfos-gpu-discovery.sh" \
  "$SPARK_HOME/sbin/start-worker.sh" -c "$CORES_PER_WORKER" -m 8G "$MASTER"
echo "master: $MASTER   workers: $SPARK_WORKER_INSTANCES"
echo "submit with: spark-submit --master $MASTER --conf spark.executor.resource.gpu.amount=1 --conf spark.task.resource.gpu.amount=1 <program>"
