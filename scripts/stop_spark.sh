#!/bin/bash
# Stop what scripts/start_spark.sh started (counterpart of the reference's scripts/stop_spark.sh).
: "${SPARK_HOME:?set SPARK_HOME}"
"$SPARK_HOME/sbin/stop-worker.sh"; "$SPARK_HOME/sbin/stop-master.sh"
