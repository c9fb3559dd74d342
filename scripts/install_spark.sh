#!/bin/bash
# Optional: install a real Apache Spark next to this framework (counterpart of the reference's
# scripts/install_spark.sh).  The framework runs without it (bundled sparklite engine); with
# pyspark importable `tensorflowonspark_b200._spark` binds to it instead.
#   SPARK_TGZ=/path/to/spark-3.x-bin-hadoop3.tgz scripts/install_spark.sh     (offline)
#   scripts/install_spark.sh                                                  (downloads)
set -e
: "${SPARK_VERSION:=3.5.1}" "${HADOOP_VERSION:=3}" "${SPARK_HOME:=/opt/spark}"
command -v java >/dev/null || { echo "a JDK (8/11/17) is required: install one and set JAVA_HOME" >&2; exit 1; }
if [ -z "$SPARK_TGZ" ]; then
  SPARK_TGZ=spark-${SPARK_VERSION}-bin-hadoop${HADOOP_VERSION}.tgz
  curl -fLO "https://archive.apache.org/dist/spark/spark-${SPARK_VERSION}/${SPARK_TGZ}"
fi
mkdir -p "$SPARK_HOME"
tar -xf "$SPARK_TGZ" -C "$SPARK_HOME" --strip-components=1
python -m pip install "pyspark==${SPARK_VERSION}" || echo "pip install pyspark failed (offline?): add \$SPARK_HOME/python to PYTHONPATH instead"
echo "Spark installed in $SPARK_HOME"
