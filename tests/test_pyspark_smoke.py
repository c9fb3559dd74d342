"""Runs only where a real pyspark (+ JVM) is importable (reference tests/README.md:10 runs its
suite on a Spark Standalone cluster): the same TFCluster plumbing test on ``local[2]`` of real
Spark, proving the `_spark.py` binding is not sparklite-only.  Skipped on this image."""
import importlib.util
import shutil

import pytest

_real = importlib.util.find_spec("pyspark") is not None and shutil.which("java") is not None
if _real:
  import pyspark
  _real = not getattr(pyspark, "__sparklite__", False)

pytestmark = pytest.mark.skipif(not _real, reason="real pyspark + JVM not installed")


def _fn(args, ctx):
  feed = ctx.get_data_feed()
  total = 0
  while not feed.should_stop():
    batch = feed.next_batch(10)
    total += sum(batch)
  with open(args["out"] + str(ctx.task_index), "w") as f:
    f.write(str(total))


def test_tfcluster_inputmode_spark_on_real_pyspark(tmp_path):
  from pyspark import SparkConf, SparkContext
  from tensorflowonspark_b200 import TFCluster, _spark
  assert _spark.BACKEND == "pyspark"
  sc = SparkContext(conf=SparkConf().setMaster("local-cluster[2,1,1024]").setAppName("tfos-smoke"))
  try:
    cluster = TFCluster.run(sc, _fn, {"out": str(tmp_path / "sum")}, 2, 0,
                            input_mode=TFCluster.InputMode.SPARK)
    cluster.train(sc.parallelize(range(1000), 4))
    cluster.shutdown()
    got = sum(int(open(str(tmp_path / "sum") + str(i)).read()) for i in range(2))
    assert got == sum(range(1000))
  finally:
    sc.stop()
