"""Spark binding: real pyspark when importable, otherwise the bundled sparklite engine."""
try:
  import pyspark as _ps
  if getattr(_ps, "__sparklite__", False):
    raise ImportError("pyspark alias installed by sparklite")
  from pyspark import BarrierTaskContext, SparkConf, SparkContext, TaskContext  # noqa: F401
  from pyspark.sql import Row, SparkSession  # noqa: F401
  from pyspark.streaming import DStream  # noqa: F401
  from pyspark.ml import Estimator, Model  # noqa: F401
  from pyspark.ml.param import Param, Params, TypeConverters  # noqa: F401
  from pyspark import keyword_only  # noqa: F401
  BACKEND = "pyspark"
except ImportError:
  from .sparklite import (BarrierTaskContext, DStream, Row, SparkConf, SparkContext,  # noqa: F401
                          SparkSession, TaskContext)
  from .sparklite.ml import (Estimator, Model, Param, Params, TypeConverters,  # noqa: F401
                             keyword_only)
  BACKEND = "sparklite"
