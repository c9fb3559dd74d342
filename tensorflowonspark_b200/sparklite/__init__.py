"""sparklite: the pyspark-compatible local engine used when pyspark is absent.

``from tensorflowonspark_b200.sparklite import SparkContext, SparkConf`` mirrors
``from pyspark import ...``; ``install_as_pyspark()`` registers the package under
the name ``pyspark`` (and ``pyspark.sql`` / ``.ml`` / ``.streaming``) so that
unmodified reference driver scripts keep working.
"""
import sys
import types

from . import core, ml, sql, streaming
from .core import (RDD, BarrierTaskContext, SparkConf, SparkContext, SparkJobError,  # noqa: F401
                   TaskContext)
from .sql import DataFrame, Row, SparkSession  # noqa: F401
from .streaming import DStream, StreamingContext  # noqa: F401


def install_as_pyspark(force=False):
  """Expose sparklite under the ``pyspark`` module names if real pyspark is not importable."""
  if not force:
    try:
      import pyspark  # noqa: F401
      if not getattr(pyspark, "__sparklite__", False):
        return False
      return True
    except ImportError:
      pass
  pkg = types.ModuleType("pyspark")
  pkg.__sparklite__ = True
  pkg.__path__ = []
  for name in ("SparkContext", "SparkConf", "RDD", "TaskContext", "BarrierTaskContext"):
    setattr(pkg, name, getattr(core, name))
  pkg.sql = types.ModuleType("pyspark.sql")
  for name in ("SparkSession", "DataFrame", "Row", "Column"):
    setattr(pkg.sql, name, getattr(sql, name))
  pkg.sql.types = types.ModuleType("pyspark.sql.types")
  for name in ("DataType", "StringType", "BinaryType", "BooleanType", "IntegerType", "LongType",
               "FloatType", "DoubleType", "ArrayType", "StructField", "StructType"):
    setattr(pkg.sql.types, name, getattr(sql, name))
  pkg.sql.functions = types.ModuleType("pyspark.sql.functions")
  pkg.sql.functions.udf = sql.udf
  pkg.sql.functions.col = sql.col
  pkg.ml = types.ModuleType("pyspark.ml")
  for name in ("Estimator", "Model", "Transformer", "Pipeline", "PipelineModel"):
    setattr(pkg.ml, name, getattr(ml, name))
  pkg.ml.param = types.ModuleType("pyspark.ml.param")
  for name in ("Param", "Params", "TypeConverters"):
    setattr(pkg.ml.param, name, getattr(ml, name))
  pkg.ml.pipeline = types.ModuleType("pyspark.ml.pipeline")
  for name in ("Estimator", "Model", "Pipeline", "PipelineModel"):
    setattr(pkg.ml.pipeline, name, getattr(ml, name))
  pkg.keyword_only = ml.keyword_only
  pkg.streaming = types.ModuleType("pyspark.streaming")
  pkg.streaming.StreamingContext = streaming.StreamingContext
  pkg.streaming.DStream = streaming.DStream
  mods = {"pyspark": pkg, "pyspark.sql": pkg.sql, "pyspark.sql.types": pkg.sql.types,
          "pyspark.sql.functions": pkg.sql.functions, "pyspark.ml": pkg.ml,
          "pyspark.ml.param": pkg.ml.param, "pyspark.ml.pipeline": pkg.ml.pipeline,
          "pyspark.streaming": pkg.streaming}
  sys.modules.update(mods)
  return True
