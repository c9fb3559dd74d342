"""DataFrame <-> TFRecord conversion without TensorFlow or a JVM.

Python parity with tensorflowonspark/dfutil.py:18-212 (``saveAsTFRecords``, ``loadTFRecords``,
``toTFExample``, ``fromTFExample``, ``infer_schema``, ``isLoadedDF``) and with the richer Scala
side (DFUtil.scala:21-259: schema *hints*; SimpleTypeParser.scala:27-64: the
``struct<name:type,...>`` grammar, here :func:`parse_schema`).  Records are framed and the
``tf.train.Example`` protos are encoded by the native codec in csrc/tfrecord.cc.
"""
import logging
import re

from . import tfrecord
from ._spark import BACKEND

logger = logging.getLogger(__name__)

if BACKEND == "pyspark":  # pragma: no cover - exercised only where pyspark is installed
  from pyspark.sql import Row
  from pyspark.sql.types import (ArrayType, BinaryType, BooleanType, DoubleType, FloatType,
                                 IntegerType, LongType, StringType, StructField, StructType)
else:
  from .sparklite.sql import (ArrayType, BinaryType, BooleanType, DoubleType, FloatType,
                              IntegerType, LongType, Row, StringType, StructField, StructType)

#: DataFrames produced by :func:`loadTFRecords`, keyed by object id -> source directory
loadedDF = {}

_INT_TYPES = ("tinyint", "smallint", "int", "bigint", "long", "boolean")
_FLOAT_TYPES = ("float", "double")
_BYTES_TYPES = ("string", "binary")


def isLoadedDF(df):
  """True when ``df`` is exactly a DataFrame returned by :func:`loadTFRecords` (identity, not
  equality: any transformation yields a new DataFrame that is no longer "loaded")."""
  return id(df) in loadedDF and loadedDF[id(df)][0] is df


TFRECORD_INPUT_FORMAT = "org.tensorflow.hadoop.io.TFRecordFileInputFormat"
TFRECORD_OUTPUT_FORMAT = "org.tensorflow.hadoop.io.TFRecordFileOutputFormat"
_KEY_CLASS = "org.apache.hadoop.io.BytesWritable"
_VALUE_CLASS = "org.apache.hadoop.io.NullWritable"


def saveAsTFRecords(df, output_dir):
  """Write a DataFrame as TFRecord files (one ``part-r-NNNNN`` per partition) of Examples.

  Goes through ``RDD.saveAsNewAPIHadoopFile`` with the tensorflow-hadoop output format, exactly
  like the reference (dfutil.py:29-41): under real pyspark the Hadoop layer resolves
  ``hdfs://`` / ``viewfs://`` paths and writes on the executors' shared filesystem; under
  sparklite the same call lands in the native TFRecord writer on the local filesystem."""
  tf_rdd = df.rdd.mapPartitions(toTFExample(df.dtypes))
  tf_rdd.saveAsNewAPIHadoopFile(output_dir, TFRECORD_OUTPUT_FORMAT, keyClass=_KEY_CLASS,
                                valueClass=_VALUE_CLASS)


def loadTFRecords(sc, input_dir, binary_features=[], schema_hint=None):
  """Load TFRecord files of Examples as a DataFrame (``sc.newAPIHadoopFile`` with the
  tensorflow-hadoop input format, reference dfutil.py:44-80 - any filesystem the bound Spark
  backend can read).

  The schema is inferred from the first record (a feature holding one value becomes a scalar
  column, several values an array column); ``binary_features`` lists bytes features that must
  stay binary instead of being decoded as UTF-8 strings; ``schema_hint`` (a StructType or a
  ``struct<...>`` string) overrides the inference for the named columns - the Scala-side
  capability (DFUtil.scala:35-55).
  """
  tfr_rdd = sc.newAPIHadoopFile(input_dir, TFRECORD_INPUT_FORMAT, keyClass=_KEY_CLASS,
                                valueClass=_VALUE_CLASS)
  first = tfr_rdd.take(1)
  if not first:
    raise IOError("no TFRecord records under {}".format(input_dir))
  hint = parse_schema(schema_hint) if isinstance(schema_hint, str) else schema_hint
  schema = infer_schema(bytes(first[0][0]), binary_features, hint)
  names = schema.names
  rows = tfr_rdd.mapPartitions(lambda it: fromTFExample(it, binary_features, schema))
  df = rows.map(lambda r: tuple(r[n] for n in names)).toDF(schema)
  loadedDF[id(df)] = (df, input_dir)
  return df


def _kind_of(dtype):
  base = dtype
  m = re.match(r"array<(.*)>$", dtype)
  if m:
    base = m.group(1)
  if base in _INT_TYPES:
    return "int64", bool(m)
  if base in _FLOAT_TYPES:
    return "float", bool(m)
  if base in _BYTES_TYPES:
    return "bytes", bool(m)
  raise Exception("Unsupported dtype for TFRecord export: {}".format(dtype))


def toTFExample(dtypes):
  """``mapPartitions`` function turning Rows into ``(serialized Example, None)`` pairs.

  ``dtypes`` is ``DataFrame.dtypes``: [(column, simple type string)].
  """
  plan = [(name,) + _kind_of(dt) for name, dt in dtypes]

  def _convert(iterator):
    out = []
    for row in iterator:
      feats = {}
      for i, (name, kind, is_array) in enumerate(plan):
        v = row[i]
        vals = list(v) if is_array else [v]
        if kind == "bytes":
          vals = [x.encode("utf-8") if isinstance(x, str) else bytes(x) for x in vals]
        elif kind == "int64":
          vals = [int(x) for x in vals]
        else:
          vals = [float(x) for x in vals]
        feats[name] = (kind, vals)
      out.append((bytearray(tfrecord.encode_example(feats)), None))
    return out

  return _convert


def infer_schema(example, binary_features=[], hint=None):
  """StructType for a serialized Example: int64 -> long, float -> double, bytes -> string (or
  binary when listed), more than one value -> array.  Columns are sorted by name."""
  feats = tfrecord.decode_example(bytes(example))
  hinted = {f.name: f for f in hint.fields} if hint is not None else {}
  fields = []
  for name in sorted(feats):
    if name in hinted:
      fields.append(hinted[name])
      continue
    kind, vals = feats[name]
    if kind == "int64":
      base = LongType()
    elif kind == "float":
      base = DoubleType()
    else:
      base = BinaryType() if name in binary_features else StringType()
    fields.append(StructField(name, ArrayType(base) if len(vals) > 1 else base))
  return StructType(fields)


def _cast(base, v, binary):
  if isinstance(base, (IntegerType, LongType)):
    return int(v)
  if isinstance(base, BooleanType):
    return bool(v)
  if isinstance(base, (FloatType, DoubleType)):
    return float(v)
  if isinstance(base, BinaryType) or binary:
    return bytearray(v)
  return v.decode("utf-8") if isinstance(v, (bytes, bytearray)) else v


def fromTFExample(iter, binary_features=[], schema=None):  # noqa: A002 (reference arg name)
  """``mapPartitions`` function turning serialized Examples into Rows (fields sorted by name)."""
  out = []
  by_name = {f.name: f.dataType for f in schema.fields} if schema is not None else {}
  for rec in iter:
    if isinstance(rec, tuple):
      rec = rec[0]
    feats = tfrecord.decode_example(bytes(rec))
    names = sorted(feats)
    values = []
    for name in names:
      kind, vals = feats[name]
      binary = name in binary_features
      dt = by_name.get(name)
      if dt is not None:
        if isinstance(dt, ArrayType):
          values.append([_cast(dt.elementType, v, binary) for v in vals])
        else:
          values.append(_cast(dt, vals[0], binary) if vals else None)
        continue
      if kind == "bytes":
        conv = [bytearray(v) if binary else v.decode("utf-8") for v in vals]
      else:
        conv = list(vals)
      values.append(conv[0] if len(conv) == 1 else conv)
    row = Row(*values)
    row.__fields__ = names
    out.append(row)
  return out


# -------------------------------------------------------- struct<...> grammar
_BASE = {
    "binary": BinaryType, "boolean": BooleanType, "int": IntegerType, "long": LongType,
    "bigint": LongType, "float": FloatType, "double": DoubleType, "string": StringType,
}


def parse_schema(text):
  """Parse ``struct<name:type,...>`` with base types binary / boolean / int / long / bigint /
  float / double / string and ``array<base>`` (the grammar of SimpleTypeParser.scala:34-64)."""
  s = text.strip()
  m = re.match(r"^struct<(.*)>$", s, re.S)
  if not m:
    raise ValueError("schema must look like struct<name:type,...>: {}".format(text))
  body, fields, depth, cur = m.group(1), [], 0, ""
  for ch in body:
    if ch == "<":
      depth += 1
    elif ch == ">":
      depth -= 1
    if ch == "," and depth == 0:
      fields.append(cur)
      cur = ""
    else:
      cur += ch
  if cur.strip():
    fields.append(cur)
  out = []
  for f in fields:
    if ":" not in f:
      raise ValueError("bad field '{}' in schema {}".format(f, text))
    name, typ = f.split(":", 1)
    name, typ = name.strip(), typ.strip()
    am = re.match(r"^array<(\w+)>$", typ)
    base = am.group(1) if am else typ
    if base not in _BASE:
      raise ValueError("unsupported type '{}' in schema {}".format(typ, text))
    dt = _BASE[base]()
    out.append(StructField(name, ArrayType(dt) if am else dt))
  return StructType(out)
