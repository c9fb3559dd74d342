"""Minimal pyspark.sql work-alike: Row, data types, DataFrame, SparkSession.

Covers what TensorFlowOnSpark uses: ``createDataFrame`` / ``toDF``,
``select``, ``.rdd``, ``.dtypes``, ``.columns``, ``.schema`` with
``simpleString()``, ``collect/count/take/show``, ``withColumn`` + ``udf`` (used
by the reference examples/mnist/keras/mnist_pipeline.py:135-143), json output
(reference Inference.scala:79).
"""
import json
import os

from .core import RDD, SparkContext


# ----------------------------------------------------------------------- types
class DataType(object):
  name = "data"

  def simpleString(self):
    return self.name

  def __eq__(self, other):
    return type(self) is type(other) and self.__dict__ == other.__dict__

  def __ne__(self, other):
    return not self == other

  def __hash__(self):
    return hash(self.simpleString())

  def __repr__(self):
    return type(self).__name__ + "()"


def _simple(cls_name, s):
  return type(cls_name, (DataType,), {"name": s})


StringType = _simple("StringType", "string")
BinaryType = _simple("BinaryType", "binary")
BooleanType = _simple("BooleanType", "boolean")
IntegerType = _simple("IntegerType", "int")
LongType = _simple("LongType", "bigint")
FloatType = _simple("FloatType", "float")
DoubleType = _simple("DoubleType", "double")


class ArrayType(DataType):

  def __init__(self, elementType, containsNull=True):
    self.elementType = elementType
    self.containsNull = containsNull

  def simpleString(self):
    return "array<{}>".format(self.elementType.simpleString())

  def __eq__(self, other):
    return isinstance(other, ArrayType) and self.elementType == other.elementType

  def __hash__(self):
    return hash(self.simpleString())

  def __repr__(self):
    return "ArrayType({!r})".format(self.elementType)


class StructField(object):

  def __init__(self, name, dataType, nullable=True, metadata=None):
    self.name, self.dataType, self.nullable = name, dataType, nullable

  def simpleString(self):
    return "{}:{}".format(self.name, self.dataType.simpleString())

  def __eq__(self, other):
    return isinstance(other, StructField) and self.name == other.name and \
        self.dataType == other.dataType

  def __repr__(self):
    return "StructField({!r}, {!r})".format(self.name, self.dataType)


class StructType(DataType):

  def __init__(self, fields=None):
    self.fields = list(fields or [])

  def add(self, field, data_type=None, nullable=True):
    self.fields.append(field if isinstance(field, StructField) else StructField(field, data_type))
    return self

  @property
  def names(self):
    return [f.name for f in self.fields]

  def fieldNames(self):
    return self.names

  def simpleString(self):
    return "struct<{}>".format(",".join(f.simpleString() for f in self.fields))

  def __iter__(self):
    return iter(self.fields)

  def __len__(self):
    return len(self.fields)

  def __getitem__(self, k):
    if isinstance(k, str):
      return next(f for f in self.fields if f.name == k)
    return self.fields[k]

  def __eq__(self, other):
    return isinstance(other, StructType) and self.fields == other.fields

  def __hash__(self):
    return hash(self.simpleString())

  def __repr__(self):
    return "StructType({!r})".format(self.fields)


def _infer_type(v):
  if isinstance(v, bool):
    return BooleanType()
  if isinstance(v, int):
    return LongType()
  if isinstance(v, float):
    return DoubleType()
  if isinstance(v, str):
    return StringType()
  if isinstance(v, (bytes, bytearray)):
    return BinaryType()
  if isinstance(v, (list, tuple)):
    return ArrayType(_infer_type(v[0]) if len(v) else StringType())
  try:
    import numpy as np
    if isinstance(v, np.ndarray):
      return ArrayType(_infer_type(v.reshape(-1)[0].item()) if v.size else DoubleType())
    if isinstance(v, np.generic):
      return _infer_type(v.item())
  except ImportError:
    pass
  return StringType()


# ------------------------------------------------------------------------- Row
class Row(tuple):
  """Tuple with optional field names: Row(a=1, b=2) or Row(*values)."""

  def __new__(cls, *args, **kwargs):
    if args and kwargs:
      raise ValueError("Can not use both args and kwargs to create Row")
    if kwargs:
      row = tuple.__new__(cls, list(kwargs.values()))
      row.__fields__ = list(kwargs.keys())
      return row
    return tuple.__new__(cls, args)

  def asDict(self, recursive=False):
    if not hasattr(self, "__fields__"):
      raise TypeError("Cannot convert a Row class into dict")
    return dict(zip(self.__fields__, self))

  def __getattr__(self, item):
    if item.startswith("__"):
      raise AttributeError(item)
    try:
      return self[self.__fields__.index(item)]
    except (ValueError, AttributeError):
      raise AttributeError(item)

  def __getitem__(self, item):
    if isinstance(item, str):
      return tuple.__getitem__(self, self.__fields__.index(item))
    return tuple.__getitem__(self, item)

  def __call__(self, *args):
    row = Row(*args)
    row.__fields__ = list(self)
    return row

  def __reduce__(self):
    if hasattr(self, "__fields__"):
      return (_create_row, (self.__fields__, tuple(self)))
    return tuple.__reduce__(self)

  def __repr__(self):
    if hasattr(self, "__fields__"):
      return "Row({})".format(", ".join("{}={!r}".format(k, v)
                                        for k, v in zip(self.__fields__, tuple(self))))
    return "<Row({})>".format(", ".join(repr(f) for f in self))


def _create_row(fields, values):
  row = Row(*values)
  row.__fields__ = list(fields)
  return row


class _NamedRows(object):
  """Picklable stage turning plain tuples into named Rows."""

  def __init__(self, names):
    self.names = names

  def __call__(self, index, it):
    names = self.names
    for t in it:
      yield _create_row(names, tuple(t))


# ---------------------------------------------------------------------- column
class Column(object):
  """A deferred per-row expression: a column reference or a udf applied to columns."""

  def __init__(self, fn, name, dataType=None):
    self.fn, self.name, self.dataType = fn, name, dataType

  def alias(self, name):
    return Column(self.fn, name, self.dataType)


def col(name):
  return Column(lambda row, names: row[names.index(name)], name)


def udf(f=None, returnType=None):

  def wrap(fn):

    def apply(*cols):
      cols = [col(c) if isinstance(c, str) else c for c in cols]
      return Column(lambda row, names: fn(*[c.fn(row, names) for c in cols]),
                    getattr(fn, "__name__", "udf"), returnType)

    return apply

  if f is None or isinstance(f, DataType):
    returnType = f if isinstance(f, DataType) else returnType
    return wrap
  return wrap(f)


# ------------------------------------------------------------------- DataFrame
class DataFrame(object):

  def __init__(self, rdd, schema, session=None):
    self._rdd = rdd  # RDD of plain tuples, in schema order
    self.schema = schema
    self.sparkSession = session
    self.sql_ctx = session

  @property
  def columns(self):
    return self.schema.names

  @property
  def dtypes(self):
    return [(f.name, f.dataType.simpleString()) for f in self.schema.fields]

  @property
  def rdd(self):
    return self._rdd.mapPartitionsWithIndex(_NamedRows(self.schema.names))

  def select(self, *cols):
    if len(cols) == 1 and isinstance(cols[0], (list, tuple)):
      cols = tuple(cols[0])
    names = self.schema.names
    specs = [col(c) if isinstance(c, str) else c for c in cols]
    idx_only = all(isinstance(c, str) for c in cols)
    fields = []
    for c, s in zip(cols, specs):
      if isinstance(c, str):
        fields.append(self.schema[c])
      else:
        fields.append(StructField(s.name, s.dataType or StringType()))
    if idx_only:
      idx = [names.index(c) for c in cols]
      rdd = self._rdd.map(lambda t: tuple(t[i] for i in idx))
    else:
      rdd = self._rdd.map(lambda t: tuple(s.fn(t, names) for s in specs))
    return DataFrame(rdd, StructType(fields), self.sparkSession)

  def withColumn(self, name, column):
    names = self.schema.names
    fields = [f for f in self.schema.fields if f.name != name] + \
        [StructField(name, column.dataType or StringType())]
    keep = [i for i, n in enumerate(names) if n != name]
    rdd = self._rdd.map(lambda t: tuple(t[i] for i in keep) + (column.fn(t, names),))
    return DataFrame(rdd, StructType(fields), self.sparkSession)

  def drop(self, *cols):
    return self.select(*[c for c in self.columns if c not in cols])

  def collect(self):
    names = self.schema.names
    return [_create_row(names, tuple(t)) for t in self._rdd.collect()]

  def take(self, n):
    return self.collect()[:n]

  def head(self, n=None):
    rows = self.take(n or 1)
    return rows if n is not None else (rows[0] if rows else None)

  def first(self):
    return self.head()

  def count(self):
    return self._rdd.count()

  def limit(self, n):
    data = self._rdd.take(n)
    return DataFrame(self._rdd.ctx.parallelize(data, 1), self.schema, self.sparkSession)

  def cache(self):
    return self

  def repartition(self, n):
    return DataFrame(self._rdd.repartition(n), self.schema, self.sparkSession)

  def printSchema(self):
    print("root")
    for f in self.schema.fields:
      print(" |-- {}: {} (nullable = true)".format(f.name, f.dataType.simpleString()))

  def show(self, n=20, truncate=True):
    print("\t".join(self.columns))
    for r in self.take(n):
      print("\t".join(str(v)[:20] if truncate else str(v) for v in r))

  def toPandas(self):
    import pandas as pd
    return pd.DataFrame([tuple(r) for r in self._rdd.collect()], columns=self.columns)

  @property
  def write(self):
    return DataFrameWriter(self)


class DataFrameWriter(object):

  def __init__(self, df):
    self.df = df
    self._mode = "error"

  def mode(self, m):
    self._mode = m
    return self

  def json(self, path):
    from .core import _strip_scheme
    path = _strip_scheme(path)
    if os.path.exists(path) and self._mode != "overwrite":
      raise IOError("path {} already exists".format(path))
    os.makedirs(path, exist_ok=True)
    names = self.df.columns

    def enc(v):
      if isinstance(v, (bytes, bytearray)):
        import base64
        return base64.b64encode(bytes(v)).decode("ascii")
      if isinstance(v, (list, tuple)):
        return [enc(x) for x in v]
      return v

    def write(index, it):
      with open(os.path.join(path, "part-{:05d}.json".format(index)), "w") as fh:
        for t in it:
          fh.write(json.dumps({n: enc(v) for n, v in zip(names, t)}) + "\n")
      return iter([])

    self.df._rdd.mapPartitionsWithIndex(write).count()
    open(os.path.join(path, "_SUCCESS"), "w").close()


# ---------------------------------------------------------------- SparkSession
class _Builder(object):

  def __init__(self):
    self._conf = {}

  def master(self, m):
    self._conf["spark.master"] = m
    return self

  def appName(self, n):
    self._conf["spark.app.name"] = n
    return self

  def config(self, key=None, value=None, conf=None):
    if conf is not None:
      self._conf.update(dict(conf.getAll()))
    elif key is not None:
      self._conf[key] = str(value)
    return self

  def enableHiveSupport(self):
    return self

  def getOrCreate(self):
    from .core import SparkConf
    if SparkSession._active is not None and not SparkSession._active.sparkContext._stopped:
      return SparkSession._active
    conf = SparkConf().setAll(self._conf.items())
    return SparkSession(SparkContext.getOrCreate(conf))


class _BuilderDescriptor(object):

  def __get__(self, obj, objtype=None):
    return _Builder()


class SparkSession(object):
  _active = None
  builder = _BuilderDescriptor()

  def __init__(self, sparkContext):
    self.sparkContext = sparkContext
    self._sc = sparkContext
    SparkSession._active = self

  def createDataFrame(self, data, schema=None, samplingRatio=None):
    sc = self.sparkContext
    if isinstance(data, RDD):
      rdd = data
      first = rdd.first()
    else:
      try:
        import pandas as pd
        if isinstance(data, pd.DataFrame):
          schema = schema or list(data.columns)
          data = [tuple(r) for r in data.itertuples(index=False, name=None)]
      except ImportError:
        pass
      data = list(data)
      first = data[0] if data else ()
      rdd = sc.parallelize(data)
    if isinstance(first, dict):
      keys = sorted(first.keys())
      rdd = rdd.map(lambda d: tuple(d[k] for k in keys))
      schema = schema or keys
      first = tuple(first[k] for k in keys)
    elif isinstance(first, Row) and hasattr(first, "__fields__") and schema is None:
      schema = list(first.__fields__)
    if not isinstance(first, (tuple, list)):
      rdd = rdd.map(lambda v: (v,))
      first = (first,)
    if isinstance(schema, StructType):
      st = schema
    else:
      names = list(schema) if schema else ["_{}".format(i + 1) for i in range(len(first))]
      st = StructType([StructField(n, _infer_type(v)) for n, v in zip(names, first)])
    return DataFrame(rdd.map(lambda t: tuple(t)), st, self)

  @property
  def read(self):
    return DataFrameReader(self)

  def stop(self):
    self.sparkContext.stop()
    SparkSession._active = None


class DataFrameReader(object):

  def __init__(self, session):
    self.session = session

  def json(self, path):
    from .core import _list_files
    rows = []
    for f in _list_files(path):
      with open(f) as fh:
        rows.extend(json.loads(line) for line in fh if line.strip())
    keys = sorted(rows[0].keys()) if rows else []
    return self.session.createDataFrame([tuple(r.get(k) for k in keys) for r in rows], keys)

  def csv(self, path, header=False, inferSchema=False):
    from .core import _list_files
    rows = []
    for f in _list_files(path):
      with open(f) as fh:
        rows.extend(tuple(line.rstrip("\n").split(",")) for line in fh if line.strip())
    names = None
    if header and rows:
      names, rows = list(rows[0]), rows[1:]
    if inferSchema:

      def conv(v):
        for t in (int, float):
          try:
            return t(v)
          except ValueError:
            pass
        return v

      rows = [tuple(conv(v) for v in r) for r in rows]
    return self.session.createDataFrame(rows, names)
