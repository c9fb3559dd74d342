"""pyspark.ml work-alikes used by the ML pipeline layer: Param / Params /
TypeConverters / Estimator / Model / Pipeline / keyword_only.

The reference builds its 19 ``Has*`` mix-ins on these (tensorflowonspark/
pipeline.py:52-296) and relies on ``_set``, ``_setDefault``, ``getOrDefault``,
``_copyValues`` and ``extractParamMap`` semantics, which are reproduced here.
"""
import copy
import functools


class Param(object):

  def __init__(self, parent, name, doc, typeConverter=None):
    self.parent = parent if isinstance(parent, str) else getattr(parent, "uid", str(parent))
    self.name = str(name)
    self.doc = str(doc)
    self.typeConverter = typeConverter or TypeConverters.identity

  def __hash__(self):
    return hash(self.name)

  def __eq__(self, other):
    return isinstance(other, Param) and self.name == other.name

  def __repr__(self):
    return "Param(name={!r}, doc={!r})".format(self.name, self.doc)


class TypeConverters(object):

  @staticmethod
  def identity(v):
    return v

  @staticmethod
  def toInt(v):
    if isinstance(v, bool) or int(v) != v:
      raise TypeError("Could not convert {!r} to int".format(v))
    return int(v)

  @staticmethod
  def toFloat(v):
    return float(v)

  @staticmethod
  def toString(v):
    if isinstance(v, (str, bytes)):
      return v if isinstance(v, str) else v.decode("utf-8")
    raise TypeError("Could not convert {!r} to string".format(v))

  @staticmethod
  def toBoolean(v):
    if isinstance(v, bool):
      return v
    raise TypeError("Boolean Param requires value of type bool. Found {}".format(type(v)))

  @staticmethod
  def toList(v):
    if isinstance(v, (list, tuple)):
      return list(v)
    raise TypeError("Could not convert {!r} to list".format(v))

  @staticmethod
  def toListString(v):
    return [TypeConverters.toString(x) for x in TypeConverters.toList(v)]

  @staticmethod
  def toListInt(v):
    return [TypeConverters.toInt(x) for x in TypeConverters.toList(v)]

  @staticmethod
  def toListFloat(v):
    return [float(x) for x in TypeConverters.toList(v)]


_uid_counter = [0]


class Params(object):
  """Holder of Params declared as class attributes (instances get their own copies)."""

  def __init__(self):
    _uid_counter[0] += 1
    self.uid = "{}_{:04x}".format(type(self).__name__, _uid_counter[0])
    self._paramMap = {}
    self._defaultParamMap = {}
    for klass in type(self).__mro__:
      for name, value in vars(klass).items():
        if isinstance(value, Param) and name not in self.__dict__:
          setattr(self, name, Param(self, value.name, value.doc, value.typeConverter))

  @property
  def params(self):
    return sorted((v for v in self.__dict__.values() if isinstance(v, Param)),
                  key=lambda p: p.name)

  def _resolve(self, param):
    return self.getParam(param) if isinstance(param, str) else self.getParam(param.name)

  def hasParam(self, name):
    return isinstance(getattr(self, name, None), Param)

  def getParam(self, name):
    p = getattr(self, name, None)
    if not isinstance(p, Param):
      raise ValueError("Cannot find param with name {}.".format(name))
    return p

  def isSet(self, param):
    return self._resolve(param) in self._paramMap

  def hasDefault(self, param):
    return self._resolve(param) in self._defaultParamMap

  def isDefined(self, param):
    return self.isSet(param) or self.hasDefault(param)

  def getOrDefault(self, param):
    p = self._resolve(param)
    if p in self._paramMap:
      return self._paramMap[p]
    return self._defaultParamMap[p]

  def extractParamMap(self, extra=None):
    m = dict(self._defaultParamMap)
    m.update(self._paramMap)
    if extra:
      m.update(extra)
    return m

  def explainParams(self):
    return "\n".join("{}: {} (current: {})".format(
        p.name, p.doc, self.getOrDefault(p) if self.isDefined(p) else "undefined")
        for p in self.params)

  def _set(self, **kwargs):
    for name, value in kwargs.items():
      p = self.getParam(name)
      if value is not None:
        try:
          value = p.typeConverter(value)
        except (TypeError, ValueError) as e:
          raise TypeError('Invalid param value given for param "{}". {}'.format(p.name, e))
      self._paramMap[p] = value
    return self

  def set(self, param, value):
    return self._set(**{self._resolve(param).name: value})

  def _setDefault(self, **kwargs):
    for name, value in kwargs.items():
      p = self.getParam(name)
      if value is not None:
        value = p.typeConverter(value)
      self._defaultParamMap[p] = value
    return self

  def clear(self, param):
    self._paramMap.pop(self._resolve(param), None)

  def copy(self, extra=None):
    that = copy.copy(self)
    that._paramMap = dict(self._paramMap)
    that._defaultParamMap = dict(self._defaultParamMap)
    return self._copyValues(that, extra)

  def _copyValues(self, to, extra=None):
    m = self.extractParamMap(extra)
    for p, v in m.items():
      if to.hasParam(p.name):
        if p in self._defaultParamMap and p not in self._paramMap and not (extra and p in extra):
          to._defaultParamMap[to.getParam(p.name)] = v
        else:
          to._paramMap[to.getParam(p.name)] = v
    return to


def keyword_only(func):
  """Decorator storing the keyword arguments of the call in ``self._input_kwargs``."""

  @functools.wraps(func)
  def wrapper(self, *args, **kwargs):
    if args:
      raise TypeError("Method {} forces keyword arguments.".format(func.__name__))
    self._input_kwargs = kwargs
    return func(self, **kwargs)

  return wrapper


class Transformer(Params):

  def transform(self, dataset, params=None):
    if params:
      return self.copy(params)._transform(dataset)
    return self._transform(dataset)

  def _transform(self, dataset):
    raise NotImplementedError()


class Model(Transformer):
  pass


class Estimator(Params):

  def fit(self, dataset, params=None):
    if params:
      return self.copy(params)._fit(dataset)
    return self._fit(dataset)

  def _fit(self, dataset):
    raise NotImplementedError()


class PipelineModel(Model):

  def __init__(self, stages):
    super(PipelineModel, self).__init__()
    self.stages = stages

  def _transform(self, dataset):
    for s in self.stages:
      dataset = s.transform(dataset)
    return dataset


class Pipeline(Estimator):

  def __init__(self, stages=None):
    super(Pipeline, self).__init__()
    self._stages = list(stages or [])

  def setStages(self, stages):
    self._stages = list(stages)
    return self

  def getStages(self):
    return self._stages

  def _fit(self, dataset):
    fitted = []
    for s in self._stages:
      if isinstance(s, Estimator):
        m = s.fit(dataset)
        fitted.append(m)
        dataset = m.transform(dataset)
      else:
        fitted.append(s)
        dataset = s.transform(dataset)
    return PipelineModel(fitted)
