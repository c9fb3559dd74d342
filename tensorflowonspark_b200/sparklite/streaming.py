"""Minimal pyspark.streaming work-alike: StreamingContext / DStream.

The reference feeds a DStream through ``TFCluster.train`` with ``foreachRDD``
and stops it via ``awaitTerminationOrTimeout`` polling (tensorflowonspark/
TFCluster.py:83-89,147-153; examples/mnist/estimator/mnist_spark_streaming.py:
115-142).  Sources: ``queueStream`` and ``textFileStream`` (new files in a
directory are one micro-batch per interval).
"""
import os
import threading
import time

from .core import _list_files


class DStream(object):

  def __init__(self, ssc):
    self._ssc = ssc
    self._actions = []
    self._transforms = []

  def _derive(self, f):
    d = DStream(self._ssc)
    d._parent, d._fn = self, f
    self._ssc._streams.append(d)
    return d

  def map(self, f):
    return self._derive(lambda rdd: rdd.map(f))

  def mapPartitions(self, f):
    return self._derive(lambda rdd: rdd.mapPartitions(f))

  def transform(self, f):
    return self._derive(f)

  def foreachRDD(self, func):
    self._actions.append(func)

  def _rdd_for(self, batch):
    if hasattr(self, "_parent"):
      r = self._parent._rdd_for(batch)
      return None if r is None else self._fn(r)
    return batch.get(id(self))

  def _fire(self, batch):
    rdd = self._rdd_for(batch)
    if rdd is None:
      return
    for a in self._actions:
      try:
        a(rdd)
      except TypeError:
        a(time.time(), rdd)


class _QueueSource(DStream):

  def __init__(self, ssc, rdds, oneAtATime, default):
    super(_QueueSource, self).__init__(ssc)
    self._queue = [r if hasattr(r, "_partitions") else ssc.sparkContext.parallelize(r)
                   for r in rdds]
    self._one, self._default = oneAtATime, default

  def _next(self):
    if self._queue:
      if self._one:
        return self._queue.pop(0)
      rdds, self._queue = self._queue, []
      return self._ssc.sparkContext.union(rdds)
    return self._default


class _FileSource(DStream):

  def __init__(self, ssc, directory):
    super(_FileSource, self).__init__(ssc)
    self._dir = directory
    self._seen = set(_list_files(directory)) if os.path.isdir(directory) else set()

  def _next(self):
    files = [f for f in _list_files(self._dir) if f not in self._seen]
    if not files:
      return None
    self._seen.update(files)
    return self._ssc.sparkContext.textFile(",".join(files))


class StreamingContext(object):

  def __init__(self, sparkContext, batchDuration=1):
    self.sparkContext = sparkContext
    self._sc = sparkContext
    self._interval = float(batchDuration)
    self._streams = []
    self._sources = []
    self._thread = None
    self._stop = threading.Event()
    self._terminated = threading.Event()
    self._error = None

  def queueStream(self, rdds, oneAtATime=True, default=None):
    s = _QueueSource(self, list(rdds), oneAtATime, default)
    self._sources.append(s)
    self._streams.append(s)
    return s

  def textFileStream(self, directory):
    s = _FileSource(self, directory)
    self._sources.append(s)
    self._streams.append(s)
    return s

  def _loop(self):
    try:
      while not self._stop.is_set():
        t0 = time.time()
        batch = {}
        for s in self._sources:
          r = s._next()
          if r is not None:
            batch[id(s)] = r
        if batch:
          for d in list(self._streams):
            if d._actions:
              d._fire(batch)
        self._stop.wait(max(0.0, self._interval - (time.time() - t0)))
    except Exception as e:  # surfaced by awaitTermination*
      self._error = e
    finally:
      self._terminated.set()

  def start(self):
    self._thread = threading.Thread(target=self._loop, name="sparklite-streaming", daemon=True)
    self._thread.start()

  def awaitTermination(self, timeout=None):
    self._terminated.wait(timeout)
    if self._error is not None:
      raise self._error

  def awaitTerminationOrTimeout(self, timeout):
    done = self._terminated.wait(timeout)
    if self._error is not None:
      raise self._error
    return done

  def stop(self, stopSparkContext=True, stopGraceFully=False):
    self._stop.set()
    if self._thread is not None:
      self._thread.join(self._interval + 30 if stopGraceFully else 5)
    self._terminated.set()
    if stopSparkContext:
      self.sparkContext.stop()
