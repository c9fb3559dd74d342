"""Input pipelines for InputMode.TENSORFLOW programs: every worker reads its own files.

The reference leaves this to ``tf.data`` in user code (examples/mnist/keras/mnist_tf_ds.py:41-50:
``Dataset.list_files -> repeat -> shuffle -> interleave(TFRecordDataset) -> map(parse) -> batch``).
This module offers the same chain as plain generators over the native TFRecord codec
(csrc/tfrecord.cc): nothing here depends on TensorFlow.

  ds = data.TFRecordPipeline(pattern, epochs=3, shuffle_buffer=10000, seed=rank)
  for images, labels in ds.shard(world, rank).map(parse).batch(64): ...

``decode(spec)`` instead of ``map(fn)`` parses a whole batch of records natively into dense
arrays (tfrecord.decode_batch: no python object per value, ~13x the per-record path on
MNIST-shaped Examples):

  spec = {"image": ("int64", 784, "uint8"), "label": ("int64", 1)}
  for images, labels in ds.decode(spec).batch(64): ...       # uint8 [64, 784], int64 [64, 1]
"""
import glob
import os
import random

from .. import tfrecord


def list_files(pattern):
  """Sorted files matching a glob pattern, a directory (its ``part-*`` files) or a comma list;
  plain / ``file://`` paths and URIs of remote filesystems (``hdfs://`` ..., utils/fs.py - there
  the wildcard may only be in the file name)."""
  import fnmatch
  from . import fs
  out = []
  for piece in str(pattern).split(","):
    if fs.is_local(piece):
      piece = fs.local(piece)
      if os.path.isdir(piece):
        out.extend(sorted(glob.glob(os.path.join(piece, "part-*"))))
      else:
        out.extend(sorted(glob.glob(piece)))
    elif fs.isdir(piece):
      out.extend(fs.join(piece, n) for n in sorted(fs.listdir(piece)) if n.startswith("part-"))
    else:
      d, name = fs.dirname(piece), fs.basename(piece)
      if fs.isdir(d):
        out.extend(fs.join(d, n) for n in sorted(fs.listdir(d)) if fnmatch.fnmatch(n, name))
  return out


class TFRecordPipeline(object):
  """list_files -> (shard) -> repeat(epochs) + per-epoch file shuffle -> interleave of
  ``cycle_length`` open files -> shuffle buffer -> map -> batch.  Iterating yields either parsed
  examples or, after :meth:`batch`, tuples of stacked numpy columns."""

  def __init__(self, pattern, epochs=1, shuffle_buffer=0, cycle_length=4, seed=0, verify=True):
    self.files = list_files(pattern)
    if not self.files:
      raise IOError("no TFRecord files match {}".format(pattern))
    self.epochs, self.shuffle_buffer, self.cycle_length = epochs, shuffle_buffer, cycle_length
    self.seed, self.verify = seed, verify
    self._parse = None
    self._batch = None
    self._spec = None

  def shard(self, num_shards, index):
    """Keep every ``num_shards``-th file (AutoShardPolicy.FILE); falls back to sharding by
    record (AutoShardPolicy.DATA) when there are fewer files than shards."""
    if len(self.files) >= num_shards:
      self.files = self.files[index::num_shards]
      self._record_shard = None
    else:
      self._record_shard = (num_shards, index)
    return self

  def map(self, fn):
    self._parse = fn
    return self

  def decode(self, spec, threads=None):
    """Parse with the native batch decoder: ``spec`` = ``{feature: (kind, length[, dtype])}`` (see
    :func:`tfrecord.decode_batch`); batches then are tuples of ``[n, length]`` arrays in the
    order of ``spec``.  Needs :meth:`batch`; mutually exclusive with :meth:`map`."""
    self._spec, self._threads = dict(spec), threads
    return self

  def batch(self, n, drop_remainder=True):
    self._batch = (int(n), drop_remainder)
    return self

  # ------------------------------------------------------------------ stages
  def _records(self):
    rng = random.Random(self.seed)
    shard = getattr(self, "_record_shard", None)
    k = 0
    for _ in range(self.epochs):
      files = list(self.files)
      rng.shuffle(files)
      # interleave: round-robin over up to cycle_length open files
      pending = files[::-1]
      active = []
      while pending or active:
        while pending and len(active) < self.cycle_length:
          active.append(iter(tfrecord.read_records(pending.pop(), self.verify)))
        for it in list(active):
          try:
            rec = next(it)
          except StopIteration:
            active.remove(it)
            continue
          k += 1
          if shard is None or (k - 1) % shard[0] == shard[1]:
            yield rec

  def _shuffled(self, it):
    if self.shuffle_buffer <= 1:
      for x in it:
        yield x
      return
    rng = random.Random(self.seed + 1)
    buf = []
    for x in it:
      if len(buf) < self.shuffle_buffer:
        buf.append(x)
        continue
      j = rng.randrange(len(buf))
      buf[j], x = x, buf[j]
      yield x
    rng.shuffle(buf)
    for x in buf:
      yield x

  def __iter__(self):
    import numpy as np
    it = self._shuffled(self._records())
    if self._spec is not None:
      if self._parse is not None or self._batch is None:
        raise ValueError("decode(spec) replaces map(fn) and needs batch(n)")
      n, drop = self._batch
      names = list(self._spec)
      rows = []
      for rec in it:
        rows.append(rec)
        if len(rows) == n:
          cols = tfrecord.decode_batch(rows, self._spec, self._threads)
          yield tuple(cols[k] for k in names)
          rows = []
      if rows and not drop:
        cols = tfrecord.decode_batch(rows, self._spec, self._threads)
        yield tuple(cols[k] for k in names)
      return
    if self._parse is not None:
      it = (self._parse(r) for r in it)
    if self._batch is None:
      for x in it:
        yield x
      return
    n, drop = self._batch
    rows = []
    for x in it:
      rows.append(x)
      if len(rows) == n:
        yield tuple(np.stack(c) for c in zip(*rows))
        rows = []
    if rows and not drop:
      yield tuple(np.stack(c) for c in zip(*rows))


def decode_png_gray(data):
  """PNG bytes -> uint8 array [H, W] (tensorflow_datasets stores MNIST images as PNGs)."""
  import io
  import numpy as np
  from PIL import Image
  return np.asarray(Image.open(io.BytesIO(bytes(data))).convert("L"), dtype=np.uint8)
