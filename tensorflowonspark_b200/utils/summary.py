"""TensorBoard event files without TensorFlow: what the TensorBoard a node launches has to read.

The reference starts a ``tensorboard --logdir`` subprocess on worker:0 / chief:0
(tensorflowonspark/TFSparkNode.py:293-329) and its examples fill that directory through the Keras
``TensorBoard`` callback or ``tf.estimator``'s summary hooks (examples/mnist/keras/mnist_tf.py:62,
examples/mnist/estimator/mnist_spark.py:100).  With neither available, this module writes the same
``events.out.tfevents.*`` format directly: TFRecord framing (native masked-CRC32C writer,
csrc/tfrecord.cc) around hand-encoded ``Event`` / ``Summary`` protobuf messages - scalars and
histograms, which is what those callbacks emit for a training loop.

  w = SummaryWriter(ctx.absolute_path(args.model_dir))
  w.add_scalar("loss", loss, step)
  w.add_histogram("fc/weights", weights, step)
  w.close()

``read_events`` is the inverse (used by the tests and by ``tools``-side post-processing).
"""
import os
import socket
import struct
import threading
import time

from .. import tfrecord

_FILE_VERSION = b"brain.Event:2"
_counter = [0]
_lock = threading.Lock()


# ------------------------------------------------------------------ protobuf wire helpers
def _varint(v):
  v &= 0xffffffffffffffff
  out = bytearray()
  while v >= 0x80:
    out.append((v & 0x7f) | 0x80)
    v >>= 7
  out.append(v)
  return bytes(out)


def _key(field, wire):
  return _varint((field << 3) | wire)


def _ld(field, payload):
  return _key(field, 2) + _varint(len(payload)) + payload


def _f64(field, v):
  return _key(field, 1) + struct.pack("<d", float(v))


def _f32(field, v):
  return _key(field, 5) + struct.pack("<f", float(v))


def _packed_f64(field, vals):
  return _ld(field, b"".join(struct.pack("<d", float(v)) for v in vals)) if len(vals) else b""


def _event(wall_time, step=None, file_version=None, summary=None):
  # Event: wall_time = 1 (double), step = 2 (int64), file_version = 3 (string), summary = 5
  out = _f64(1, wall_time)
  if step is not None:
    out += _key(2, 0) + _varint(int(step))
  if file_version is not None:
    out += _ld(3, file_version)
  if summary is not None:
    out += _ld(5, summary)
  return out


def _scalar_value(tag, value):
  # Summary.Value: tag = 1 (string), simple_value = 2 (float)
  return _ld(1, _ld(1, tag.encode("utf-8")) + _f32(2, value))


def _default_edges():
  # TensorFlow's exponential bucket layout: +-1e-12 * 1.1^k up to 1e20, symmetric around zero
  pos, v = [], 1e-12
  while v < 1e20:
    pos.append(v)
    v *= 1.1
  return [-x for x in reversed(pos)] + [0.0] + pos + [float("inf")]


_EDGES = None


def _histogram_value(tag, values):
  # HistogramProto: min = 1, max = 2, num = 3, sum = 4, sum_squares = 5 (double),
  # bucket_limit = 6, bucket = 7 (packed double); only the occupied range of buckets is written
  import numpy as np
  global _EDGES
  if _EDGES is None:
    _EDGES = np.asarray(_default_edges())
  v = np.asarray(values, dtype=np.float64).ravel()
  v = v[np.isfinite(v)]
  if v.size == 0:
    v = np.zeros(1)
  idx = np.searchsorted(_EDGES, v, side="left")      # first edge >= value: limits are inclusive
  counts = np.bincount(idx, minlength=len(_EDGES))
  lo, hi = int(idx.min()), int(idx.max())
  lo = max(0, lo - 1)                                 # one empty bucket on the left closes the range
  histo = (_f64(1, v.min()) + _f64(2, v.max()) + _f64(3, v.size) + _f64(4, v.sum()) +
           _f64(5, float((v * v).sum())) + _packed_f64(6, _EDGES[lo:hi + 1]) +
           _packed_f64(7, counts[lo:hi + 1]))
  return _ld(1, _ld(1, tag.encode("utf-8")) + _ld(5, histo))


# ------------------------------------------------------------------ writer
class SummaryWriter(object):
  """Appends scalar / histogram summaries to one ``events.out.tfevents.*`` file in ``logdir``.

  Events are buffered and handed to the record writer every ``flush_secs`` seconds or
  ``max_queue`` events, whichever comes first, and on ``flush`` / ``close`` (a training loop calls
  ``add_scalar`` per step; the file is touched a few times a minute).  Thread-safe."""

  def __init__(self, logdir, filename_suffix="", flush_secs=10.0, max_queue=64):
    from . import fs
    with _lock:
      _counter[0] += 1
      uid = _counter[0]
    name = "events.out.tfevents.{:010d}.{}.{}.{}{}".format(
        int(time.time()), socket.gethostname(), os.getpid(), uid, filename_suffix)
    # a log directory on a remote filesystem (hdfs://, s3:// ... - utils/fs.py): records are
    # appended to a local spool file which replaces the remote copy at every flush (event files
    # of a training run are small, and TensorBoard re-reads them periodically anyway)
    self.remote = None
    if fs.is_local(logdir):
      logdir = fs.local(logdir)
      os.makedirs(logdir, exist_ok=True)
    else:
      import tempfile
      fs.makedirs(logdir)
      self.remote = fs.join(str(logdir), name)
      logdir = tempfile.mkdtemp(prefix="tfos-events-")
    self.path = os.path.join(logdir, name)
    self.flush_secs, self.max_queue = flush_secs, max_queue
    self._pending, self._last = [], time.time()
    self._mutex = threading.Lock()
    self._closed = False
    tfrecord.write_records(self.path, [_event(time.time(), file_version=_FILE_VERSION)])
    self._upload()

  def _add(self, step, value, wall_time):
    with self._mutex:
      if self._closed:
        raise ValueError("SummaryWriter is closed")
      self._pending.append(_event(time.time() if wall_time is None else wall_time, step, summary=value))
      due = len(self._pending) >= self.max_queue or time.time() - self._last >= self.flush_secs
    if due:
      self.flush()

  def add_scalar(self, tag, value, step, wall_time=None):
    self._add(step, _scalar_value(tag, float(value)), wall_time)

  def add_scalars(self, values, step, wall_time=None):
    """Several tags of one step in ONE event (``{"loss": .., "images_per_s": ..}``)."""
    body = b"".join(_scalar_value(t, float(v)) for t, v in sorted(values.items()))
    self._add(step, body, wall_time)

  def add_histogram(self, tag, values, step, wall_time=None):
    if hasattr(values, "detach"):
      values = values.detach().float().cpu().numpy()
    self._add(step, _histogram_value(tag, values), wall_time)

  def flush(self):
    with self._mutex:
      batch, self._pending = self._pending, []
      self._last = time.time()
      if batch:
        tfrecord.write_records(self.path, batch, append=True)
        self._upload()

  def _upload(self):
    if self.remote is not None:
      from . import fs
      fs.copy_from_local(self.path, self.remote)

  def close(self):
    self.flush()
    self._closed = True

  def __enter__(self):
    return self

  def __exit__(self, *exc):
    self.close()


# ------------------------------------------------------------------ reader
def _fields(buf):
  p, n = 0, len(buf)
  while p < n:
    k, shift = 0, 0
    while True:
      b = buf[p]
      p += 1
      k |= (b & 0x7f) << shift
      shift += 7
      if b < 0x80:
        break
    field, wire = k >> 3, k & 7
    if wire == 0:
      v, shift = 0, 0
      while True:
        b = buf[p]
        p += 1
        v |= (b & 0x7f) << shift
        shift += 7
        if b < 0x80:
          break
      yield field, wire, v
    elif wire == 1:
      yield field, wire, buf[p:p + 8]
      p += 8
    elif wire == 5:
      yield field, wire, buf[p:p + 4]
      p += 4
    elif wire == 2:
      ln, shift = 0, 0
      while True:
        b = buf[p]
        p += 1
        ln |= (b & 0x7f) << shift
        shift += 7
        if b < 0x80:
          break
      yield field, wire, buf[p:p + ln]
      p += ln
    else:
      raise ValueError("unsupported protobuf wire type {}".format(wire))


def _doubles(b):
  return list(struct.unpack("<{}d".format(len(b) // 8), b))


def read_events(path):
  """Events of one file as dicts: ``{"wall_time", "step", "file_version"?, "scalars": {tag: v},
  "histograms": {tag: {"min","max","num","sum","sum_squares","bucket_limit","bucket"}}}``."""
  out = []
  for rec in tfrecord.read_records(path):
    ev = {"wall_time": 0.0, "step": 0, "scalars": {}, "histograms": {}}
    for f, _, v in _fields(rec):
      if f == 1:
        ev["wall_time"] = struct.unpack("<d", v)[0]
      elif f == 2:
        ev["step"] = v
      elif f == 3:
        ev["file_version"] = bytes(v).decode()
      elif f == 5:
        for sf, _, sv in _fields(v):
          if sf != 1:
            continue
          tag, simple, histo = None, None, None
          for vf, _, vv in _fields(sv):
            if vf == 1:
              tag = bytes(vv).decode("utf-8")
            elif vf == 2:
              simple = struct.unpack("<f", vv)[0]
            elif vf == 5:
              names = {1: "min", 2: "max", 3: "num", 4: "sum", 5: "sum_squares"}
              histo = {"bucket_limit": [], "bucket": []}
              for hf, _, hv in _fields(vv):
                if hf in names:
                  histo[names[hf]] = struct.unpack("<d", hv)[0]
                elif hf == 6:
                  histo["bucket_limit"] = _doubles(hv)
                elif hf == 7:
                  histo["bucket"] = _doubles(hv)
          if simple is not None:
            ev["scalars"][tag] = simple
          if histo is not None:
            ev["histograms"][tag] = histo
    out.append(ev)
  return out


def event_files(logdir):
  """The event files of ``logdir``, oldest first."""
  return sorted(os.path.join(logdir, f) for f in os.listdir(logdir) if f.startswith("events.out.tfevents."))
