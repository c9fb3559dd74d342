"""Python face of the shared-memory pinned ring (csrc/feed.cc).

Feeder tasks pack a block of uniform numeric rows column-by-column into a ring slot
(:func:`pack_rows`) and post a ``marker.RingBlock`` descriptor on the manager queue; the
training process maps the same segment, optionally page-locks it (``ring.pin()``), and either
expands the block back into rows (:func:`unpack_rows`) or DMA-copies the columns straight to
the GPU (``feed.DevicePrefetcher.push_ring_slot``).

A pure-Python ring with the same interface (backed by ``multiprocessing.shared_memory``)
is used only until the native extension has been built; it cannot be page-locked.
"""
import logging
import os
import struct
import time
import uuid

import numpy as np

from . import _build, marker

logger = logging.getLogger(__name__)

DEFAULT_SLOTS = 4  # 4 x 16 MiB: stays cache resident, 64 MiB of first-touch faults (measured: tools/bench_feed.py)
DEFAULT_SLOT_BYTES = 16 << 20


class _PyRing(object):
  """Fallback ring: same slot protocol, per-slot sequence numbers in shared memory."""
  _HDR = 64

  def __init__(self, name, create, nslots=0, slot_bytes=0):
    from multiprocessing import shared_memory
    self._name = name.lstrip("/")
    if create:
      slot_bytes = (slot_bytes + 4095) // 4096 * 4096
      size = self._HDR + 32 * nslots + slot_bytes * nslots
      self._shm = shared_memory.SharedMemory(name=self._name, create=True, size=size)
      struct.pack_into("<IIQQQI", self._shm.buf, 0, 0x74664f53, nslots, slot_bytes, 0, 0, 0)
      for i in range(nslots):
        struct.pack_into("<QQqI", self._shm.buf, self._HDR + 32 * i, i, 0, 0, 0)
    else:
      self._shm = shared_memory.SharedMemory(name=self._name)
    _, self.nslots, self.slot_bytes = struct.unpack_from("<IIQ", self._shm.buf, 0)
    self._owner = create
    self.pinned = False
    import fcntl
    self._lockf = open("/tmp/.tfos_ring_{}.lock".format(self._name), "a+")
    self._fcntl = fcntl

  def _locked(self):
    ring = self

    class _L(object):

      def __enter__(self_inner):
        ring._fcntl.flock(ring._lockf, ring._fcntl.LOCK_EX)

      def __exit__(self_inner, *a):
        ring._fcntl.flock(ring._lockf, ring._fcntl.LOCK_UN)

    return _L()

  def _seq(self, i):
    return struct.unpack_from("<Q", self._shm.buf, self._HDR + 32 * i)[0]

  def _acquire(self, field_off, want_delta, timeout):
    deadline = time.time() + timeout
    while True:
      with self._locked():
        pos = struct.unpack_from("<Q", self._shm.buf, field_off)[0]
        if self._seq(pos % self.nslots) == pos + want_delta:
          struct.pack_into("<Q", self._shm.buf, field_off, pos + 1)
          return pos
        closed = struct.unpack_from("<I", self._shm.buf, 32)[0]
        head = struct.unpack_from("<Q", self._shm.buf, 16)[0]
      if want_delta == 1 and closed and head == pos:
        return -2
      if time.time() > deadline or (want_delta == 0 and closed):
        return -1
      time.sleep(0.0005)

  def acquire_write(self, timeout):
    return self._acquire(16, 0, timeout)

  def acquire_read(self, timeout):
    return self._acquire(24, 1, timeout)

  def commit_write(self, pos, nbytes, nrows, tag=0):
    i = pos % self.nslots
    struct.pack_into("<QqI", self._shm.buf, self._HDR + 32 * i + 8, nbytes, tag, nrows)
    struct.pack_into("<Q", self._shm.buf, self._HDR + 32 * i, pos + 1)

  def release_read(self, pos):
    struct.pack_into("<Q", self._shm.buf, self._HDR + 32 * (pos % self.nslots), pos + self.nslots)

  def meta(self, pos):
    nbytes, tag, nrows = struct.unpack_from("<QqI", self._shm.buf,
                                            self._HDR + 32 * (pos % self.nslots) + 8)
    return nbytes, nrows, tag

  def slot_view(self, pos):
    off = self._HDR + 32 * self.nslots + self.slot_bytes * (pos % self.nslots)
    return self._shm.buf[off:off + self.slot_bytes]

  def depth(self):
    head, tail = struct.unpack_from("<QQ", self._shm.buf, 16)
    return head - tail

  def close(self):
    struct.pack_into("<I", self._shm.buf, 32, 1)

  def closed(self):
    return struct.unpack_from("<I", self._shm.buf, 32)[0] != 0

  def pin(self):
    raise RuntimeError("the pure-Python ring cannot be page-locked; build the native extension")

  def h2d(self, *a, **k):
    raise RuntimeError("h2d needs the native ring")

  def __del__(self):
    try:
      self._shm.close()
      if self._owner:
        self._shm.unlink()
    except Exception:
      pass


def _ring_cls():
  C = _build.load(required=False)
  return C.ShmRing if C is not None else _PyRing


def new_name():
  return "/tfos_ring_{}_{}".format(os.getpid(), uuid.uuid4().hex[:8])


def _unlink(name):
  try:
    os.unlink("/dev/shm/" + name.lstrip("/"))
  except OSError:
    pass


def create(nslots=DEFAULT_SLOTS, slot_bytes=DEFAULT_SLOT_BYTES, name=None):
  import atexit
  name = name or new_name()
  ring = _ring_cls()(name, True, nslots, slot_bytes)
  atexit.register(_unlink, name)
  _cleanups.append((_unlink, name))  # executors torn down by SIGTERM run these (run_cleanups)
  return name, ring


_cleanups = []


def run_cleanups():
  """Unlink every ring this process created (called from the executors' SIGTERM handler, which
  must not run arbitrary third-party atexit hooks half-way through an import)."""
  while _cleanups:
    fn, arg = _cleanups.pop()
    try:
      fn(arg)
    except Exception:
      pass


_attached = {}


def attach(name):
  """Map an existing ring (cached per process)."""
  r = _attached.get(name)
  if r is None:
    r = _ring_cls()(name, False, 0, 0)
    _attached[name] = r
  return r


def _as_columns(rows):
  """rows (sequence of equal-length sequences of numbers/arrays) -> list of 2-D numpy arrays, or None."""
  if not rows:
    return None
  first = rows[0]
  if isinstance(first, (list, tuple)) and first:
    # a row of same-typed scalars (e.g. a CSV line: label, pixel0..pixel783) is ONE [n, len]
    # matrix, not len(row) columns
    kind = None
    if all(isinstance(v, (int, np.integer)) and not isinstance(v, (bool, np.bool_)) for v in first):
      kind = "iu"
    elif all(isinstance(v, (float, np.floating)) for v in first):
      kind = "f"
    if kind is not None:
      try:
        arr = np.asarray(rows)
        if arr.ndim == 2 and arr.dtype.kind in kind:
          return [np.ascontiguousarray(arr)], False
      except Exception:
        pass
  if isinstance(first, (list, tuple)):
    ncol = len(first)
    cols = []
    for c in range(ncol):
      try:
        arr = np.asarray([r[c] for r in rows])
      except Exception:
        return None
      if arr.dtype == object or arr.dtype.kind not in "biuf":
        return None
      cols.append(np.ascontiguousarray(arr))
    return cols, True
  try:
    arr = np.asarray(rows)
  except Exception:
    return None
  if arr.dtype == object or arr.dtype.kind not in "biuf":
    return None
  return [np.ascontiguousarray(arr)], False


def rows_per_slot(ring, sample_row):
  """How many rows shaped like ``sample_row`` fit one ring slot (FEED_CHUNK-sized if unknown)."""
  try:
    cols = sample_row if isinstance(sample_row, (list, tuple)) else [sample_row]
    nbytes = sum(int(np.asarray(c).nbytes) for c in cols)
    if nbytes <= 0:
      return 1 << 20
    return max(1, int((ring.slot_bytes - 64 * (len(cols) + 1)) // (nbytes + 8)))
  except Exception:
    return 1 << 20


def _uniform_array_columns(rows):
  """[(dtype, shape)] per column when every row is a tuple/list of equally shaped ndarrays."""
  first = rows[0]
  if not isinstance(first, (list, tuple)) or not first:
    return None
  spec = []
  for c in first:
    if not isinstance(c, np.ndarray) or c.dtype.kind not in "biuf":
      return None
    spec.append((c.dtype, c.shape))
  ncol = len(spec)
  for r in rows:
    if len(r) != ncol:
      return None
    for c, (dt, shape) in zip(r, spec):
      if not isinstance(c, np.ndarray) or c.dtype != dt or c.shape != shape:
        return None
  return spec


def _pack_arrays_direct(ring, rows, spec, timeout):
  """Rows of ndarrays go straight into the slot (np.stack(out=slot view)): one copy, no
  intermediate [n, ...] array."""
  n = len(rows)
  sizes = [int(np.prod(shape, dtype=np.int64)) * dt.itemsize * n for dt, shape in spec]
  if sum((s + 63) // 64 * 64 for s in sizes) > ring.slot_bytes:
    return None
  pos = ring.acquire_write(timeout)
  if pos < 0:
    raise RuntimeError("timed out waiting for a free ring slot (consumer stalled?)")
  view = np.frombuffer(ring.slot_view(pos), dtype=np.uint8)
  layout, off = [], 0
  for j, ((dt, shape), nb) in enumerate(zip(spec, sizes)):
    dst = view[off:off + nb].view(dt).reshape((n,) + tuple(shape))
    np.stack([r[j] for r in rows], out=dst)
    layout.append((off, nb, dt.str, tuple(shape)))
    off = (off + nb + 63) // 64 * 64
  ring.commit_write(pos, off, n, 1)
  return marker.RingBlock(pos, n, layout)


def pack_rows(ring, rows, timeout=600.0):
  """Write a block of rows into the next free slot; returns a RingBlock or None if the rows are
  not uniform numeric data (caller falls back to the queue path)."""
  if rows:
    spec = _uniform_array_columns(rows)
    if spec is not None:
      blk = _pack_arrays_direct(ring, rows, spec, timeout)
      if blk is not None:
        return blk
  packed = _as_columns(rows)
  if packed is None:
    return None
  cols, tupled = packed
  total = sum(int(c.nbytes) + 64 for c in cols)
  if total > ring.slot_bytes:
    return None
  pos = ring.acquire_write(timeout)
  if pos < 0:
    raise RuntimeError("timed out waiting for a free ring slot (consumer stalled?)")
  view = np.frombuffer(ring.slot_view(pos), dtype=np.uint8)
  layout, off = [], 0
  for c in cols:
    nb = int(c.nbytes)
    view[off:off + nb] = c.reshape(-1).view(np.uint8)
    layout.append((off, nb, c.dtype.str, tuple(c.shape[1:])))
    off = (off + nb + 63) // 64 * 64
  ring.commit_write(pos, off, len(rows), 1 if tupled else 0)
  return marker.RingBlock(pos, len(rows), layout if tupled else [layout[0] + ("flat",)])


def unpack_columns(ring, block):
  """Zero-copy numpy views (one per column) of a ring block; valid until release_read(pos)."""
  view = np.frombuffer(ring.slot_view(block.pos), dtype=np.uint8)
  cols = []
  for entry in block.layout:
    off, nb, dt, shape = entry[:4]
    cols.append(view[off:off + nb].view(np.dtype(dt)).reshape((block.nrows,) + tuple(shape)))
  return cols


def unpack_rows(ring, block):
  """Expand a ring block back into python rows (copying out of the slot)."""
  cols = unpack_columns(ring, block)
  flat = len(block.layout) == 1 and len(block.layout[0]) == 5
  if flat:
    return cols[0].tolist()
  lists = [c.tolist() for c in cols]
  return [list(r) for r in zip(*lists)]
