"""TFRecord framing + tf.train.Example codec (no TensorFlow, no JVM).

Native implementation: csrc/tfrecord.cc (masked CRC32C, slicing-by-8; hand-rolled
protobuf wire codec).  The pure-Python twin below is used only until the
extension has been built; both are bit-compatible and tested against each other.

Replaces lib/tensorflow-hadoop-1.0-SNAPSHOT.jar (TFRecordFileInputFormat /
OutputFormat) and the Example conversion of the reference
(tensorflowonspark/dfutil.py:84-212, DFUtil.scala:119-258).
"""
import struct

from . import _build

_MASK_DELTA = 0xa282ead8
_table = None


def _native():
  return _build.load(required=False)


def _py_crc32c(data):
  global _table
  if _table is None:
    t = []
    for i in range(256):
      c = i
      for _ in range(8):
        c = (c >> 1) ^ 0x82f63b78 if c & 1 else c >> 1
      t.append(c)
    _table = t
  c = 0xffffffff
  for b in data:
    c = _table[(c ^ b) & 0xff] ^ (c >> 8)
  return c ^ 0xffffffff


def crc32c(data):
  n = _native()
  return n.crc32c(bytes(data)) if n is not None else _py_crc32c(bytes(data))


def masked_crc32c(data):
  crc = crc32c(data)
  return (((crc >> 15) | (crc << 17)) + _MASK_DELTA) & 0xffffffff


def _remote(path):
  from .utils import fs
  return None if fs.is_local(path) else fs


def write_records(path, records, append=False):
  """TFRecord file from an iterable of serialized records.  ``path`` may be a URI of a remote
  filesystem (utils/fs.py): the file is then written locally and uploaded (no ``append``)."""
  records = [bytes(r) for r in records]
  rfs = _remote(path)
  if rfs is not None:
    if append:
      raise ValueError("cannot append to a TFRecord file on a remote filesystem: " + str(path))
    import os
    import tempfile
    fd, tmp = tempfile.mkstemp(suffix=".tfrecord")
    os.close(fd)
    try:
      write_records(tmp, records)
      rfs.copy_from_local(tmp, path)
    finally:
      os.remove(tmp)
    return None
  path = path[len("file://"):] if str(path).startswith("file://") else path
  n = _native()
  if n is not None:
    return n.tfrecord_write(path, records, append)
  with open(path, "ab" if append else "wb") as f:
    for r in records:
      hdr = struct.pack("<Q", len(r))
      f.write(hdr)
      f.write(struct.pack("<I", masked_crc32c(hdr)))
      f.write(r)
      f.write(struct.pack("<I", masked_crc32c(r)))


def read_records(path, verify=True):
  """Records of one TFRecord file (CRCs checked unless ``verify=False``).  A URI of a remote
  filesystem (``hdfs://`` ..., utils/fs.py) is streamed through that filesystem and framed here,
  with the native CRC32C."""
  rfs = _remote(path)
  if rfs is not None:
    with rfs.open_read(path) as f:
      return _read_stream(f, verify, str(path))
  path = path[len("file://"):] if str(path).startswith("file://") else path
  n = _native()
  if n is not None:
    return n.tfrecord_read(path, verify)
  with open(path, "rb") as f:
    return _read_stream(f, verify, path)


def _read_stream(f, verify, path):
  out = []
  while True:
    hdr = f.read(12)
    if not hdr:
      break
    if len(hdr) != 12:
      raise IOError("truncated TFRecord header in " + path)
    (length,), (crc,) = struct.unpack("<Q", hdr[:8]), struct.unpack("<I", hdr[8:])
    if verify and masked_crc32c(hdr[:8]) != crc:
      raise IOError("corrupt TFRecord length CRC in " + path)
    data = f.read(length)
    tail = f.read(4)
    if len(data) != length or len(tail) != 4:
      raise IOError("truncated TFRecord payload in " + path)
    if verify and masked_crc32c(data) != struct.unpack("<I", tail)[0]:
      raise IOError("corrupt TFRecord data CRC in " + path)
    out.append(data)
  return out


# ------------------------------------------------------------ Example codec
def _varint(v):
  v &= 0xffffffffffffffff
  out = bytearray()
  while v >= 0x80:
    out.append((v & 0x7f) | 0x80)
    v >>= 7
  out.append(v)
  return bytes(out)


def _ld(field, payload):
  return _varint((field << 3) | 2) + _varint(len(payload)) + payload


def _py_encode(features):
  feats = b""
  for name, (kind, vals) in features.items():
    if kind == "bytes":
      feature = _ld(1, b"".join(_ld(1, bytes(v)) for v in vals))
    elif kind == "float":
      packed = b"".join(struct.pack("<f", float(v)) for v in vals)
      feature = _ld(2, _ld(1, packed) if packed else b"")
    elif kind == "int64":
      packed = b"".join(_varint(int(v)) for v in vals)
      feature = _ld(3, _ld(1, packed) if packed else b"")
    else:
      raise ValueError("unknown feature kind " + kind)
    feats += _ld(1, _ld(1, name.encode("utf-8")) + _ld(2, feature))
  return _ld(1, feats)


class _R(object):

  def __init__(self, b, p=0, e=None):
    self.b, self.p, self.e = b, p, len(b) if e is None else e

  def done(self):
    return self.p >= self.e

  def varint(self):
    v = shift = 0
    while True:
      x = self.b[self.p]
      self.p += 1
      v |= (x & 0x7f) << shift
      if not x & 0x80:
        return v
      shift += 7

  def sub(self):
    n = self.varint()
    r = _R(self.b, self.p, self.p + n)
    self.p += n
    return r

  def skip(self, wt):
    if wt == 0:
      self.varint()
    elif wt == 1:
      self.p += 8
    elif wt == 2:
      self.sub()
    elif wt == 5:
      self.p += 4
    else:
      raise ValueError("unsupported wire type")

  def raw(self):
    return bytes(self.b[self.p:self.e])


def _signed(v):
  return v - (1 << 64) if v >= (1 << 63) else v


def _py_decode(data):
  out = {}
  ex = _R(bytes(data))
  while not ex.done():
    tag = ex.varint()
    if tag >> 3 != 1 or tag & 7 != 2:
      ex.skip(tag & 7)
      continue
    feats = ex.sub()
    while not feats.done():
      t2 = feats.varint()
      if t2 >> 3 != 1 or t2 & 7 != 2:
        feats.skip(t2 & 7)
        continue
      entry = feats.sub()
      name, kind, values = "", "bytes", []
      while not entry.done():
        t3 = entry.varint()
        field, wt = t3 >> 3, t3 & 7
        if field == 1 and wt == 2:
          name = entry.sub().raw().decode("utf-8")
        elif field == 2 and wt == 2:
          feature = entry.sub()
          while not feature.done():
            t4 = feature.varint()
            kf = t4 >> 3
            if t4 & 7 != 2:
              feature.skip(t4 & 7)
              continue
            lst = feature.sub()
            kind = {1: "bytes", 2: "float"}.get(kf, "int64")
            while not lst.done():
              t5 = lst.varint()
              wt5 = t5 & 7
              if t5 >> 3 != 1:
                lst.skip(wt5)
              elif kf == 1:
                values.append(lst.sub().raw())
              elif kf == 2:
                if wt5 == 2:
                  raw = lst.sub().raw()
                  values.extend(struct.unpack("<{}f".format(len(raw) // 4), raw))
                else:
                  values.append(struct.unpack("<f", lst.b[lst.p:lst.p + 4])[0])
                  lst.p += 4
              else:
                if wt5 == 2:
                  pk = lst.sub()
                  while not pk.done():
                    values.append(_signed(pk.varint()))
                else:
                  values.append(_signed(lst.varint()))
        else:
          entry.skip(wt)
      out[name] = (kind, values)
  return out


def encode_example(features):
  """features: dict name -> (kind, list); kind in {'bytes', 'float', 'int64'} -> serialized Example."""
  n = _native()
  if n is not None:
    return n.example_encode({k: (v[0], list(v[1])) for k, v in features.items()})
  return _py_encode(features)


def decode_example(data):
  """serialized Example -> dict name -> (kind, list)."""
  n = _native()
  if n is not None:
    return dict(n.example_decode(bytes(data)))
  return _py_decode(data)


_OUT_DTYPES = {"bytes": "uint8", "float": "float32", "int64": "int64"}


def decode_batch(records, spec, threads=None):
  """N serialized Examples -> ``{name: ndarray [N, length]}`` without a python object per value.

  ``spec``: ``{name: (kind, length)}`` or ``{name: (kind, length, dtype)}`` - kind 'int64' decodes
  to int64 (default), int32 or uint8; 'float' to float32; 'bytes' (ONE value of exactly
  ``length`` bytes, e.g. raw pixels) to uint8.  Every record must hold exactly ``length`` values
  of every requested feature; other features are skipped without being materialised.  The native
  decoder (csrc/tfrecord.cc: decode_batch) walks the wire format once, writes straight into the
  arrays, releases the GIL and splits large batches over threads."""
  import numpy as np
  records = [r if isinstance(r, bytes) else bytes(r) for r in records]
  norm = []
  for name, s in spec.items():
    kind, length = s[0], int(s[1])
    dt = np.dtype(s[2]).name if len(s) > 2 and s[2] is not None else _OUT_DTYPES[kind]
    norm.append((name, kind, length, dt))
  n = _native()
  if n is not None:
    if threads is None:
      threads = min(8, max(1, len(records) // 4096))   # threads pay off for large batches only
    return dict(n.example_decode_batch(records, norm, int(threads)))
  out = {name: np.empty((len(records), length), dtype=dt) for name, _, length, dt in norm}
  for r, rec in enumerate(records):
    ex = _py_decode(rec)
    for name, kind, length, dt in norm:
      if name not in ex:
        raise RuntimeError("record {}: feature '{}' has no values, {} requested".format(r, name, length))
      k, vals = ex[name]
      if k != kind:
        raise RuntimeError("feature '{}' has another type than requested".format(name))
      if kind == "bytes":
        if len(vals) != 1 or len(vals[0]) != length:
          raise RuntimeError("bytes feature '{}' is not one value of the requested length".format(name))
        out[name][r] = np.frombuffer(vals[0], dtype=np.uint8)
      else:
        if len(vals) != length:
          raise RuntimeError("record {}: feature '{}' has {} values, {} requested".format(
              r, name, len(vals), length))
        out[name][r] = np.asarray(vals).astype(dt)
  return out
