"""File access for checkpoints, exports and event files on whatever filesystem a path names.

The reference writes these through TensorFlow's filesystem layer, so ``model_dir`` / ``export_dir``
may be ``hdfs://``, ``viewfs://``, ``s3a://`` ... (``ctx.absolute_path`` qualifies relative paths
with the cluster's ``defaultFS``, tensorflowonspark/TFNode.py:32-67) - which is what makes a
checkpoint written by the chief on one host visible to the evaluator, the restarted job or the
serving executors on another.  Here plain paths and ``file://`` go through ``os`` / ``open``
unchanged; every other scheme is resolved with ``pyarrow.fs.FileSystem.from_uri`` (HDFS through
libhdfs, S3, GCS, ... - whatever the installed pyarrow build and the host's client libraries
provide), one filesystem object per ``scheme://authority``.  Only the handful of operations the
callers need is exposed; ``write_atomic`` is "write a sibling temp file, then rename" on both
kinds (atomic on POSIX and HDFS).
"""
import os
import posixpath
import threading
import uuid

_lock = threading.Lock()
_filesystems = {}

# spelling differences between Hadoop-style URIs and what pyarrow expects
_SCHEME_ALIASES = {"s3a": "s3", "s3n": "s3"}


def is_local(path):
  p = str(path)
  return "://" not in p or p.startswith("file://")


def local(path):
  """The plain OS path of a local path (``file://`` stripped)."""
  p = str(path)
  return p[len("file://"):] if p.startswith("file://") else p


def _resolve(path):
  """(pyarrow filesystem, path inside it) for a non-local URI.  ``from_uri`` is called once per
  ``scheme://authority`` (it may open a connection); how the inner path is spelled - with the
  bucket in front (object stores), as an absolute path (HDFS) or relative (the in-memory test
  filesystem) - is learnt from that first answer and applied to later paths."""
  import pyarrow.fs as pafs
  p = str(path)
  scheme, rest = p.split("://", 1)
  authority, _, tail = rest.partition("/")
  key = (scheme, authority)
  with _lock:
    hit = _filesystems.get(key)
    if hit is None:
      uri = _SCHEME_ALIASES.get(scheme, scheme) + "://" + rest
      try:
        fs, inner = pafs.FileSystem.from_uri(uri)
      except Exception as e:
        raise IOError("cannot open the filesystem of {!r}: {} (is the client library for '{}' "
                      "installed on this host?)".format(p, e, scheme))
      if authority and inner.startswith(authority):
        style = "bucket"
      elif inner.startswith("/"):
        style = "absolute"
      else:
        style = "relative"
      hit = _filesystems[key] = (fs, style)
  fs, style = hit
  if style == "bucket":
    return fs, (authority + "/" + tail).rstrip("/")
  if style == "absolute":
    return fs, "/" + tail
  return fs, tail


def join(base, *parts):
  if is_local(base):
    return os.path.join(str(base), *parts)
  return posixpath.join(str(base), *parts)


def basename(path):
  return posixpath.basename(str(path).rstrip("/"))


def dirname(path):
  p = str(path)
  if is_local(p):
    return os.path.dirname(p)
  head, _ = p.rstrip("/").rsplit("/", 1)
  return head


def exists(path):
  if is_local(path):
    return os.path.exists(local(path))
  import pyarrow.fs as pafs
  fs, p = _resolve(path)
  return fs.get_file_info(p).type != pafs.FileType.NotFound


def isdir(path):
  if is_local(path):
    return os.path.isdir(local(path))
  import pyarrow.fs as pafs
  fs, p = _resolve(path)
  return fs.get_file_info(p).type == pafs.FileType.Directory


def makedirs(path):
  if is_local(path):
    os.makedirs(local(path), exist_ok=True)
    return
  fs, p = _resolve(path)
  fs.create_dir(p, recursive=True)


def listdir(path):
  """Base names of the entries of a directory."""
  if is_local(path):
    return os.listdir(local(path))
  import pyarrow.fs as pafs
  fs, p = _resolve(path)
  return [posixpath.basename(i.path) for i in fs.get_file_info(pafs.FileSelector(p))]


def remove(path):
  if is_local(path):
    os.remove(local(path))
    return
  fs, p = _resolve(path)
  fs.delete_file(p)


def open_read(path):
  """Binary, seekable file object."""
  if is_local(path):
    return open(local(path), "rb")
  fs, p = _resolve(path)
  return fs.open_input_file(p)


def read_text(path):
  with open_read(path) as f:
    return f.read().decode("utf-8")


def write_atomic(path, writer):
  """``writer(fileobj)`` fills a temp file next to ``path`` which then replaces ``path``."""
  d = dirname(path)
  tmp = join(d, ".tmp-{}-{}".format(uuid.uuid4().hex[:12], basename(path)))
  if is_local(path):
    os.makedirs(local(d) or ".", exist_ok=True)
    lt, lp = local(tmp), local(path)
    try:
      with open(lt, "wb") as f:
        writer(f)
        f.flush()
        os.fsync(f.fileno())
      os.replace(lt, lp)
    except BaseException:
      if os.path.exists(lt):
        os.remove(lt)
      raise
    return
  import pyarrow.fs as pafs
  fs, p = _resolve(path)
  _, pt = _resolve(tmp)
  fs.create_dir(posixpath.dirname(p), recursive=True)
  try:
    with fs.open_output_stream(pt) as f:
      writer(f)
    if fs.get_file_info(p).type != pafs.FileType.NotFound:
      fs.delete_file(p)               # (object stores and some HDFS versions do not overwrite on move)
    fs.move(pt, p)
  except BaseException:
    try:
      fs.delete_file(pt)
    except Exception:
      pass
    raise


def write_text(path, text):
  write_atomic(path, lambda f: f.write(text.encode("utf-8")))


def copy_from_local(local_path, path):
  """Upload one local file (atomic at the destination)."""
  def writer(f):
    with open(local_path, "rb") as src:
      while True:
        chunk = src.read(1 << 20)
        if not chunk:
          break
        f.write(chunk)
  write_atomic(path, writer)
