"""File-descriptor hand-off between the rank processes of one host (SCM_RIGHTS over AF_UNIX).

cuMem allocations and NVLS multicast objects are shared as POSIX file descriptors
(csrc/vmm.cc); unlike a CUDA-IPC handle a descriptor is only meaningful inside the process that
owns it, so it has to be passed through a Unix-domain socket.  Every rank runs one ``FdServer``
(abstract-namespace socket, a daemon thread) holding the descriptors it exported under string
keys; a peer connects, names a key and receives a duplicate.  Socket addresses and keys travel
over the ordinary control channel (torch.distributed / the reservation board) next to the CUDA
IPC handles.
"""
import os
import socket
import struct
import threading
import uuid


class FdServer(object):

  def __init__(self):
    self.address = "\0tfos-fd-{}-{}".format(os.getpid(), uuid.uuid4().hex[:12])
    self._fds = {}
    self._lock = threading.Lock()
    self._sock = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    self._sock.bind(self.address)
    self._sock.listen(64)
    self._closed = False
    self._thread = threading.Thread(target=self._serve, name="tfos-fdshare", daemon=True)
    self._thread.start()

  def register(self, key, fd):
    with self._lock:
      self._fds[key] = fd

  def _serve(self):
    while not self._closed:
      try:
        conn, _ = self._sock.accept()
      except OSError:
        return
      try:
        (n,) = struct.unpack(">I", _read_exact(conn, 4))
        key = _read_exact(conn, n).decode("utf-8")
        with self._lock:
          fd = self._fds.get(key)
        if fd is None:
          conn.sendall(b"N")
        else:
          socket.send_fds(conn, [b"Y"], [fd])
      except Exception:
        pass
      finally:
        conn.close()

  def close(self):
    self._closed = True
    try:
      self._sock.close()
    except OSError:
      pass
    with self._lock:
      for fd in self._fds.values():
        try:
          os.close(fd)
        except OSError:
          pass
      self._fds = {}


def _read_exact(conn, n):
  buf = b""
  while len(buf) < n:
    chunk = conn.recv(n - len(buf))
    if not chunk:
      raise EOFError("fd share peer closed the connection")
    buf += chunk
  return buf


def fetch_fd(address, key, timeout=60.0):
  """Receive a duplicate of the descriptor a peer registered under ``key``."""
  s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
  s.settimeout(timeout)
  try:
    s.connect(address)
    k = key.encode("utf-8")
    s.sendall(struct.pack(">I", len(k)) + k)
    msg, fds, _, _ = socket.recv_fds(s, 1, 1)
    if msg != b"Y" or not fds:
      raise KeyError("peer has no descriptor named {!r}".format(key))
    return fds[0]
  finally:
    s.close()
