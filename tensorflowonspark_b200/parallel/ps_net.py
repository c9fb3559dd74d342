"""Parameter server over TCP: the transport for ps and worker nodes on DIFFERENT hosts.

The peer-mapped parameter server (parallel/ps.py) addresses the ps GPU's memory through CUDA IPC
(or POSIX shared memory on CPU executors): one machine.  The reference's parameter server is a
network service - ``tf.train.Server`` / gRPC between executors that Spark places anywhere
(tensorflowonspark/TFNode.py:126-132; ``TFCluster.run(num_ps=...)``, TFCluster.py:260-262) - so a
cluster whose ps and workers sit on several hosts needs a wire protocol.  This module is that
protocol with the SAME surface as ``PSServer`` / ``PSClient`` (``pull / push / push_sparse`` in
plain mode; ``pull_model / push_grads / set_lr`` in slot mode, which is what ``PSWorker`` drives):

  * the ps node listens on a TCP port published on the reservation board; every client opens one
    connection per server and speaks a 32-byte header + raw little-endian payload (no pickling
    of tensors: gradients and parameters cross as the bytes they are);
  * server side, one thread per connection receives into a reusable buffer and applies under a
    lock - on a GPU ps through the same ``ps_apply`` kernel as the peer-mapped server (state,
    fp32 master and bf16 serving copy resident on the GPU, the received gradient staged through
    pinned memory), on a CPU ps through the numpy twin;
  * ``pull_model`` ships bf16 for the decayed weights and fp32 for the tail the model consumes
    in fp32 (batch-norm scale / offset, biases, running statistics): half the bytes of an fp32
    pull, like the peer-mapped path;
  * ``push_grads`` does not wait for its acknowledgement until the next push on that connection:
    the apply of push i overlaps the worker's step i + 1 (the slot back-pressure of ps.py,
    depth one).

``ps.attach`` picks this transport when the cluster spec names more than one host (or
``TFOS_PS_TRANSPORT=tcp``); workers follow whatever the servers published.
"""
import logging
import socket
import struct
import threading
import time

import numpy as np

from . import ps as _ps

logger = logging.getLogger(__name__)

HDR = struct.Struct("<B3xIqdQ")     # op, client, seq / aux, scalar, payload bytes
OP_PULL, OP_PULL_MODEL, OP_PUSH_GRADS, OP_PUSH_DENSE, OP_PUSH_SPARSE, OP_SET_LR, OP_STATS, OP_BYE = \
    1, 2, 3, 4, 5, 6, 7, 8
SPARSE = struct.Struct("<IIq")      # rows, width, base


def _recv_exact(sock, view):
  got, n = 0, len(view)
  while got < n:
    k = sock.recv_into(view[got:], n - got)
    if k == 0:
      raise ConnectionError("peer closed the connection after {} of {} bytes".format(got, n))
    got += k


def _send(sock, op, client=0, seq=0, scalar=0.0, payload=b""):
  payload = memoryview(payload).cast("B") if not isinstance(payload, (bytes, bytearray)) else payload
  sock.sendall(HDR.pack(op, client, seq, scalar, len(payload)))
  if len(payload):
    sock.sendall(payload)


def _recv(sock, buf=None):
  head = bytearray(HDR.size)
  _recv_exact(sock, memoryview(head))
  op, client, seq, scalar, nbytes = HDR.unpack(head)
  if buf is None or len(buf) < nbytes:
    buf = bytearray(nbytes)
  view = memoryview(buf)[:nbytes]
  if nbytes:
    _recv_exact(sock, view)
  return op, client, seq, scalar, view, buf


def _bf16_bytes(x_f32):
  """fp32 numpy -> bf16 (round-to-nearest-even) as raw uint16."""
  import torch
  return torch.from_numpy(np.ascontiguousarray(x_f32)).to(torch.bfloat16).view(torch.int16).numpy()


def _bf16_to_f32(raw_u16):
  import torch
  return torch.from_numpy(np.ascontiguousarray(raw_u16).view(np.int16)).view(torch.bfloat16).float().numpy()


class NetPSServer(object):
  """Runs inside a 'ps' node: hosts parameters [lo, hi) of the flat vector behind a TCP port."""

  def __init__(self, ctx, numel, init=None, optimizer=None, lr=0.1, momentum=0.9,
               weight_decay=0.0, beta1=0.9, beta2=0.999, eps=1e-7, decay_end=None,
               ema_begin=None, grad_scale=1.0):
    from .. import util
    self.ctx, self.numel = ctx, int(numel)
    self.slot_mode = optimizer is not None
    self.opt = _ps.OPTS[optimizer] if self.slot_mode else 0
    if self.slot_mode and self.numel % 8:
      raise ValueError("slot mode needs a parameter count that is a multiple of 8")
    self.lo, self.hi = _ps._slices(self.numel, _ps._ps_count(ctx))[ctx.task_index]
    n = self.n = self.hi - self.lo
    self.decay_end = self.numel if decay_end is None else int(decay_end)
    self.ema_begin = self.numel if ema_begin is None else int(ema_begin)
    self.clients = _ps._clients(ctx)
    self.cuda = _ps._use_cuda(ctx)
    hyper = [lr, momentum, weight_decay, grad_scale, beta1, beta2, eps, 0.0]
    start = None
    if init is not None:
      start = (init.detach().float().cpu().numpy() if hasattr(init, "detach")
               else np.asarray(init, dtype=np.float32)).reshape(-1)[self.lo:self.hi]
    if self.cuda:
      import torch
      torch.cuda.set_device(0)
      dev = torch.device("cuda", 0)
      self.master = torch.zeros(max(n, 8), dtype=torch.float32, device=dev)[:n]
      self.state1, self.state2 = torch.zeros_like(self.master), torch.zeros_like(self.master)
      self.wbf16 = torch.zeros(max(n, 8), dtype=torch.bfloat16, device=dev)[:n]
      self.slot = torch.zeros(max(n, 8), dtype=torch.float32, device=dev)[:n]
      self.hyper = torch.tensor(hyper, dtype=torch.float32, device=dev)
      self._applied = torch.zeros(8, dtype=torch.int32, device=dev)
      self._counter = torch.zeros(8, dtype=torch.int32, device=dev)
      self._stage = torch.zeros(max(n, 8), dtype=torch.float32).pin_memory()
      if start is not None:
        self.master.copy_(torch.from_numpy(np.ascontiguousarray(start)))
      self.wbf16.copy_(self.master)
      torch.cuda.synchronize()
      self.array, self.tensor = None, self.master
    else:
      self.master = np.zeros(n, dtype=np.float32)
      self.state1, self.state2 = np.zeros(n, dtype=np.float32), np.zeros(n, dtype=np.float32)
      self.hyper = np.asarray(hyper, dtype=np.float32)
      if start is not None:
        self.master[:] = start
      self.array, self.tensor = self.master, None
    self.lock = threading.Lock()
    self.applies = 0
    self._polled = 0
    self._closing = False
    self.sock = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
    self.sock.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
    self.sock.bind(("", 0))
    self.sock.listen(128)
    self.port = self.sock.getsockname()[1]
    self.host = util.get_ip_address()
    self._threads = []
    self._acceptor = threading.Thread(target=self._accept_loop, name="ps-net-accept", daemon=True)
    self._acceptor.start()
    board = _ps._board(ctx)
    board.put("ps/{}/{}".format(ctx.cluster_id, ctx.task_index), {
        "kind": "tcp", "host": self.host, "port": self.port, "lo": self.lo, "hi": self.hi,
        "numel": self.numel, "slot_mode": self.slot_mode, "clients": [list(c) for c in self.clients],
        "decay_end": self.decay_end, "ema_begin": self.ema_begin, "opt": self.opt})
    board.close()
    logger.info("ps:%d serving parameters [%d, %d) over tcp://%s:%d (%s%s)", ctx.task_index, self.lo,
                self.hi, self.host, self.port, "GPU" if self.cuda else "CPU",
                ", slot mode " + optimizer if self.slot_mode else "")

  # ------------------------------------------------------------------ connections
  def _accept_loop(self):
    while not self._closing:
      try:
        conn, addr = self.sock.accept()
      except OSError:
        return
      conn.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
      t = threading.Thread(target=self._serve, args=(conn, addr), name="ps-net-conn", daemon=True)
      t.start()
      self._threads.append(t)

  def _serve(self, conn, addr):
    buf = None
    try:
      while True:
        op, client, seq, scalar, view, buf = _recv(conn, buf)
        if op == OP_BYE:
          return
        reply = self._dispatch(op, client, seq, scalar, view)
        _send(conn, op, client, seq, float(self.applies), reply)
    except ConnectionError:
      pass
    except Exception:
      logger.exception("ps connection from %s failed", addr)
    finally:
      conn.close()

  def _dispatch(self, op, client, seq, scalar, view):
    with self.lock:
      if op == OP_PULL:
        return self._values_locked().tobytes()
      if op == OP_PULL_MODEL:
        return self._model_bytes(int(seq))
      if op == OP_PUSH_GRADS:
        self._apply(np.frombuffer(view, dtype=np.float32))
        return b""
      if op == OP_PUSH_DENSE:
        self._axpy(0, np.frombuffer(view, dtype=np.float32), -scalar)
        self.applies += 1
        return b""
      if op == OP_PUSH_SPARSE:
        rows, width, base = SPARSE.unpack(bytes(view[:SPARSE.size]))
        idx = np.frombuffer(view[SPARSE.size:SPARSE.size + 8 * rows], dtype=np.int64)
        g = np.frombuffer(view[SPARSE.size + 8 * rows:], dtype=np.float32).reshape(rows, width)
        self._sparse(g, idx, width, base, -scalar)
        return b""
      if op == OP_SET_LR:
        if self.cuda:
          self.hyper[0:1].fill_(float(scalar))
        else:
          self.hyper[0] = scalar
        return b""
      if op == OP_STATS:
        return struct.pack("<q", self.applies)
    raise ValueError("unknown parameter-server request {}".format(op))

  # ------------------------------------------------------------------ state access (lock held)
  def _values_locked(self):
    if self.cuda:
      return self.master.detach().cpu().numpy()
    return self.master

  def _model_bytes(self, decay_end):
    """bf16 for the owned elements below ``decay_end``, fp32 for the rest."""
    a = min(max(decay_end - self.lo, 0), self.n)
    if self.cuda:
      import torch
      head = self.wbf16[:a].cpu().view(torch.int16).numpy()
      tail = self.master[a:].cpu().numpy()
    else:
      head, tail = _bf16_bytes(self.master[:a]), self.master[a:]
    return head.tobytes() + tail.tobytes()

  def _axpy(self, off, g, k):
    """master[off : off + len(g)] += k * g   (plain mode: the workers' own learning rate)."""
    if self.cuda:
      import torch
      t = torch.from_numpy(np.ascontiguousarray(g)).to(self.master.device)
      self.master[off:off + g.size].add_(t, alpha=float(k))
      self.wbf16[off:off + g.size].copy_(self.master[off:off + g.size])
    else:
      self.master[off:off + g.size] += np.float32(k) * g

  def _sparse(self, g, idx, width, base, k):
    for row, r in zip(g, idx):
      s = base + int(r) * width
      b0, b1 = max(s, self.lo), min(s + width, self.hi)
      if b0 < b1:                 # the part of the row this server owns
        self._axpy(b0 - self.lo, row[b0 - s:b1 - s], k)
    self.applies += 1

  def _apply(self, g):
    """Slot mode: one optimizer step with the received gradient slice (+ running-stat deltas)."""
    n = self.n
    if g.size != n:
      raise ValueError("gradient of {} elements for a slice of {}".format(g.size, n))
    if self.cuda:
      import torch
      from .. import ops
      self._stage[:n].copy_(torch.from_numpy(g))
      self.slot.copy_(self._stage[:n], non_blocking=True)
      if self.opt == 2:
        self.hyper[7:8].add_(1.0)
      ops.K.ps_apply({
          "master": self.master.data_ptr(), "state1": self.state1.data_ptr(),
          "state2": self.state2.data_ptr(), "wbf16": self.wbf16.data_ptr(),
          "slot": self.slot.data_ptr(), "hyper": self.hyper.data_ptr(),
          "n": n, "lo": self.lo, "decay_end": self.decay_end, "ema_begin": self.ema_begin,
          "applied_flag": self._applied.data_ptr(), "seq": self.applies + 1,
          "block_counter": self._counter.data_ptr(), "opt": self.opt})
      torch.cuda.current_stream().synchronize()   # the staging buffer is reused by the next push
    else:
      _ps.apply_numpy(self.master, self.state1, self.state2, self.hyper, g, self.lo, self.decay_end,
                      self.ema_begin, self.opt)
    self.applies += 1

  # ------------------------------------------------------------------ PSServer surface
  def poll_once(self):
    """Requests are applied by the connection threads as they arrive; this reports how many
    were applied since the previous call (same contract as PSServer.poll_once)."""
    done, self._polled = self.applies - self._polled, self.applies
    return done

  def values(self):
    with self.lock:
      return np.array(self._values_locked(), copy=True)

  def serve_forever(self, poll=1.0):
    while not self._closing:
      time.sleep(poll)

  def close(self):
    self._closing = True
    try:
      self.sock.close()
    except OSError:
      pass


class NetPSClient(object):
  """Worker-side handle on parameter servers reached over TCP (same surface as ps.PSClient)."""

  def __init__(self, ctx, timeout=600, parts=None):
    self.ctx = ctx
    if parts is None:
      board = _ps._board(ctx)
      parts = [board.get("ps/{}/{}".format(ctx.cluster_id, i), timeout)
               for i in range(_ps._ps_count(ctx))]
      board.close()
    self.parts = parts
    self.numel = parts[0]["numel"]
    self.slot_mode = bool(parts[0].get("slot_mode"))
    self.cuda = _ps._use_cuda(ctx)          # where THIS worker's tensors live
    me = [ctx.job_name, ctx.task_index]
    self.client_id = parts[0]["clients"].index(me) if me in parts[0]["clients"] else 0
    self.num_clients = len(parts[0]["clients"])
    self.pushes = 0
    self.socks, self._pending, self._bufs = [], [], []
    for p in parts:
      s = socket.create_connection((p["host"], p["port"]), timeout=timeout)
      s.settimeout(None)
      s.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
      self.socks.append(s)
      self._pending.append(0)
      self._bufs.append(None)
    self._running_pulled = None
    self._host_grads = None

  # ------------------------------------------------------------------ plumbing
  def _drain(self, i):
    while self._pending[i]:
      _, _, _, _, _, self._bufs[i] = _recv(self.socks[i], self._bufs[i])
      self._pending[i] -= 1

  def _call(self, i, op, seq=0, scalar=0.0, payload=b""):
    self._drain(i)
    _send(self.socks[i], op, self.client_id, seq, scalar, payload)
    _, _, _, applied, view, self._bufs[i] = _recv(self.socks[i], self._bufs[i])
    return view, applied

  @staticmethod
  def _host(x):
    if hasattr(x, "detach"):
      return x.detach().float().cpu().numpy().reshape(-1)
    return np.asarray(x, dtype=np.float32).reshape(-1)

  # ------------------------------------------------------------------ plain mode
  def pull(self, out_fp32=None, out_bf16=None):
    flat = np.empty(self.numel, dtype=np.float32)
    for i, p in enumerate(self.parts):
      view, _ = self._call(i, OP_PULL)
      flat[p["lo"]:p["hi"]] = np.frombuffer(view, dtype=np.float32)
    import torch
    t = torch.from_numpy(flat)
    if out_fp32 is not None:
      if hasattr(out_fp32, "copy_"):
        out_fp32.copy_(t)
      else:
        out_fp32[:] = flat
    if out_bf16 is not None:
      out_bf16.copy_(t)
    if self.cuda and (out_fp32 is not None or out_bf16 is not None):
      return out_fp32 if out_fp32 is not None else out_bf16
    return flat

  def push(self, grad, lr, scale=1.0):
    g = self._host(grad)
    for i, p in enumerate(self.parts):
      self._call(i, OP_PUSH_DENSE, scalar=float(lr) * float(scale),
                 payload=np.ascontiguousarray(g[p["lo"]:p["hi"]]))

  def push_sparse(self, grad_rows, indices, width, base=0, lr=1.0, scale=1.0):
    g = (grad_rows.detach().float().cpu().numpy() if hasattr(grad_rows, "detach")
         else np.asarray(grad_rows, dtype=np.float32)).reshape(-1, width)
    ix = (indices.detach().cpu().numpy() if hasattr(indices, "detach") else np.asarray(indices)
          ).astype(np.int64).reshape(-1)
    first = base + ix * width
    for i, p in enumerate(self.parts):
      sel = (first < p["hi"]) & (first + width > p["lo"])   # rows that touch this server's slice
      if not sel.any():
        continue
      rows = np.ascontiguousarray(g[sel])
      payload = SPARSE.pack(rows.shape[0], width, base) + ix[sel].tobytes() + rows.tobytes()
      self._call(i, OP_PUSH_SPARSE, scalar=float(lr) * float(scale), payload=payload)

  # ------------------------------------------------------------------ slot mode
  def set_lr(self, lr):
    for i in range(len(self.parts)):
      self._call(i, OP_SET_LR, scalar=float(lr))

  def pull_model(self, weights=None, aux32=None, running=None, decay_end=None, total=None):
    """GPU worker: fills bf16 ``weights`` [total], fp32 ``aux32`` (= [decay_end, total)) and the
    ``running`` statistics tail; CPU worker: returns the flat fp32 vector."""
    if weights is None:
      return self.pull()
    import torch
    total = int(total if total is not None else weights.numel())
    decay_end = int(decay_end if decay_end is not None else total)
    dev = weights.device
    if running is not None and (self._running_pulled is None
                                or self._running_pulled.numel() != running.numel()):
      self._running_pulled = torch.zeros_like(running)
    for i, p in enumerate(self.parts):
      lo, hi = p["lo"], p["hi"]
      view, _ = self._call(i, OP_PULL_MODEL, seq=decay_end)
      a = min(max(decay_end - lo, 0), hi - lo)
      if a:
        head = torch.from_numpy(np.frombuffer(view[:2 * a], dtype=np.int16).copy()).view(torch.bfloat16)
        weights[lo:lo + a].copy_(head.to(dev, non_blocking=True))
      tail = torch.from_numpy(np.frombuffer(view[2 * a:], dtype=np.float32).copy()).to(dev)
      t0 = lo + a                                       # global index of tail[0]
      m = min(hi, total)                                # model part of the tail: [t0, m)
      if m > t0:
        weights[t0:m].copy_(tail[:m - t0])
        if aux32 is not None:
          aux32[t0 - decay_end:m - decay_end].copy_(tail[:m - t0])
      if hi > total and running is not None:            # running statistics: [max(t0, total), hi)
        r0 = max(t0, total)
        running[r0 - total:hi - total].copy_(tail[r0 - t0:])
        self._running_pulled[r0 - total:hi - total].copy_(tail[r0 - t0:])

  def push_grads(self, grads, running=None, total=None):
    """Hand this step's gradients (flat fp32, length ``total``) to the servers; the non-trainable
    tail, if any, is sent as (pulled - current) running statistics.  Returns without waiting for
    the apply: the acknowledgement is collected before the next request on the connection."""
    self.pushes += 1
    if hasattr(grads, "detach") and grads.is_cuda:
      import torch
      total = int(total if total is not None else grads.numel())
      if self._host_grads is None or self._host_grads.numel() != self.numel:
        self._host_grads = torch.zeros(self.numel, dtype=torch.float32).pin_memory()
      self._host_grads[:total].copy_(grads[:total], non_blocking=True)
      if running is not None and self.numel > total:
        self._host_grads[total:].copy_(self._running_pulled - running, non_blocking=True)
      torch.cuda.current_stream().synchronize()
      g = self._host_grads.numpy()
    else:
      g = self._host(grads)
      if g.size < self.numel:
        tail = np.zeros(self.numel - g.size, dtype=np.float32)
        if running is not None and self._running_pulled is not None:
          tail = self._host(self._running_pulled) - self._host(running)
        g = np.concatenate([g, tail])
    for i, p in enumerate(self.parts):
      self._drain(i)                                    # back-pressure: push i - 1 was applied
      _send(self.socks[i], OP_PUSH_GRADS, self.client_id, self.pushes, 0.0,
            np.ascontiguousarray(g[p["lo"]:p["hi"]]))
      self._pending[i] += 1

  def applies(self):
    """Number of requests each server has applied so far (drains outstanding pushes first)."""
    return [struct.unpack("<q", bytes(self._call(i, OP_STATS)[0]))[0] for i in range(len(self.parts))]

  def close(self):
    for i, s in enumerate(self.socks):
      try:
        self._drain(i)
        _send(s, OP_BYE)
      except Exception:
        pass
      s.close()
    self.socks = []
