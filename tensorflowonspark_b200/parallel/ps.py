"""Asynchronous parameter server on GPU (and a shared-memory twin for CPU executors).

The reference reaches this through TensorFlow: ``tf.train.Server`` between-graph replication in
TF1 (tensorflowonspark/TFNode.py:126-132) or ``ParameterServerStrategy`` in TF2 (examples/mnist/
estimator/mnist_spark_streaming.py:86); TFoS itself only reserves ``num_ps`` executors for the
'ps' role and parks them on a control queue (TFCluster.py:260-262, TFSparkNode.py:442-458).

Here a 'ps' node owns a slice of the flat fp32 parameter vector **resident on its GPU**,
publishes the CUDA IPC handle on the reservation board, and sleeps.  Workers map that memory
and, without any barrier (Hogwild-style asynchronous SGD):

  * ``pull()``  - ``ps_pull`` kernel: peer loads of the parameters over NVLink, writing the local
    fp32 copy and the bf16 compute copy in one pass;
  * ``push()``  - ``ps_push_dense`` kernel: ``w_ps += -lr * scale * g`` applied *in the PS GPU's
    memory* with remote ``red.add.f32`` - the gradient never lands in a staging buffer;
  * ``push_sparse()`` - row-indexed scatter-apply for embedding-style (IndexedSlices) gradients.

On CPU-only executors the same interface is backed by POSIX shared memory + numpy (used by the
CPU test-suite and BASELINE config #1 style runs).
"""
import logging
import time

import numpy as np

logger = logging.getLogger(__name__)


def _use_cuda(ctx):
  try:
    import torch
    return bool(getattr(ctx, "gpus", None)) and torch.cuda.is_available()
  except Exception:
    return False


def _board(ctx):
  from .. import reservation
  return reservation.Client(ctx.server_addr)


def _ps_count(ctx):
  return len(ctx.cluster_spec.get("ps", []))


def _slices(numel, num_ps):
  per = ((numel + num_ps - 1) // num_ps + 7) // 8 * 8
  return [(min(numel, i * per), min(numel, (i + 1) * per)) for i in range(num_ps)]


class PSServer(object):
  """Runs inside a 'ps' node: hosts parameters [lo, hi) of the flat vector."""

  def __init__(self, ctx, numel, init=None):
    self.ctx = ctx
    self.numel = int(numel)
    self.lo, self.hi = _slices(self.numel, _ps_count(ctx))[ctx.task_index]
    n = self.hi - self.lo
    key = "ps/{}/{}".format(ctx.cluster_id, ctx.task_index)
    board = _board(ctx)
    self.cuda = _use_cuda(ctx)
    if self.cuda:
      import torch
      from .. import ops
      torch.cuda.set_device(0)
      ptr, handle = ops.C().symm_alloc(max(256, n * 4))
      self.tensor = ops.C().tensor_from_ptr(ptr, [n], "f32")
      if init is not None:
        self.tensor.copy_(torch.as_tensor(init, dtype=torch.float32).reshape(-1)[self.lo:self.hi])
      torch.cuda.synchronize()
      board.put(key, {"kind": "cuda", "handle": handle, "lo": self.lo, "hi": self.hi,
                      "numel": self.numel})
    else:
      from multiprocessing import shared_memory
      self.shm = shared_memory.SharedMemory(create=True, size=max(8, n * 4))
      self.array = np.ndarray((n,), dtype=np.float32, buffer=self.shm.buf)
      self.array[:] = 0 if init is None else np.asarray(init, dtype=np.float32).reshape(-1)[
          self.lo:self.hi]
      board.put(key, {"kind": "shm", "name": self.shm.name, "lo": self.lo, "hi": self.hi,
                      "numel": self.numel})
    board.close()
    logger.info("ps:%d serving parameters [%d, %d) on %s", ctx.task_index, self.lo, self.hi,
                "GPU" if self.cuda else "shared memory")

  def values(self):
    if self.cuda:
      return self.tensor.detach().cpu().numpy().copy()
    return self.array.copy()

  def serve_forever(self, poll=1.0):
    """Park until the driver stops this node (the node runtime kills the process)."""
    while True:
      time.sleep(poll)

  def close(self):
    if not self.cuda:
      try:
        self.shm.close()
        self.shm.unlink()
      except Exception:
        pass


class PSClient(object):
  """Worker-side handle on all parameter servers of the cluster."""

  def __init__(self, ctx, timeout=600):
    self.ctx = ctx
    board = _board(ctx)
    self.parts = []
    for i in range(_ps_count(ctx)):
      info = board.get("ps/{}/{}".format(ctx.cluster_id, i), timeout)
      self.parts.append(info)
    board.close()
    self.numel = self.parts[0]["numel"]
    self.cuda = self.parts[0]["kind"] == "cuda"
    if self.cuda:
      import torch
      from .. import ops
      torch.cuda.set_device(0)
      self.ptrs = [ops.C().symm_open(p["handle"]) for p in self.parts]
      self.hyper = torch.zeros(8, dtype=torch.float32, device="cuda")
    else:
      from multiprocessing import shared_memory
      self.shms = [shared_memory.SharedMemory(name=p["name"]) for p in self.parts]
      self.arrays = [np.ndarray((p["hi"] - p["lo"],), dtype=np.float32, buffer=s.buf)
                     for p, s in zip(self.parts, self.shms)]

  # ------------------------------------------------------------------ pull
  def pull(self, out_fp32=None, out_bf16=None):
    """Fetch the current parameters.  GPU: fills the given device tensors (either may be None);
    CPU: returns a numpy copy (and fills ``out_fp32`` if it is a numpy array / torch tensor)."""
    if self.cuda:
      from .. import ops
      for p, ptr in zip(self.parts, self.ptrs):
        lo, hi = p["lo"], p["hi"]
        ops.K.ps_pull(ptr, out_fp32[lo:hi] if out_fp32 is not None else None,
                      out_bf16[lo:hi] if out_bf16 is not None else None, hi - lo)
      return out_fp32 if out_fp32 is not None else out_bf16
    flat = np.concatenate([a for a in self.arrays]) if len(self.arrays) > 1 else self.arrays[0].copy()
    if out_fp32 is not None:
      if hasattr(out_fp32, "numpy"):
        out_fp32.copy_(__import__("torch").from_numpy(flat))
      else:
        out_fp32[:] = flat
    return flat

  # ------------------------------------------------------------------ push
  def push(self, grad, lr, scale=1.0):
    """Apply ``w -= lr * scale * grad`` on the servers, asynchronously and without locking."""
    if self.cuda:
      from .. import ops
      self.hyper[0] = float(lr)
      self.hyper[3] = float(scale)
      for p, ptr in zip(self.parts, self.ptrs):
        ops.K.ps_push_dense(ptr, grad[p["lo"]:p["hi"]], self.hyper)
      return
    g = grad.detach().cpu().numpy() if hasattr(grad, "detach") else np.asarray(grad)
    g = g.reshape(-1).astype(np.float32)
    for p, a in zip(self.parts, self.arrays):
      a -= (lr * scale) * g[p["lo"]:p["hi"]]

  def push_sparse(self, grad_rows, indices, width, base=0, lr=1.0, scale=1.0):
    """Row-sparse update of a [rows, width] table stored at flat offset ``base``.  A row that
    straddles a server boundary is split: each server applies the elements it owns."""
    if self.cuda:
      import torch
      from .. import ops
      self.hyper[0] = float(lr)
      self.hyper[3] = float(scale)
      idx = indices.to(torch.int32)
      for p, ptr in zip(self.parts, self.ptrs):
        lo, hi = p["lo"], p["hi"]
        r0, r1 = max(0, (lo - base + width - 1) // width), max(0, (hi - base) // width)
        sel = (idx >= r0) & (idx < r1)
        if r1 > r0 and bool(sel.any()):
          rows = grad_rows[sel].contiguous()
          local = (idx[sel] - r0).contiguous()
          off = base + r0 * width - lo
          ops.K.ps_push_sparse(ptr + 4 * off, rows, local, self.hyper)
        # rows cut by this server's boundaries: the owned element range goes through the dense kernel
        for edge in (lo, hi):
          r = (edge - base) // width
          s0 = base + r * width
          if edge <= base or s0 == edge or s0 + width <= lo or s0 >= hi:
            continue
          a, b = max(s0, lo), min(s0 + width, hi)
          for k in torch.nonzero(idx == r).flatten().tolist():
            seg = grad_rows[k].reshape(-1)[a - s0:b - s0].to(torch.float32).contiguous()
            ops.K.ps_push_dense(ptr + 4 * (a - lo), seg, self.hyper)
      return
    g = grad_rows.detach().cpu().numpy() if hasattr(grad_rows, "detach") else np.asarray(grad_rows)
    ix = indices.detach().cpu().numpy() if hasattr(indices, "detach") else np.asarray(indices)
    for p, a in zip(self.parts, self.arrays):
      for row, r in zip(g, ix):
        s = base + int(r) * width
        b0, b1 = max(s, p["lo"]), min(s + width, p["hi"])   # the part of the row this server owns
        if b0 < b1:
          a[b0 - p["lo"]:b1 - p["lo"]] -= (lr * scale) * row[b0 - s:b1 - s]

  def close(self):
    if self.cuda:
      from .. import ops
      for ptr in self.ptrs:
        try:
          ops.C().symm_close(ptr)
        except Exception:
          pass
    else:
      for s in self.shms:
        s.close()


def attach(ctx, params=None, init=None):
  """PSServer on a 'ps' node (``params`` = number of elements or an initial flat tensor),
  PSClient everywhere else."""
  if ctx.job_name == "ps":
    if params is None:
      raise ValueError("a ps node must be told the parameter count: start_cluster_server(params=N)")
    if hasattr(params, "numel") or isinstance(params, np.ndarray):
      init = params
      numel = int(params.numel() if hasattr(params, "numel") else params.size)
    else:
      numel = int(params)
    return PSServer(ctx, numel, init)
  return PSClient(ctx)
