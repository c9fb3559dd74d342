"""In-process helpers for the user's ``map_fun`` (reference: tensorflowonspark/TFNode.py).

* :class:`DataFeed` - the consumer end of InputMode.SPARK (reference TFNode.py:234-342), same
  ``next_batch / should_stop / batch_results / terminate`` contract, plus a device fast path
  (``next_batch_tensors``) that moves ring blocks pinned-host -> GPU on a copy stream.
* :func:`hdfs_path` - path normalisation against the cluster's default filesystem (:32-67).
* :func:`start_cluster_server` - where the reference started a ``tf.train.Server`` (TF1 only,
  :70-154), this returns a :class:`ClusterServer` handle: a NCCL/gloo process group for sync
  data parallelism or a GPU-resident parameter server for the async path.
* :func:`export_saved_model`, :func:`release_port`, and the deprecated module-level stubs.
"""
import getpass
import logging
import os
import queue as _queue_mod

from . import marker

logger = logging.getLogger(__name__)

#: schemes understood by Hadoop-compatible filesystems; paths carrying one are already absolute
HADOOP_SCHEMES = ("adl://", "file://", "hdfs://", "oss://", "s3://", "s3a://", "s3n://", "swift://",
                  "viewfs://", "wasb://")


def hdfs_path(ctx, path):
  """Absolute, scheme-qualified version of ``path`` for this cluster's default filesystem."""
  if path.startswith(HADOOP_SCHEMES):
    return path
  fs = ctx.defaultFS
  if path.startswith("/"):
    return fs + path
  if fs.startswith(("hdfs://", "viewfs://")):
    return "{}/user/{}/{}".format(fs, getpass.getuser(), path)
  if fs.startswith("file://"):
    return "{}/{}/{}".format(fs, ctx.working_dir[1:], path)
  logger.warning("Unknown scheme %s with relative path: %s", fs, path)
  return "{}/{}".format(fs, path)


def local_path(path):
  """Strip a ``file://`` scheme so the path can be opened with ordinary file APIs."""
  return path[len("file://"):] if path.startswith("file://") else path


class ClusterServer(object):
  """What :func:`start_cluster_server` hands back in place of a ``tf.train.Server``.

  ``target`` names the rendezvous (``tcp://host:port``) the way ``server.target`` named the gRPC
  endpoint; ``join()`` parks a ps node until the driver stops it, like ``server.join()``.
  """

  def __init__(self, ctx, group=None, ps=None):
    self.ctx, self.group, self.ps = ctx, group, ps
    self.target = "tcp://{}:{}".format(os.environ.get("MASTER_ADDR", "127.0.0.1"),
                                       os.environ.get("MASTER_PORT", "0"))

  def join(self):
    if self.ps is not None:
      self.ps.serve_forever()
    else:
      import time
      while True:
        time.sleep(3600)


def start_cluster_server(ctx, num_gpus=1, rdma=False, backend=None, async_ps=None, params=None,
                         **ps_kwargs):
  """Join this node to the cluster's communication substrate.

  Returns ``(cluster_spec, server)`` like the reference.  With ``num_ps == 0`` every worker
  enters a ``torch.distributed`` process group (NCCL when a GPU is visible, gloo otherwise);
  when the cluster has ps nodes, ps processes host the parameters on their GPU and workers get
  push/pull handles (parallel/ps.py).  ``rdma`` is accepted for signature compatibility: NVLink
  peer access supersedes the reference's ``grpc+verbs`` switch (TFNode.py:129-130).
  """
  from .parallel import process_group
  has_ps = bool(ctx.cluster_spec.get("ps"))
  if async_ps is None:
    async_ps = has_ps
  if async_ps:
    from .parallel import ps as ps_mod
    # ps nodes need the parameter count (or initial values); workers just attach
    # (``optimizer=...`` + hyper-parameters: the ps keeps the optimizer state - "slot mode")
    handle = ps_mod.attach(ctx, params=params, **ps_kwargs)
    return ctx.cluster_spec, ClusterServer(ctx, ps=handle)
  group = process_group.init_from_ctx(ctx, backend=backend)
  return ctx.cluster_spec, ClusterServer(ctx, group=group)


def export_saved_model(sess, export_dir, tag_set="serve", signatures=None):
  """Write an inference artefact to ``export_dir`` (weights + a JSON signature).

  The reference's version (TFNode.py:162-211) wrote a TF1 SavedModel from a session; here the
  first argument is any object with ``state_dict()`` (a native-engine model or a torch module).
  ``signatures`` maps signature keys to ``{'inputs': {alias: tensor_name}, 'outputs': {...}}``.
  """
  from .utils import checkpoint
  return checkpoint.export_model(sess, export_dir, tag_set=tag_set, signatures=signatures)


def release_port(ctx):
  """Close the socket that has been holding this node's reserved port (reference :214-221)."""
  if getattr(ctx, "tmp_socket", None) is not None:
    logger.info("releasing reserved port %s", ctx.tmp_socket.getsockname())
    ctx.tmp_socket.close()
    ctx.tmp_socket = None
  else:
    logger.warning("release_port() called but no port is being held")


def next_batch(mgr, batch_size, qname="input"):
  """*DEPRECATED*: use :class:`DataFeed`."""
  raise Exception("DEPRECATED: Use TFNode.DataFeed class instead")


def batch_results(mgr, results, qname="output"):
  """*DEPRECATED*: use :class:`DataFeed`."""
  raise Exception("DEPRECATED: Use TFNode.DataFeed class instead")


def terminate(mgr, qname="input"):
  """*DEPRECATED*: use :class:`DataFeed`."""
  raise Exception("DEPRECATED: Use TFNode.DataFeed class instead")


def _value(proxy):
  return proxy._getvalue() if hasattr(proxy, "_getvalue") else proxy


class DataFeed(object):
  """Consumer of the rows Spark feeds into this node.

  Args:
    mgr: this executor's TFManager.
    train_mode: True when feeding training data (no results are returned to Spark).
    qname_in / qname_out: queue names.
    input_mapping: ``{column: tensor_name}``; when given, batches are dicts
      ``{tensor_name: [column values]}`` with columns taken in ``sorted(input_mapping)`` order,
      otherwise batches are lists of rows.
  """

  def __init__(self, mgr, train_mode=True, qname_in="input", qname_out="output",
               input_mapping=None):
    self.mgr = mgr
    self.train_mode = train_mode
    self.qname_in, self.qname_out = qname_in, qname_out
    self.done_feeding = False
    self.input_tensors = [t for _, t in sorted(input_mapping.items())] if input_mapping else None
    self.queue_in = mgr.get_queue(qname_in)
    self.queue_out = mgr.get_queue(qname_out) if not train_mode else None
    self._block, self._off = None, 0   # block currently being drained
    self._ring = None
    self._prefetch = None
    self._held_pos, self._retired, self._calls = None, [], 0   # ring slots not yet given back
    self._views_out, self._last_take_call = False, 0

  # ------------------------------------------------------------- internals
  def _attach_ring(self):
    if self._ring is None:
      info = _value(self.mgr.get("ring"))
      if not info:
        raise RuntimeError("feeder posted a ring block but no ring is registered")
      from . import shmring
      self._ring = shmring.attach(info["name"])
      if getattr(self, "_want_pin", False):
        try:
          self._ring.pin()
          logger.info("feed ring %s page-locked: batches are DMA'd straight from its slots",
                      info["name"])
        except Exception as e:
          logger.warning("could not page-lock the feed ring: %s", e)
          self._want_pin = False
    return self._ring

  def pin_ring(self):
    """Page-lock the shared-memory feed ring in THIS process (``cudaHostRegister``), so that the
    zero-copy views returned by :meth:`next_batch_arrays` are DMA sources: the prefetcher then
    issues ``cudaMemcpyAsync`` straight from the ring slot and skips its staging copy
    (feed.DevicePrefetcher.push_arrays).  A slot is therefore kept one call longer (3 instead
    of 2): it may only go back to the feeders after the copy that reads it has completed, which
    the prefetcher's back-pressure guarantees two pushes later.  Returns True when pinned."""
    try:
      import torch
      if not torch.cuda.is_available():
        return False
      self._want_pin = True
      self.HOLD_CALLS = max(self.HOLD_CALLS, 3)
      if _value(self.mgr.get("ring")):
        self._attach_ring()
      return True
    except Exception as e:   # an unpinned ring only costs the staging copy
      logger.warning("could not page-lock the feed ring: %s", e)
      self._want_pin = False
      return False

  def _expand(self, item):
    """Turn one queue item into a block: ('rows', list) or ('cols', [ndarray per column], tupled).

    Ring blocks stay in their shared-memory slot: the columns are zero-copy views, and the slot
    goes back to the feeders ``HOLD_CALLS`` ``next_batch*`` calls after its last row was handed
    out (see :meth:`_retire`) - so a batch returned by ``next_batch_arrays`` stays valid while
    the caller stages / copies it, and nothing is copied just to free the slot."""
    if isinstance(item, marker.RingBlock):
      ring = self._attach_ring()
      from . import shmring
      cols = shmring.unpack_columns(ring, item)
      tupled = not (len(item.layout) == 1 and len(item.layout[0]) == 5)
      self._held_pos = item.pos
      return ("cols", cols, tupled, item.nrows)
    if isinstance(item, marker.Rows):
      return ("rows", item.rows, True, len(item.rows))
    return ("rows", [item], True, 1)

  def _pull(self):
    """One blocking queue read.  Returns 'eof', 'end_partition' or 'rows'."""
    item = self.queue_in.get(block=True)
    if item is None:
      logger.info("next_batch() got None: end of feed")
      self.queue_in.task_done()
      self.done_feeding = True
      self._retire_current()
      return "eof"
    if isinstance(item, marker.EndPartition):
      self.queue_in.task_done()
      self._retire_current()
      return "end_partition"
    self._retire_current()
    self._block, self._off = self._expand(item), 0
    self.queue_in.task_done()
    return "rows"

  HOLD_CALLS = 2   # a returned batch stays valid for this many further next_batch* calls

  def _retire_current(self):
    """The block being drained is exhausted: queue its ring slot for release."""
    pos = getattr(self, "_held_pos", None)
    if pos is not None:
      self._held_pos = None
      if self._views_out:       # the caller may still be reading views of this slot
        self._retired.append((pos, self._last_take_call))
      else:                     # only copies (python rows) left the slot: free it right away
        try:
          self._attach_ring().release_read(pos)
        except Exception:
          pass
    self._views_out = False

  def _release_retired(self, everything=False):
    while self._retired and (everything or self._calls - self._retired[0][1] >= self.HOLD_CALLS):
      pos, _ = self._retired.pop(0)
      try:
        self._attach_ring().release_read(pos)
      except Exception:
        pass

  def _remaining(self):
    return 0 if self._block is None else self._block[3] - self._off

  def _take_rows(self, k):
    kind, data, tupled, _ = self._block
    lo, hi = self._off, self._off + k
    self._off = hi
    if kind == "rows":
      return data[lo:hi]
    if not tupled:
      return data[0][lo:hi].tolist()
    lists = [c[lo:hi].tolist() for c in data]
    return [list(r) for r in zip(*lists)]

  def _take_arrays(self, k):
    import numpy as np
    kind, data, tupled, _ = self._block
    lo, hi = self._off, self._off + k
    self._off = hi
    if kind == "cols":
      self._views_out = self._held_pos is not None
      self._last_take_call = self._calls   # the hold period counts from the last hand-out
      return [c[lo:hi] for c in data]
    rows = data[lo:hi]
    if rows and isinstance(rows[0], (list, tuple)):
      return [np.asarray([r[c] for r in rows]) for c in range(len(rows[0]))]
    return [np.asarray(rows)]

  def _fill(self, batch_size, take):
    """Common batching loop: ``take(k)`` pops k rows of the current block."""
    self._calls += 1
    self._release_retired()
    parts, count = [], 0
    while count < batch_size:
      if self._remaining() == 0:
        if self.done_feeding:
          break
        what = self._pull()
        if what == "eof":
          break
        if what == "end_partition":
          if not self.train_mode and count > 0:
            break
          continue
      k = min(self._remaining(), batch_size - count)
      parts.append(take(k))
      count += k
    return parts

  # ------------------------------------------------------------------- API
  def next_batch(self, batch_size):
    """Up to ``batch_size`` rows (fewer at end of feed / end of an inference partition)."""
    rows = [r for part in self._fill(batch_size, self._take_rows) for r in part]
    if self.input_tensors is None:
      return rows
    tensors = {t: [] for t in self.input_tensors}
    for row in rows:
      for i, t in enumerate(self.input_tensors):
        tensors[t].append(row[i])
    return tensors

  def next_batch_arrays(self, batch_size):
    """Column-major fast path: a list with one numpy array per column (``[n, ...]`` each), or a
    dict keyed by tensor name when ``input_mapping`` was given.  Ring blocks are sliced, never
    expanded into python objects."""
    import numpy as np
    parts = self._fill(batch_size, self._take_arrays)
    if not parts:
      cols = []
    else:
      ncol = len(parts[0])
      cols = [parts[0][c] if len(parts) == 1 else np.concatenate([p[c] for p in parts])
              for c in range(ncol)]
    if self.input_tensors is None:
      return cols
    return dict(zip(self.input_tensors, cols))

  def should_stop(self):
    """True once the end-of-feed marker has been consumed."""
    return self.done_feeding and self._remaining() == 0

  def batch_results(self, results):
    """Return one result per input row of the last batch to Spark (inference mode)."""
    results = list(results)
    self.queue_out.put(marker.Rows(results), block=True)

  def terminate(self):
    """Stop consuming: flag the executor as terminating and drain whatever is still queued."""
    logger.info("terminate() invoked")
    self.mgr.set("state", "terminating")
    self._retire_current()
    self._release_retired(everything=True)
    self._block, self._off = None, 0
    dropped = 0
    while True:
      try:
        item = self.queue_in.get(block=True, timeout=5)
        if isinstance(item, marker.RingBlock):
          try:
            self._attach_ring().release_read(item.pos)
          except Exception:
            pass
        self.queue_in.task_done()
        dropped += 1
      except _queue_mod.Empty:
        break
    logger.info("dropped %d queued item(s)", dropped)
    # Tell the driver right away.  (The reference only relays this through the *next* feeder
    # task, tensorflowonspark/TFSparkNode.py:520-531, so a stream that goes quiet never stops.)
    try:
      addr = _value(self.mgr.get("server_addr"))
      if addr:
        from . import reservation
        client = reservation.Client(tuple(addr))
        client.request_stop()
        client.close()
    except Exception as e:
      logger.debug("could not notify the reservation server: %s", e)

  # -------------------------------------------------------- device fast path
  def next_batch_tensors(self, batch_size, device=None, dtypes=None):
    """Like :meth:`next_batch` but returns one device tensor per column.

    Rows are staged in page-locked host memory and copied with ``cudaMemcpyAsync`` on a side
    stream (feed.DevicePrefetcher); on a CPU-only host plain tensors are returned.
    """
    import numpy as np
    import torch
    batch = self.next_batch_arrays(batch_size)
    if self.input_tensors is None:
      names, cols = list(range(len(batch))), batch
    else:
      names = self.input_tensors
      cols = [batch[t] for t in names]
    use_cuda = torch.cuda.is_available() and (device is None or str(device).startswith("cuda"))
    arrs = [np.ascontiguousarray(np.asarray(c, dtype=(dtypes[i] if dtypes else None)))
            for i, c in enumerate(cols)]
    if not use_cuda:
      out = [torch.from_numpy(a) for a in arrs]
      return out if self.input_tensors is None else dict(zip(names, out))
    # page-locked staging and the device slots belong to a prefetcher that lives as long as the
    # feed (re-created only if the batch geometry changes): no cudaHostAlloc / page-lock per batch
    from .feed import DevicePrefetcher
    specs = [(tuple(a.shape), torch.from_numpy(a[:0]).dtype) for a in arrs]
    pf = getattr(self, "_prefetcher", None)
    if pf is None or self._prefetch_specs != specs:
      pf = self._prefetcher = DevicePrefetcher(specs, device or "cuda", depth=3)
      self._prefetch_specs = specs
    elif getattr(self, "_prefetch_unreleased", False):
      # the kernels that read the PREVIOUS batch have been enqueued on the current stream by now:
      # only from this point on may a later copy reuse that batch's device slot
      pf.release()
    pf.push_arrays(arrs)
    slots = pf.pop()
    self._prefetch_unreleased = True
    out = list(slots)   # valid until the caller asks for the batch after the next one
    return out if self.input_tensors is None else dict(zip(names, out))
