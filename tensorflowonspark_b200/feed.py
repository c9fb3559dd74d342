"""Host -> device input pipeline: pinned staging + side-stream copies.

``DevicePrefetcher`` is the device half of the InputMode.SPARK fast path: batches
arrive in page-locked host memory (either plain pinned tensors or slots of the
shared-memory ring written by feeder tasks, csrc/feed.cc), are copied with
``cudaMemcpyAsync`` on a dedicated copy stream into a small pool of device
staging buffers, and are handed to the compute stream through CUDA events - the
copy of batch i+1 overlaps the training step of batch i.

The reference has no equivalent: rows cross two process boundaries one pickled
RPC at a time and reach TF through ``tf.data.Dataset.from_generator``
(tensorflowonspark/TFNode.py:278-300, examples/mnist/keras/mnist_spark.py:33-47).
"""
import collections

import torch


class DevicePrefetcher(object):

  def __init__(self, specs, device, depth=2):
    """specs: list of (shape, dtype) for the tensors of one batch."""
    self.device = torch.device(device)
    self.depth = depth
    self.copy_stream = torch.cuda.Stream(device=self.device)
    self.slots = [[torch.empty(s, dtype=dt, device=self.device) for (s, dt) in specs]
                  for _ in range(depth)]
    self.ready = [torch.cuda.Event() for _ in range(depth)]
    self.consumed = [torch.cuda.Event() for _ in range(depth)]
    self._used = [False] * depth
    self._w = 0
    self._queue = collections.deque()
    self.bytes_per_batch = sum(int(torch.empty(s, dtype=dt).numel()) * torch.empty(0, dtype=dt)
                               .element_size() for (s, dt) in specs)
    self._specs = specs
    self._host = None   # pinned staging owned by the prefetcher (push_arrays)
    self.direct_copies = self.staged_copies = 0   # tensors DMA'd in place / through staging

  def push_arrays(self, arrays):
    """Stage one batch given as numpy arrays / CPU tensors (e.g. zero-copy views of a feed-ring
    slot from ``DataFeed.next_batch_arrays``) and enqueue its async copy.

    The page-locked staging buffers belong to the prefetcher and are reused round-robin; before
    one is overwritten the host WAITS for the copy that last read it.  That wait is the
    back-pressure of the input pipeline: the host can run at most ``depth`` batches ahead of the
    device, however long the device is stalled (e.g. on a slower peer in the all-reduce) - an
    unprotected staging buffer would silently feed the wrong batch."""
    if self._host is None:
      self._host = [[torch.empty(s, dtype=dt).pin_memory() for (s, dt) in self._specs]
                    for _ in range(self.depth)]
    i = self._w
    if self._used[i]:
      self.ready[i].synchronize()
    import numpy as np
    sources = []
    for dst, src, (shape, dt) in zip(self._host[i], arrays, self._specs):
      a = src.numpy() if isinstance(src, torch.Tensor) else np.asarray(src)
      # a view of a page-locked region (the feed ring after DataFeed.pin_ring(), or any pinned
      # tensor) of the right dtype is a DMA source as it is: no staging copy
      if a.flags["C_CONTIGUOUS"] and a.size == dst.numel():
        try:
          t = torch.from_numpy(a.reshape(dst.shape))
          if t.dtype == dt and t.is_pinned():
            sources.append(t)
            self.direct_copies += 1
            continue
        except (TypeError, ValueError, RuntimeError):
          pass
      # plain single-threaded memcpy (or cast): torch's copy_ fans a 3 MB copy out over the
      # whole intra-op pool, which crawls when several node processes share a CPU quota
      np.copyto(dst.numpy(), a.reshape(dst.shape), casting="unsafe")
      sources.append(dst)
      self.staged_copies += 1
    self.push(sources)

  def acquire_host(self):
    """The page-locked staging tensors of the NEXT push, for callers that assemble a batch in
    place (row by row) instead of handing over ready arrays; waits for the copy that last read
    them (same back-pressure as :meth:`push_arrays`).  Follow with :meth:`push_host`."""
    if self._host is None:
      self._host = [[torch.empty(s, dtype=dt).pin_memory() for (s, dt) in self._specs]
                    for _ in range(self.depth)]
    i = self._w
    if self._used[i]:
      self.ready[i].synchronize()
    return self._host[i]

  def push_host(self):
    self.staged_copies += len(self._specs)
    self.push(self._host[self._w])

  def push(self, host_tensors):
    """Enqueue an async copy of one batch (pinned host tensors) into the next staging slot."""
    i = self._w
    self._w = (self._w + 1) % self.depth
    with torch.cuda.stream(self.copy_stream):
      if self._used[i]:
        self.copy_stream.wait_event(self.consumed[i])
      for dst, src in zip(self.slots[i], host_tensors):
        dst.copy_(src, non_blocking=True)
      self.ready[i].record(self.copy_stream)
    self._used[i] = True
    self._queue.append(i)

  def push_ring_slot(self, ring, pos, layout):
    """Same, straight from a pinned shared-memory ring slot.  layout: [(offset, nbytes)] per tensor."""
    i = self._w
    self._w = (self._w + 1) % self.depth
    with torch.cuda.stream(self.copy_stream):
      if self._used[i]:
        self.copy_stream.wait_event(self.consumed[i])
      for dst, (off, nbytes) in zip(self.slots[i], layout):
        ring.h2d(pos, off, dst.data_ptr(), nbytes, self.copy_stream.cuda_stream)
      self.ready[i].record(self.copy_stream)
    self._used[i] = True
    self._queue.append(i)

  def pop(self):
    """Device tensors of the oldest pushed batch; the current stream waits for its copy."""
    i = self._queue.popleft()
    torch.cuda.current_stream(self.device).wait_event(self.ready[i])
    self._last = i
    return self.slots[i]

  def release(self):
    """Mark the batch returned by the last pop() as consumed (after the kernels reading it are enqueued)."""
    self.consumed[self._last].record(torch.cuda.current_stream(self.device))

  def pending(self):
    return len(self._queue)
