"""Device-side timing and profiler hooks.

The reference has no tracing beyond launching TensorBoard.  Here:
* :class:`DeviceTimer` - CUDA-event phase timer (the only timing BASELINE.json accepts: on the
  device, on the launching stream, max over ranks);
* :func:`nvtx_range` - NVTX ranges around step phases for Nsight tools;
* :func:`exposed_ms` - how long a stream had to wait for another one (non-overlapped collective time).
"""
import contextlib


class DeviceTimer(object):
  """``with t.phase('fwd'): ...`` records start/stop events; ``t.summary()`` -> {phase: ms}."""

  def __init__(self, device=None):
    import torch
    self.torch = torch
    self.device = device
    self.records = {}

  @contextlib.contextmanager
  def phase(self, name):
    torch = self.torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    try:
      yield
    finally:
      e1.record()
      self.records.setdefault(name, []).append((e0, e1))

  def summary(self, reduce="mean"):
    self.torch.cuda.synchronize(self.device)
    out = {}
    for name, evs in self.records.items():
      ms = [a.elapsed_time(b) for a, b in evs]
      out[name] = sum(ms) / len(ms) if reduce == "mean" else sum(ms)
    return out

  def max_over_ranks(self, group=None):
    """Per-phase mean, reduced with MAX over the ranks of ``group`` (the reportable number)."""
    import torch.distributed as dist
    s = self.summary()
    if not (dist.is_available() and dist.is_initialized()):
      return s
    names = sorted(s)
    t = self.torch.tensor([s[n] for n in names], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return dict(zip(names, t.tolist()))


@contextlib.contextmanager
def nvtx_range(name):
  try:
    import torch
    torch.cuda.nvtx.range_push(name)
    pushed = True
  except Exception:
    pushed = False
  try:
    yield
  finally:
    if pushed:
      import torch
      torch.cuda.nvtx.range_pop()


def exposed_ms(wait_begin_event, wait_end_event):
  """Milliseconds the compute stream spent blocked between two of its own events that bracket
  a wait on the communication stream (0 when the collective was fully overlapped)."""
  wait_end_event.synchronize()
  return wait_begin_event.elapsed_time(wait_end_event)
