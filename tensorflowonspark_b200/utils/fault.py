"""Fault injection for resilience tests (the reference's only injected faults
are two "FAKE exception" tests, tests/test_TFCluster.py:50-91).

``TFOS_FAULT_INJECT`` holds ';'-separated specs ``action:key=value:...``:

  raise:rank=1:step=5            raise an exception in that rank at that step
  kill:rank=0:step=3             hard-exit the process (os._exit(137)), as if OOM-killed
  delay:rank=1:step=4:secs=2.5   stall a rank (exposes missing timeouts / barrier hangs)
  drop_feed:rank=0:step=2        stop consuming the DataFeed without terminate()

Training loops call :func:`maybe_inject(rank, step)` once per step; it is a no-op when the
variable is unset.  Whatever happens must surface through the normal channels: the node's error
queue -> feeder task / shutdown() -> driver exception (never a hang).
"""
import logging
import os
import time

logger = logging.getLogger(__name__)
ENV = "TFOS_FAULT_INJECT"


class InjectedFault(Exception):
  pass


def parse(spec):
  faults = []
  for part in (spec or "").split(";"):
    part = part.strip()
    if not part:
      continue
    fields = part.split(":")
    f = {"action": fields[0]}
    for kv in fields[1:]:
      k, v = kv.split("=", 1)
      f[k] = float(v) if k == "secs" else int(v)
    faults.append(f)
  return faults


def maybe_inject(rank, step, spec=None):
  """Returns 'drop_feed' when the caller should stop consuming, else None."""
  spec = os.environ.get(ENV) if spec is None else spec
  if not spec:
    return None
  for f in parse(spec):
    if f.get("rank", rank) != rank or f.get("step", step) != step:
      continue
    logger.warning("fault injection: %s on rank %d at step %d", f["action"], rank, step)
    if f["action"] == "raise":
      raise InjectedFault("injected fault on rank {} at step {}".format(rank, step))
    if f["action"] == "kill":
      os._exit(137)
    if f["action"] == "delay":
      time.sleep(f.get("secs", 1.0))
    if f["action"] == "drop_feed":
      return "drop_feed"
  return None
