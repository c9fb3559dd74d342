"""Per-step metrics as JSON lines, one file per rank, plus a driver-side reducer.

The reference has no metrics registry (logging only).  Fields are free-form;
the conventional ones are ``step, loss, images_per_s, step_ms, exposed_allreduce_ms,
feed_queue_depth, h2d_mb_s``.  ``reduce_max`` merges the per-rank files the way multi-GPU
timings must be reported: the slowest rank defines the step.  ``Exporter`` serves the same fields
live on a Prometheus ``/metrics`` endpoint."""
import glob
import json
import os
import time


class StepLogger(object):

  def __init__(self, path, rank=0, exporter=None):
    """``exporter``: an :class:`Exporter` that mirrors every numeric field as a live gauge."""
    self.exporter = exporter
    self.path = "{}.rank{}".format(path, rank) if rank is not None else path
    d = os.path.dirname(self.path)
    if d:
      os.makedirs(d, exist_ok=True)
    self.f = open(self.path, "a")
    self.rank = rank

  def log(self, **fields):
    fields.setdefault("ts", time.time())
    fields.setdefault("rank", self.rank)
    self.f.write(json.dumps(fields) + "\n")
    self.f.flush()
    if self.exporter is not None:
      self.exporter.update(**fields)

  def close(self):
    self.f.close()


def read(path):
  with open(path) as f:
    return [json.loads(line) for line in f if line.strip()]


def reduce_max(path_prefix, key="step_ms", by="step"):
  """{step: max over ranks of ``key``} from the files ``<prefix>.rank*``."""
  out = {}
  for p in glob.glob(path_prefix + ".rank*"):
    for rec in read(p):
      if key in rec and by in rec:
        out[rec[by]] = max(out.get(rec[by], float("-inf")), rec[key])
  return out


# ------------------------------------------------------------------ live scrape endpoint
class Exporter(object):
  """``/metrics`` endpoint of one node for a Prometheus server (or ``curl``): the numeric fields
  a training loop logs become gauges ``tfos_<field>{rank="r"}``, ``step`` a counter-like gauge,
  plus ``tfos_last_log_timestamp_seconds`` - a scrape that stops moving is a stalled node, which
  is the question the reference can only answer from executor logs.

    exp = metrics.Exporter(rank=ctx.rank)            # port 0: any free port -> exp.port
    log = metrics.StepLogger(path, rank=ctx.rank, exporter=exp)
    log.log(step=s, loss=l, images_per_s=r)          # JSONL line + gauges in one call

  Uses ``prometheus_client`` with a private registry (several exporters may live in one test
  process); ``available()`` tells whether the package is installed."""

  def __init__(self, rank=0, port=0, addr="0.0.0.0", prefix="tfos"):
    import prometheus_client as prom
    self._prom, self.rank, self.prefix = prom, rank, prefix
    self.registry = prom.CollectorRegistry()
    self._gauges = {}
    self._stamp = prom.Gauge(prefix + "_last_log_timestamp_seconds", "wall clock of the last update",
                             ["rank"], registry=self.registry)
    self._server, self._thread = prom.start_http_server(port, addr=addr, registry=self.registry)
    self.port = self._server.server_address[1]

  @staticmethod
  def available():
    try:
      import prometheus_client  # noqa: F401
      return True
    except ImportError:
      return False

  def update(self, **fields):
    for k, v in fields.items():
      if isinstance(v, bool) or not isinstance(v, (int, float)) or k in ("rank", "ts"):
        continue
      g = self._gauges.get(k)
      if g is None:
        name = "{}_{}".format(self.prefix, "".join(c if c.isalnum() else "_" for c in k))
        g = self._gauges[k] = self._prom.Gauge(name, "training-loop field '{}'".format(k), ["rank"],
                                               registry=self.registry)
      g.labels(rank=str(self.rank)).set(float(v))
    self._stamp.labels(rank=str(self.rank)).set(time.time())

  def url(self, host="127.0.0.1"):
    return "http://{}:{}/metrics".format(host, self.port)

  def close(self):
    self._server.shutdown()
    self._server.server_close()
