"""Per-step metrics as JSON lines, one file per rank, plus a driver-side reducer.

The reference has no metrics registry (logging only).  Fields are free-form;
the conventional ones are ``step, loss, images_per_s, step_ms, exposed_allreduce_ms,
feed_queue_depth, h2d_mb_s``.  ``reduce_max`` merges the per-rank files the way multi-GPU
timings must be reported: the slowest rank defines the step."""
import glob
import json
import os
import time


class StepLogger(object):

  def __init__(self, path, rank=0):
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
