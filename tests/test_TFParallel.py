"""Independent parallel instances (scenarios of reference tests/test_TFParallel.py:16-53)."""
import pytest

from tensorflowonspark_b200 import TFParallel


def _fn(args, ctx):
  import os
  assert ctx.num_workers == args["n"]
  return [(ctx.worker_num, ctx.executor_id, os.environ.get("CUDA_VISIBLE_DEVICES"))]


def test_non_barrier(sc):
  out = TFParallel.run(sc, _fn, {"n": 2, "num_gpus": 0}, 2, use_barrier=False)
  assert sorted(o[0] for o in out) == [0, 1] and all(o[2] == "" for o in out)


def test_barrier(sc):
  out = TFParallel.run(sc, _fn, {"n": 2, "num_gpus": 0}, 2, use_barrier=True)
  assert sorted(o[0] for o in out) == [0, 1]


def test_barrier_needs_all_slots(sc):
  with pytest.raises(Exception, match="slots"):
    TFParallel.run(sc, _fn, {"n": 3, "num_gpus": 0}, 3, use_barrier=True)
