"""Checkpoint / export / metrics helpers and host utilities."""
import os

import pytest
import torch

from tensorflowonspark_b200 import gpu_info, util
from tensorflowonspark_b200.utils import checkpoint, metrics


def test_checkpoint_save_latest_resume_prune(tmp_path):
  d = str(tmp_path / "ckpt")
  assert checkpoint.latest_checkpoint(d) is None and checkpoint.load(d) == (0, None)
  for step in range(1, 9):
    checkpoint.save(d, step * 10, {"w": torch.full((3,), float(step))}, keep=3)
  files = sorted(f for f in os.listdir(d) if f.startswith("ckpt-"))
  assert files == ["ckpt-00000060.pt", "ckpt-00000070.pt", "ckpt-00000080.pt"]
  step, state = checkpoint.load(d)
  assert step == 80 and float(state["w"][0]) == 8.0
  # a torn write (temp file left behind) must not be picked up
  open(os.path.join(d, ".tmp-garbage.pt"), "wb").write(b"xx")
  assert checkpoint.latest_checkpoint(d).endswith("ckpt-00000080.pt")


def test_export_and_load_with_builder(tmp_path):
  from tensorflowonspark_b200.models import simple
  m = simple.Linear(2, 1)
  d = checkpoint.export_model(m, "file://" + str(tmp_path / "exp"), tag_set="serve",
                              signatures={"serving_default": {"inputs": {"x": "x"}}})
  served, sig = checkpoint.load_model(d, "serve")
  out = served(x=[[1.0, 2.0]])["y"]
  assert torch.allclose(out.cpu(), m(torch.tensor([[1.0, 2.0]])))
  assert sig["signatures"]["serving_default"]["inputs"] == {"x": "x"}
  with pytest.raises(ValueError):
    checkpoint.load_model(d, "train")


def test_metrics_reduce_max(tmp_path):
  prefix = str(tmp_path / "m.jsonl")
  for rank, ms in ((0, 10.0), (1, 12.5)):
    log = metrics.StepLogger(prefix, rank)
    log.log(step=1, step_ms=ms)
    log.log(step=2, step_ms=ms + 1)
    log.close()
  assert metrics.reduce_max(prefix) == {1: 12.5, 2: 13.5}


def test_executor_id_file(tmp_path, monkeypatch):
  monkeypatch.chdir(tmp_path)
  with pytest.raises(Exception, match="No executor_id file"):
    util.read_executor_id()
  util.write_executor_id(7)
  assert util.read_executor_id() == 7
  assert util.find_in_path("/nonexistent:" + str(tmp_path), "executor_id") == str(tmp_path / "executor_id")
  assert util.find_in_path("/nonexistent", "executor_id") is False
  assert util.get_ip_address().count(".") == 3


def test_gpu_placement_rules(monkeypatch):
  inv = ([(i, "GPU-%d" % i) for i in range(8)], {"GPU-1"})
  monkeypatch.setattr(gpu_info, "_inventory", lambda: inv)
  assert gpu_info.get_gpus(1, 0) == "0"
  assert gpu_info.get_gpus(2, 1) == "3,4"          # contiguous slice of the *free* list
  assert gpu_info.get_gpus(2, 3, format=gpu_info.AS_LIST) == ["7", "0"]  # wraps around
  assert len(set(gpu_info.get_gpus(3, -1).split(","))) == 3
  monkeypatch.setattr(gpu_info, "MAX_RETRIES", 0)
  with pytest.raises(Exception, match="Unable to find 8 free"):
    gpu_info.get_gpus(8, 0)
