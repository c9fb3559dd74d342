"""GPU tier, 2+ GPUs: fused all-reduce + optimizer parity with torch.distributed, broadcast and
device-side barrier visibility (tools/gpu_check_multi.py under torchrun)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def _free_port():
  """A port nobody listens on right now (fixed rendezvous ports collide with leftovers)."""
  import socket
  with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as sk:
    sk.bind(("127.0.0.1", 0))
    return str(sk.getsockname()[1])


def test_fused_collectives_two_ranks():
  import torch
  if torch.cuda.device_count() < 2:
    pytest.skip("needs 2 GPUs")
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  env = dict(os.environ, TFOS_FULL="0")
  p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node",
                      "2", "--master-addr", "127.0.0.1", "--master-port", _free_port(),
                      os.path.join(root, "tools", "gpu_check_multi.py")],
                     capture_output=True, text=True, timeout=600, env=env, cwd=root)
  print(p.stdout[-4000:], p.stderr[-2000:])
  assert p.returncode == 0 and "ALL OK" in p.stdout


@pytest.mark.parametrize("model", ["resnet", "unet"])
def test_lockstep_and_replica_consistency_under_cuda_graphs(model):
  """A slow rank must slow every rank down (the fused all-reduce is a real barrier inside the
  captured graph) and the replicas' bf16 weights must stay bit-identical."""
  import re
  import torch
  if torch.cuda.device_count() < 2:
    pytest.skip("needs 2 GPUs")
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node",
                      "2", "--master-addr", "127.0.0.1", "--master-port", _free_port(),
                      os.path.join(root, "tools", "lockstep_check.py"), model],
                     capture_output=True, text=True, timeout=600, cwd=root)
  print(p.stdout[-3000:], p.stderr[-2000:])
  assert p.returncode == 0
  rows = re.findall(r"rank (\d) \w+: ([\d.]+) ms/step .*= ([\d.e+-]+)", p.stdout)
  assert len(rows) == 2
  for _, ms, diff in rows:
    assert float(ms) >= 49.0 and float(diff) == 0.0


def test_spark_bootstrap_lockstep():
  """Same property with the communicator bootstrapped over the reservation board by Spark nodes."""
  import torch
  if torch.cuda.device_count() < 2:
    pytest.skip("needs 2 GPUs")
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  p = subprocess.run([sys.executable, os.path.join(root, "tools", "lockstep_spark_check.py"),
                      "--cluster_size", "2", "--input_mode", "spark"],
                     capture_output=True, text=True, timeout=600, cwd=root)
  print(p.stdout[-3000:], p.stderr[-2000:])
  assert p.returncode == 0 and "LOCKSTEP OK" in p.stdout


def test_two_level_allreduce_halves_and_inter_host_step():
  """Several-hosts path on the GPUs of one box (tools/gpu_check_hier.py): the fused kernel's
  PHASE 1 / PHASE 2 halves ("1 host x 2 GPUs") and NCCL between one-GPU "hosts" - against
  torch.distributed.all_reduce + the update in PyTorch.  (Run the tool with 4 ranks by hand for
  the "2 hosts x 2 GPUs" case; it is not part of the tier because it has not been run yet.)"""
  import torch
  if torch.cuda.device_count() < 2:
    pytest.skip("needs 2 GPUs")
  ranks = 2
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node",
                      str(ranks), "--master-addr", "127.0.0.1", "--master-port", _free_port(),
                      os.path.join(root, "tools", "gpu_check_hier.py")],
                     capture_output=True, text=True, timeout=600, cwd=root)
  print(p.stdout[-4000:], p.stderr[-2000:])
  assert p.returncode == 0 and "HIER CHECK PASSED" in p.stdout
