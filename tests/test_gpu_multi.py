"""GPU tier, 2+ GPUs: fused all-reduce + optimizer parity with torch.distributed, broadcast and
device-side barrier visibility (tools/gpu_check_multi.py under torchrun)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_fused_collectives_two_ranks():
  import torch
  if torch.cuda.device_count() < 2:
    pytest.skip("needs 2 GPUs")
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  env = dict(os.environ, TFOS_FULL="0")
  p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node",
                      "2", "--master-addr", "127.0.0.1", "--master-port", "29533",
                      os.path.join(root, "tools", "gpu_check_multi.py")],
                     capture_output=True, text=True, timeout=600, env=env, cwd=root)
  print(p.stdout[-4000:], p.stderr[-2000:])
  assert p.returncode == 0 and "ALL OK" in p.stdout
