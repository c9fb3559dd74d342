"""GPU tier: every native kernel against a plain PyTorch fp32 reference of the same op, and an
end-to-end ResNet-50 overfit through the native engine.  The check bodies live in
tools/gpu_check.py so they can also be run one by one under a timeout on the GPU box."""
import importlib.util
import os

import pytest

pytestmark = pytest.mark.gpu

_spec = importlib.util.spec_from_file_location(
    "gpu_check", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools",
                              "gpu_check.py"))
gpu_check = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(gpu_check)


def test_extension_is_loaded_not_a_fallback():
  from tensorflowonspark_b200 import _build
  mod = _build.load(required=True)
  assert mod.__file__.endswith("_tfos_b200_C.so") and "tensorflowonspark_b200/_ext" in mod.__file__


@pytest.mark.parametrize("name", sorted(gpu_check.CHECKS))
def test_kernel(name):
  import torch
  torch.manual_seed(0)
  assert gpu_check.CHECKS[name](), "numerics check '{}' failed (see captured stdout)".format(name)
