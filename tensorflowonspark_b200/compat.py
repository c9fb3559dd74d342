"""Helpers that keep user code independent of the compute backend.

The reference's versions shim TensorFlow releases (tensorflowonspark/compat.py:10-31); the
same three entry points here shim *this* runtime, so ported ``map_fun``s keep their shape.
"""
import logging
import os

logger = logging.getLogger(__name__)


def export_saved_model(model, export_dir, is_chief=False, signatures=None):
  """Export ``model`` for inference.  Every rank must call this (state gathering may be
  collective); only the chief's output lands in ``export_dir``, the others write to a scratch
  ``worker_model`` directory exactly like the reference (compat.py:15-17)."""
  from .utils import checkpoint
  target = export_dir if is_chief else os.path.join(os.path.dirname(export_dir.rstrip("/")) or ".",
                                                    "worker_model")
  return checkpoint.export_model(model, target, signatures=signatures)


def disable_auto_shard(options=None):
  """No-op: data sharding is explicit here (ctx.rank / ctx.world_size), nothing to disable."""
  return options


def is_gpu_available():
  """True when this process can use a CUDA device."""
  try:
    import torch
    return torch.cuda.is_available()
  except Exception:
    return False
