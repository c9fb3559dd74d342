"""Checkpoint / resume and the inference artefact format.

The reference delegates all of this to TensorFlow in user code and only contributes path
helpers and the ``grace_secs`` window (Keras ``ModelCheckpoint`` in
examples/mnist/keras/mnist_spark.py:51-53, Estimator ``save_checkpoints_steps`` in
examples/mnist/estimator/mnist_spark.py:94-97, SavedModel export via compat.py:10-17).
Here:

* ``save(model_dir, step, state)`` / ``latest_checkpoint`` / ``load`` - atomic (write to a
  temp name, fsync, rename) ``ckpt-<step>.pt`` files plus a ``checkpoint`` index, so a restarted
  job resumes from the newest complete file;
* ``export_model`` / ``load_model`` - ``export_dir/{weights.pt, signature.json}``; the JSON keeps
  the meaning of the pipeline params ``signature_def_key`` / ``tag_set`` / input & output
  mappings (tensorflowonspark/pipeline.py:244-296).

Every path may name a remote filesystem (``hdfs://``, ``s3://`` ...; utils/fs.py), as
``ctx.absolute_path(model_dir)`` produces on a cluster whose ``defaultFS`` is not local.
"""
import json
import logging
import re

from . import fs

logger = logging.getLogger(__name__)

INDEX = "checkpoint"


def _local(path):
  return fs.local(path) if fs.is_local(path) else str(path)


def _atomic_torch_save(obj, path):
  import torch
  fs.write_atomic(path, lambda f: torch.save(obj, f))


def _torch_load(path, map_location):
  import torch
  with fs.open_read(path) as f:
    return torch.load(f, map_location=map_location, weights_only=False)


def save(model_dir, step, state, keep=5, model=None, signatures=None):
  """Write ``state`` (any picklable / tensor dict) as checkpoint ``step``; prune old ones.

  ``model_dir`` may live on any filesystem ``utils/fs.py`` resolves (plain / ``file://`` paths,
  ``hdfs://``, ``s3://`` ...).  ``model`` (optional): the object the state came from.  Its
  ``export_builder`` is recorded in ``model_dir/signature.json`` so that the newest checkpoint can
  be *served* without an export (``load_model_dir``; the reference's TFModel falls back to
  ``tf.train.latest_checkpoint(model_dir)`` when no ``export_dir`` is given, pipeline.py:549-555)."""
  model_dir = _local(model_dir)
  path = fs.join(model_dir, "ckpt-{:08d}.pt".format(int(step)))
  _atomic_torch_save({"step": int(step), "state": state}, path)
  if model is not None and getattr(model, "export_builder", None):
    _write_signature(model_dir, model, "serve", signatures)
  fs.write_text(fs.join(model_dir, INDEX), json.dumps({"latest": fs.basename(path), "step": int(step)}))
  ckpts = sorted(f for f in fs.listdir(model_dir) if re.match(r"ckpt-\d+\.pt$", f))
  for old in ckpts[:-keep] if keep else []:
    try:
      fs.remove(fs.join(model_dir, old))
    except (OSError, IOError):
      pass
  return path


def latest_checkpoint(model_dir):
  """Path of the newest complete checkpoint in ``model_dir`` or None."""
  model_dir = _local(model_dir)
  if not fs.isdir(model_dir):
    return None
  idx = fs.join(model_dir, INDEX)
  if fs.exists(idx):
    try:
      p = fs.join(model_dir, json.loads(fs.read_text(idx))["latest"])
      if fs.exists(p):
        return p
    except Exception:
      pass
  ckpts = sorted(f for f in fs.listdir(model_dir) if re.match(r"ckpt-\d+\.pt$", f))
  return fs.join(model_dir, ckpts[-1]) if ckpts else None


def load(path_or_dir, map_location="cpu"):
  """(step, state) of a checkpoint file, or of the latest one in a directory; (0, None) if none."""
  p = _local(path_or_dir)
  if fs.isdir(p):
    p = latest_checkpoint(p)
  if not p or not fs.exists(p):
    return 0, None
  blob = _torch_load(p, map_location)
  return blob["step"], blob["state"]


def _state_of(model):
  if hasattr(model, "state_dict"):
    return model.state_dict()
  if isinstance(model, dict):
    return model
  raise TypeError("cannot export object of type {}".format(type(model)))


def _write_signature(directory, model, tag_set, signatures, builder=None):
  sig = {
      "tag_set": tag_set if isinstance(tag_set, (list, tuple)) else str(tag_set).split(","),
      "signatures": signatures or getattr(model, "export_signatures", None)
      or {"serving_default": {"inputs": {}, "outputs": {}}},
      "builder": builder or getattr(model, "export_builder", None),
      "builder_args": getattr(model, "export_builder_args", {}),
  }
  fs.write_text(fs.join(directory, "signature.json"), json.dumps(sig, indent=1))
  return sig


def export_model(model, export_dir, tag_set="serve", signatures=None, builder=None):
  """Write the inference artefact.

  Args:
    model: object with ``state_dict()`` (torch module / native-engine model) or a state dict.
    export_dir: target directory (``file://`` prefix allowed).
    tag_set: tag(s) stored with the artefact (string or list).
    signatures: ``{key: {'inputs': {alias: name}, 'outputs': {alias: name}}}``.
    builder: dotted ``module:function`` that rebuilds the model for loading
      (``fn(state_dict, **builder_args) -> callable``); defaults to the model's
      ``export_builder`` attribute when present.
  """
  export_dir = _local(export_dir)
  fs.makedirs(export_dir)
  _atomic_torch_save(_state_of(model), fs.join(export_dir, "weights.pt"))
  _write_signature(export_dir, model, tag_set, signatures, builder)
  logger.info("exported model to %s", export_dir)
  return export_dir


def _build(sig, state):
  import importlib
  if sig.get("builder"):
    mod, fn = sig["builder"].split(":")
    return getattr(importlib.import_module(mod), fn)(state, **sig.get("builder_args", {}))
  return state


def load_model_dir(model_dir, map_location="cpu"):
  """(callable, signature_json) from the NEWEST CHECKPOINT of a training ``model_dir`` - the
  inference fallback when nothing was exported (reference pipeline.py:549-555).  Needs the
  ``signature.json`` that ``save(..., model=...)`` writes next to the checkpoints."""
  model_dir = _local(model_dir)
  sig_path = fs.join(model_dir, "signature.json")
  latest = latest_checkpoint(model_dir)
  if latest is None:
    raise IOError("no checkpoint found in model_dir {}".format(model_dir))
  if not fs.exists(sig_path):
    raise IOError("{} has checkpoints but no signature.json: save them with "
                  "checkpoint.save(..., model=<model>) to make them servable".format(model_dir))
  sig = json.loads(fs.read_text(sig_path))
  _, state = load(latest, map_location)
  if isinstance(state, dict) and "model" in state and "optimizer" in state:
    state = state["model"]      # a training checkpoint: only the parameters are served
  logger.info("serving checkpoint %s", latest)
  return _build(sig, state), sig


def load_model(export_dir, tag_set=None, map_location="cpu"):
  """(callable_or_state_dict, signature_json) from an exported artefact."""
  export_dir = _local(export_dir)
  sig = json.loads(fs.read_text(fs.join(export_dir, "signature.json")))
  if tag_set:
    want = tag_set if isinstance(tag_set, (list, tuple)) else str(tag_set).split(",")
    if not set(want) <= set(sig["tag_set"]):
      raise ValueError("export at {} has tags {}, requested {}".format(export_dir, sig["tag_set"],
                                                                       want))
  state = _torch_load(fs.join(export_dir, "weights.pt"), map_location)
  return _build(sig, state), sig
