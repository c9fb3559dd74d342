"""The batch-inference application (python -m tensorflowonspark_b200.inference) - counterpart of
the reference's JVM ``Inference`` CLI (src/main/scala/com/yahoo/tensorflowonspark/Inference.scala:
30-79: load TFRecords with an optional schema hint -> TFModel.transform -> write JSON)."""
import glob
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_inference_cli_tfrecords_to_json(sc, spark, tmp_path):
  import torch
  from tensorflowonspark_b200 import dfutil
  from tensorflowonspark_b200.models import simple
  from tensorflowonspark_b200.utils import checkpoint
  # a model with known coefficients, exported the way compat.export_saved_model does
  model = simple.Linear(2, 1, input_name="x", output_name="y")
  with torch.no_grad():
    model.fc.weight.copy_(torch.tensor([[3.0, -2.0]]))
    model.fc.bias.copy_(torch.tensor([0.5]))
  export = str(tmp_path / "export")
  checkpoint.export_model(model, export, signatures={
      "serving_default": {"inputs": {"x": "x"}, "outputs": {"y": "y"}, "input_shapes": {"x": [-1, 2]}}})
  rng = np.random.RandomState(3)
  feats = rng.rand(37, 2)
  rows = [(int(i), [float(a), float(b)]) for i, (a, b) in enumerate(feats)]
  dfutil.saveAsTFRecords(spark.createDataFrame(rows, ["id", "features"]), str(tmp_path / "tfr"))
  out = str(tmp_path / "predictions")
  p = subprocess.run(
      [sys.executable, "-m", "tensorflowonspark_b200.inference", "--export_dir", export,
       "--input", str(tmp_path / "tfr"), "--schema_hint", "struct<id:bigint,features:array<float>>",
       "--input_mapping", json.dumps({"features": "x"}), "--output_mapping", json.dumps({"y": "prediction"}),
       "--output", out, "--batch_size", "8", "--cluster_size", "2", "--verbose"],
      cwd=ROOT, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""), capture_output=True, text=True, timeout=240)
  assert p.returncode == 0, (p.stdout[-2000:], p.stderr[-2000:])
  assert "features" in p.stdout                                   # --verbose printed the schema
  recs = [json.loads(line) for f in sorted(glob.glob(out + "/part-*")) for line in open(f)]
  assert len(recs) == 37 and set(recs[0]) == {"prediction"}       # only the mapped output columns
  got = sorted(r["prediction"][0] for r in recs)
  want = sorted((feats @ np.array([3.0, -2.0]) + 0.5).tolist())
  assert np.allclose(got, want, atol=1e-5)


def test_inference_cli_requires_its_options():
  from tensorflowonspark_b200 import inference
  import pytest
  with pytest.raises(SystemExit):
    inference.parse(["--export_dir", "m"])
  a = inference.parse(["--export_dir", "m", "--input", "i", "--input_mapping", "{}", "--output_mapping", "{}",
                       "--output", "o"])
  assert (a.batch_size, a.signature_def_key, a.tag_set, a.schema_hint) == (128, "serving_default", "serve", None)
