"""Whole example programs on the 2-executor CPU engine - the plumbing bar of BASELINE.json
config #1 ("MNIST InputMode.SPARK sync SGD local[2] on CPU") and of the other MNIST drivers
(reference examples/mnist/keras/{mnist_spark,mnist_tf_ds}.py, estimator/mnist_pipeline.py)."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, timeout=420, env=None):
  e = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="2")
  e.update(env or {})
  for attempt in (1, 2):   # whole programs spawn executors and bind ports: one retry on a failure
    p = subprocess.run([sys.executable] + args, capture_output=True, text=True, timeout=timeout,
                       cwd=ROOT, env=e)
    if p.returncode == 0:
      break
    print("attempt {} of {} failed:\n{}\n{}".format(attempt, args[0], p.stdout[-2000:], p.stderr[-2000:]))
  assert p.returncode == 0, (p.stdout[-3000:], p.stderr[-3000:])
  return p.stdout + p.stderr


@pytest.fixture(scope="module")
def mnist(tmp_path_factory):
  d = str(tmp_path_factory.mktemp("mnist"))
  _run(["examples/mnist/mnist_data_setup.py", "--output", d + "/data", "--train_size", "2048",
        "--test_size", "256", "--num_partitions", "4"])
  return d


def test_baseline_config_1_mnist_inputmode_spark_sync_sgd_local2(mnist):
  out = _run(["examples/mnist/mnist_spark.py", "--cluster_size", "2", "--images_labels",
              mnist + "/data/csv/train", "--epochs", "2", "--batch_size", "32",
              "--learning_rate", "0.05", "--export_dir", mnist + "/export_spark",
              "--model_dir", mnist + "/model_spark"])
  assert "gloo, world 2" in out                       # two CPU workers in one process group
  assert os.path.exists(mnist + "/export_spark/weights.pt")
  from tensorflowonspark_b200.utils import checkpoint
  step, state = checkpoint.load(mnist + "/model_spark")   # the chief's periodic weights checkpoint
  assert step > 0 and len(state) > 1
  assert os.path.exists(mnist + "/model_spark/signature.json")
  from tensorflowonspark_b200.utils import summary
  (events,) = summary.event_files(mnist + "/model_spark")   # what tensorboard --logdir would read
  scalars = [e["scalars"] for e in summary.read_events(events) if e["scalars"]]
  assert scalars and "loss" in scalars[-1]
  losses = [float(x) for x in re.findall(r"loss ([\d.]+)", out)]
  assert not losses or losses[-1] < 2.4               # (fewer than 100 steps print nothing)
  inf = _run(["examples/mnist/mnist_inference.py", "--cluster_size", "2", "--images_labels",
              mnist + "/data/tfr/test", "--export_dir", mnist + "/export_spark", "--output",
              mnist + "/pred_spark"])
  assert float(re.search(r"accuracy: ([\d.]+)", inf).group(1)) > 0.3   # chance is 0.1


def test_mnist_tf_ds_streaming_tfrecord_pipeline(mnist):
  out = _run(["examples/mnist/mnist_tf_ds.py", "--cluster_size", "2", "--images_labels",
              mnist + "/data/tfr/train/part-*", "--epochs", "2", "--batch_size", "32",
              "--num_examples", "2048", "--buffer_size", "512", "--learning_rate", "0.05",
              "--model_dir", mnist + "/model_ds", "--export_dir", mnist + "/export_ds"])
  assert out.count("saved weights to") == 2           # one weights checkpoint per epoch
  assert os.path.exists(mnist + "/export_ds/signature.json")
  assert os.path.exists(mnist + "/model_ds/signature.json")   # checkpoints are servable


def test_estimator_pipeline_train_then_serve_newest_checkpoint(mnist):
  _run(["examples/mnist/estimator/mnist_pipeline.py", "--cluster_size", "2", "--images_labels",
        mnist + "/data/csv/train", "--epochs", "2", "--batch_size", "32", "--learning_rate",
        "0.05", "--model_dir", mnist + "/model_est", "--export_dir", mnist + "/export_est"])
  assert os.path.exists(mnist + "/export_est/weights.pt")
  # inference from the export, then with NO export: TFModel falls back to model_dir's newest
  # checkpoint (reference pipeline.py:549-555)
  for export in (mnist + "/export_est", mnist + "/no_such_export"):
    out = _run(["examples/mnist/estimator/mnist_pipeline.py", "--mode", "inference", "--format",
                "tfr", "--cluster_size", "2", "--images_labels", mnist + "/data/tfr/test",
                "--model_dir", mnist + "/model_est", "--export_dir", export, "--output",
                mnist + "/pred_est"])
    assert float(re.search(r"inference accuracy: ([\d.]+)", out).group(1)) > 0.5


def test_estimator_tf_mode_with_evaluator_sidecar_and_event_files(mnist):
  """reference examples/mnist/estimator/mnist_tf.py: chief + evaluator + worker; the evaluator
  follows the chief's checkpoints and both leave TensorBoard event files under model_dir."""
  from tensorflowonspark_b200.utils import summary
  md = mnist + "/model_est_tf"
  out = _run(["examples/mnist/estimator/mnist_tf.py", "--cluster_size", "3", "--images_labels",
              mnist + "/data/tfr", "--num_examples", "2048", "--epochs", "2", "--batch_size", "32",
              "--learning_rate", "0.05", "--model_dir", md, "--export_dir", mnist + "/export_est_tf",
              "--save_checkpoints_steps", "10", "--eval_interval", "0.3"])
  assert "evaluator:0 received the stop signal" in out
  evals = [line.split() for line in open(md + "/eval.log")]
  assert evals and float(evals[-1][2]) > 0.3            # accuracy of the last evaluated checkpoint
  (train_events,) = summary.event_files(md)
  assert any("loss" in e["scalars"] for e in summary.read_events(train_events))
  (eval_events,) = summary.event_files(md + "/eval")
  acc = [e["scalars"]["accuracy"] for e in summary.read_events(eval_events) if e["scalars"]]
  assert len(acc) == len(evals) and abs(acc[-1] - float(evals[-1][2])) < 1e-5


def test_restart_from_checkpoint_after_an_injected_rank_failure(mnist):
  """SURVEY 5.3 / 5.4 recovery story: a rank raises mid-training, TFCluster.shutdown leaves the
  driver with an error, utils.recovery.run_with_restarts brings the job up again on a fresh
  context and the nodes resume from the newest complete checkpoint (not from step 0)."""
  md = mnist + "/model_resilient"
  out = _run(["examples/mnist/mnist_resilient.py", "--cluster_size", "2", "--images_labels",
              mnist + "/data/tfr", "--model_dir", md, "--max_steps", "60", "--save_steps", "10",
              "--inject", "raise:rank=1:step=25"])
  assert "injected fault on rank 1 at step 25" in out and "attempt 1 failed" in out
  assert "chief:0 starts at step 0" in out and "chief:0 starts at step 20" in out
  assert "worker:0 starts at step 20" in out             # every rank restores the same file
  assert "job finished after 2 attempt(s)" in out
  from tensorflowonspark_b200.utils import checkpoint
  step, _ = checkpoint.load(md)
  assert step == 60


def test_streaming_feed_with_async_parameter_server_stops_on_terminate(mnist, tmp_path):
  """reference examples/mnist/estimator/mnist_spark_streaming.py: DStream micro-batches feed one
  worker, parameters live on a ps node and are updated without barriers; the worker's
  ``terminate()`` reaches the driver through the reservation server and ``shutdown(ssc)`` ends
  the stream."""
  import shutil
  import time
  watched = tmp_path / "stream"
  watched.mkdir()
  log = open(str(tmp_path / "out.log"), "w")
  env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="2")
  p = subprocess.Popen([sys.executable, "examples/mnist/mnist_spark_streaming.py", "--cluster_size", "2",
                        "--images_labels", str(watched), "--max_examples", "256", "--interval", "0.3"],
                       cwd=ROOT, env=env, stdout=log, stderr=subprocess.STDOUT)
  try:
    time.sleep(4)                                        # cluster up, stream (probably) started
    parts = sorted(f for f in os.listdir(mnist + "/data/csv/train") if f.startswith("part-"))
    for i in range(90):                                  # new files keep appearing until the job ends:
      if p.poll() is not None:                           # a slow start only delays the first batch
        break
      shutil.copy(os.path.join(mnist, "data/csv/train", parts[i % len(parts)]),
                  str(watched / "f{}.csv".format(i)))
      time.sleep(1)
    rc = p.wait(timeout=60)
  finally:
    if p.poll() is None:
      p.kill()
    log.close()
  out = open(str(tmp_path / "out.log")).read()
  assert rc == 0, out[-3000:]
  assert "terminate() invoked" in out and "server done, stopping the StreamingContext" in out
  assert "slot mode" not in out and "serving parameters" in out          # plain (Hogwild) ps node


def test_remaining_mnist_drivers_tf_mode_and_estimator_spark_mode(mnist):
  """examples/mnist/mnist_tf.py (InputMode.TENSORFLOW, every worker reads its shard) and
  estimator/mnist_spark.py (InputMode.SPARK, periodic checkpoints, terminate() on max steps)
  followed by estimator/mnist_inference.py (foreachPartition with a per-executor model cache) -
  reference examples/mnist/keras/mnist_tf.py, estimator/mnist_spark.py, estimator/mnist_inference.py."""
  _run(["examples/mnist/mnist_tf.py", "--cluster_size", "2", "--images_labels", mnist + "/data/tfr/train",
        "--num_examples", "2048", "--epochs", "2", "--batch_size", "32", "--learning_rate", "0.05",
        "--export_dir", mnist + "/export_tf"])
  assert os.path.exists(mnist + "/export_tf/signature.json")
  _run(["examples/mnist/estimator/mnist_spark.py", "--cluster_size", "2", "--images_labels",
        mnist + "/data/csv/train", "--num_examples", "2048", "--epochs", "2", "--batch_size", "32",
        "--learning_rate", "0.05", "--model_dir", mnist + "/model_es", "--export_dir", mnist + "/export_es"])
  from tensorflowonspark_b200.utils import checkpoint
  step, _ = checkpoint.load(mnist + "/model_es")
  assert step >= 50                                       # 90 % of 2048 * 2 / (32 * 2) collective steps
  for export in ("/export_tf", "/export_es"):
    out = _run(["examples/mnist/estimator/mnist_inference.py", "--cluster_size", "2", "--images_labels",
                mnist + "/data/tfr/test", "--export_dir", mnist + export, "--output",
                mnist + "/pred" + export.replace("/", "_")])
    assert float(re.search(r"accuracy: ([\d.]+)", out).group(1)) > 0.3


def test_keras_style_ml_pipeline_and_the_cli_helpers(mnist):
  """examples/mnist/mnist_pipeline.py (TFEstimator.fit -> TFModel.transform + argmax, reference
  examples/mnist/keras/mnist_pipeline.py:120-146) and examples/utils/mnist_reshape.py."""
  import json
  out = _run(["examples/mnist/mnist_pipeline.py", "--cluster_size", "2", "--images_labels",
              mnist + "/data/csv/train", "--epochs", "1", "--batch_size", "32", "--learning_rate", "0.05",
              "--export_dir", mnist + "/export_pipe"])
  assert float(re.search(r"pipeline accuracy on \d+ rows: ([\d.]+)", out).group(1)) > 0.3
  row = ",".join(["7"] + [str(i % 256) for i in range(784)])
  shaped = json.loads(_run(["examples/utils/mnist_reshape.py", row]).strip().splitlines()[-1])
  assert shaped["label"] == 7 and len(shaped["image"]) == 28 and shaped["image"][1][0] == 28
