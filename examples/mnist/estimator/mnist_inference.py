"""Raw ``foreachPartition`` inference over TFRecord files - no cluster object at all: every Spark
task loads the exported model once per executor process (cached) and scores its files
(reference: examples/mnist/estimator/mnist_inference.py:24-89).

  python examples/mnist/estimator/mnist_inference.py --images_labels /tmp/mnist/tfr/test \
      --export_dir /tmp/mnist_export --output /tmp/mnist_predictions
"""
import argparse
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(HERE))))

_cache = {}


def _model(export_dir):
  if export_dir not in _cache:
    from tensorflowonspark_b200.utils import checkpoint
    _cache[export_dir] = checkpoint.load_model(export_dir, "serve")[0]
  return _cache[export_dir]


def run_partition(args):
  def _run(it):
    import numpy as np
    from tensorflowonspark_b200 import tfrecord
    model = _model(args.export_dir)
    os.makedirs(args.output, exist_ok=True)
    for path in it:
      recs = [tfrecord.decode_example(r) for r in tfrecord.read_records(path)]
      with open(os.path.join(args.output, os.path.basename(path)), "w") as out:
        for i in range(0, len(recs), args.batch_size):
          chunk = recs[i:i + args.batch_size]
          x = np.asarray([e["image"][1] for e in chunk], dtype=np.float32) / 255.0
          pred = model(image=x)["prediction"].cpu().numpy()
          for e, p in zip(chunk, pred):
            out.write("{} {}\n".format(e["label"][1][0], int(p)))
  return _run


if __name__ == "__main__":
  import glob
  from tensorflowonspark_b200._spark import SparkConf, SparkContext

  parser = argparse.ArgumentParser()
  parser.add_argument("--batch_size", type=int, default=128)
  parser.add_argument("--cluster_size", type=int, default=2)
  parser.add_argument("--images_labels", required=True)
  parser.add_argument("--export_dir", required=True)
  parser.add_argument("--output", default="predictions")
  args = parser.parse_args()
  sc = SparkContext(conf=SparkConf().setAppName("mnist_estimator_inference").set(
      "spark.executor.instances", str(args.cluster_size)))
  files = sorted(glob.glob(os.path.join(args.images_labels, "part-*")))
  sc.parallelize(files, min(len(files), 4 * args.cluster_size)).foreachPartition(
      run_partition(args))
  lines = [l.split() for f in glob.glob(os.path.join(args.output, "part-*")) for l in open(f)]
  print("predictions: {}  accuracy: {:.4f}".format(
      len(lines), sum(a == b for a, b in lines) / max(1, len(lines))))
  sc.stop()
