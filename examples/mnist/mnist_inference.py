"""Parallel batch inference with TFParallel: N independent single-GPU instances, each taking
``files[worker_num::num_workers]`` and writing ``part-NNNNN`` (reference: examples/mnist/keras/
mnist_inference.py:24-79).

  python examples/mnist/mnist_inference.py --cluster_size 2 --images_labels /tmp/mnist/tfr/test \
      --export_dir /tmp/mnist_export --output /tmp/mnist_predictions
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def inference(args, ctx):
  import glob
  import numpy as np
  from tensorflowonspark_b200 import TFNode, tfrecord
  from tensorflowonspark_b200.utils import checkpoint
  model, sig = checkpoint.load_model(TFNode.local_path(args.export_dir), "serve")
  path = TFNode.local_path(ctx.absolute_path(args.images_labels))
  files = sorted(glob.glob(os.path.join(path, "part-*")))[ctx.worker_num::ctx.num_workers]
  out_dir = TFNode.local_path(ctx.absolute_path(args.output))
  os.makedirs(out_dir, exist_ok=True)
  n = correct = 0
  with open(os.path.join(out_dir, "part-{:05d}".format(ctx.worker_num)), "w") as out:
    for f in files:
      recs = [tfrecord.decode_example(r) for r in tfrecord.read_records(f)]
      for i in range(0, len(recs), args.batch_size):
        chunk = recs[i:i + args.batch_size]
        x = np.asarray([e["image"][1] for e in chunk], dtype=np.float32) / 255.0
        y = np.asarray([e["label"][1][0] for e in chunk])
        pred = model(image=x)["prediction"].cpu().numpy()
        for label, p in zip(y, pred):
          out.write("{} {}\n".format(label, p))
        n += len(chunk)
        correct += int((pred == y).sum())
  return [(ctx.worker_num, n, correct)]


if __name__ == "__main__":
  from tensorflowonspark_b200 import TFParallel
  from tensorflowonspark_b200._spark import SparkConf, SparkContext

  parser = argparse.ArgumentParser()
  parser.add_argument("--batch_size", type=int, default=128)
  parser.add_argument("--cluster_size", type=int, default=2)
  parser.add_argument("--images_labels", required=True)
  parser.add_argument("--export_dir", required=True)
  parser.add_argument("--output", default="predictions")
  args = parser.parse_args()
  sc = SparkContext(conf=SparkConf().setAppName("mnist_inference").set(
      "spark.executor.instances", str(args.cluster_size)))
  stats = TFParallel.run(sc, inference, args, args.cluster_size, use_barrier=True)
  total, ok = sum(s[1] for s in stats), sum(s[2] for s in stats)
  print("predictions: {}  accuracy: {:.4f}".format(total, ok / max(1, total)))
  sc.stop()
