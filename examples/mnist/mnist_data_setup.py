"""Prepare MNIST-shaped data as CSV and TFRecords (reference: examples/mnist/mnist_data_setup.py:27-65,
which downloads MNIST through tensorflow_datasets).  There is no network here, so the data is either
read from local IDX files (``--idx_dir`` holding train-images-idx3-ubyte etc.) or synthesised: ten
fixed random class templates plus noise, which a CNN learns to >95 % in one epoch - enough to
exercise every code path with a meaningful accuracy signal.

  python examples/mnist/mnist_data_setup.py --output /tmp/mnist --num_partitions 10
"""
import argparse
import os
import struct
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def synthetic_mnist(n, seed=0, noise=0.25):
  rng = np.random.RandomState(1234)
  templates = (rng.rand(10, 28, 28) > 0.75).astype(np.float32)
  rng = np.random.RandomState(seed)
  labels = rng.randint(0, 10, size=n)
  images = templates[labels] * 0.8 + rng.rand(n, 28, 28).astype(np.float32) * noise
  return (np.clip(images, 0, 1) * 255).astype(np.uint8), labels.astype(np.int64)


def read_idx(idx_dir, split):
  prefix = "train" if split == "train" else "t10k"
  with open(os.path.join(idx_dir, prefix + "-images-idx3-ubyte"), "rb") as f:
    _, n, h, w = struct.unpack(">IIII", f.read(16))
    images = np.frombuffer(f.read(), dtype=np.uint8).reshape(n, h, w)
  with open(os.path.join(idx_dir, prefix + "-labels-idx1-ubyte"), "rb") as f:
    f.read(8)
    labels = np.frombuffer(f.read(), dtype=np.uint8).astype(np.int64)
  return images, labels


if __name__ == "__main__":
  from tensorflowonspark_b200 import dfutil
  from tensorflowonspark_b200._spark import SparkConf, SparkContext, SparkSession

  parser = argparse.ArgumentParser()
  parser.add_argument("--num_partitions", type=int, default=10)
  parser.add_argument("--output", default="data/mnist")
  parser.add_argument("--idx_dir", default=None, help="directory with the MNIST IDX files")
  parser.add_argument("--train_size", type=int, default=60000)
  parser.add_argument("--test_size", type=int, default=10000)
  args = parser.parse_args()

  sc = SparkContext(conf=SparkConf().setAppName("mnist_data_setup"))
  spark = SparkSession(sc)
  for split, n, seed in (("train", args.train_size, 0), ("test", args.test_size, 1)):
    if args.idx_dir:
      images, labels = read_idx(args.idx_dir, split)
    else:
      images, labels = synthetic_mnist(n, seed)
    rows = [(int(l), im.reshape(-1).tolist()) for im, l in zip(images, labels)]
    rdd = sc.parallelize(rows, args.num_partitions)
    # CSV: label,pix0,...,pix783 (what mnist_spark.py feeds through InputMode.SPARK)
    rdd.map(lambda r: ",".join(str(v) for v in [r[0]] + r[1])).saveAsTextFile(
        os.path.join(args.output, "csv", split))
    # TFRecords of Examples {image: int64[784], label: int64}
    df = spark.createDataFrame(rdd.map(lambda r: (r[1], r[0])), ["image", "label"])
    dfutil.saveAsTFRecords(df, os.path.join(args.output, "tfr", split))
    print("wrote {} {} examples under {}".format(len(rows), split, args.output))
  sc.stop()
