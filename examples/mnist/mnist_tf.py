"""MNIST with InputMode.TENSORFLOW: Spark only launches the nodes; every worker reads its own
shard of the TFRecord files (reference: examples/mnist/keras/mnist_tf_ds.py:41-50,84-117 and
mnist_tf.py:24-27 - tf.data with AutoShardPolicy.DATA there, explicit ``files[rank::world]`` here).

  python examples/mnist/mnist_tf.py --cluster_size 2 --images_labels /tmp/mnist/tfr/train
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main_fun(args, ctx):
  import glob
  import numpy as np
  import mnist_common
  from tensorflowonspark_b200 import TFNode, tfrecord
  step_fn, export_fn, desc = mnist_common.make_trainer(ctx, args.batch_size, args.learning_rate)
  print("{}:{} using {}".format(ctx.job_name, ctx.task_index, desc))
  path = TFNode.local_path(ctx.absolute_path(args.images_labels))
  files = sorted(glob.glob(os.path.join(path, "part-*")))
  mine = files[ctx.rank::ctx.world_size]
  images, labels = [], []
  for f in mine:
    for rec in tfrecord.read_records(f):
      ex = tfrecord.decode_example(rec)
      images.append(ex["image"][1])
      labels.append(ex["label"][1][0])
  images, labels = np.asarray(images, dtype=np.uint8), np.asarray(labels)
  # all ranks must run the same number of collective steps
  per_rank = args.num_examples // ctx.world_size
  steps_per_epoch = int(per_rank * 0.9) // args.batch_size
  timer = mnist_common.StepTimer()
  rng = np.random.RandomState(ctx.rank)
  step = 0
  for epoch in range(args.epochs):
    order = rng.permutation(len(images))
    for i in range(steps_per_epoch):
      idx = order[(i * args.batch_size) % max(1, len(order) - args.batch_size):][:args.batch_size]
      loss = step_fn(images[idx], labels[idx])
      timer.tick(step, loss, args.batch_size * ctx.world_size)
      step += 1
  if args.export_dir:
    export_fn(args.export_dir, ctx.is_chief)


if __name__ == "__main__":
  from tensorflowonspark_b200 import TFCluster
  from tensorflowonspark_b200._spark import SparkConf, SparkContext

  parser = argparse.ArgumentParser()
  parser.add_argument("--batch_size", type=int, default=64)
  parser.add_argument("--cluster_size", type=int, default=2)
  parser.add_argument("--epochs", type=int, default=3)
  parser.add_argument("--images_labels", required=True, help="path to the MNIST TFRecords")
  parser.add_argument("--num_examples", type=int, default=60000)
  parser.add_argument("--learning_rate", type=float, default=1e-3)
  parser.add_argument("--export_dir", default="mnist_export")
  parser.add_argument("--tensorboard", action="store_true")
  args = parser.parse_args()
  sc = SparkContext(conf=SparkConf().setAppName("mnist_tf").set("spark.executor.instances",
                                                                 str(args.cluster_size)))
  cluster = TFCluster.run(sc, main_fun, args, args.cluster_size, num_ps=0,
                          tensorboard=args.tensorboard, input_mode=TFCluster.InputMode.TENSORFLOW,
                          master_node="chief")
  cluster.shutdown(grace_secs=5)
  sc.stop()
