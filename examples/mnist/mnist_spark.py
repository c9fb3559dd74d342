"""MNIST with InputMode.SPARK: Spark feeds CSV rows to the nodes through TFNode.DataFeed.

Driver-script shape of the reference (examples/mnist/keras/mnist_spark.py:74-107): build the
cluster with ``TFCluster.run(..., input_mode=InputMode.SPARK, master_node='chief')``, feed with
``cluster.train(rdd, epochs)``, then ``cluster.shutdown()``.  Only the body of ``main_fun``
changed: a PyTorch/native-kernel training loop instead of Keras under MultiWorkerMirroredStrategy.

  python examples/mnist/mnist_spark.py --cluster_size 2 --images_labels /tmp/mnist/csv/train \
      --model_dir /tmp/mnist_model --export_dir /tmp/mnist_export
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main_fun(args, ctx):
  import mnist_common
  from tensorflowonspark_b200 import TFNode
  from tensorflowonspark_b200.utils import checkpoint
  trainer = mnist_common.Trainer(ctx, args.batch_size, args.learning_rate)
  print("{}:{} using {}".format(ctx.job_name, ctx.task_index, trainer.desc))
  model_dir = TFNode.local_path(ctx.absolute_path(args.model_dir)) if args.model_dir else None
  feed = ctx.get_data_feed(train_mode=True)
  # every step is a collective: stop at 90 % of the expected steps so that a worker whose
  # partitions were shorter never leaves its peers waiting (same guard as the reference, :62-66)
  steps = int(args.num_examples * args.epochs * 0.9 / (args.batch_size * ctx.num_workers))
  timer = mnist_common.StepTimer(logdir=model_dir if ctx.is_chief else None)
  done, loss = 0, None
  for step in range(steps):
    # columnar fast path: one [B, 785] array sliced out of the shared-memory ring - the rows are
    # never expanded into python lists (the reference pulls them one by one, next_batch(1))
    cols = feed.next_batch_arrays(args.batch_size)
    if not cols or len(cols[0]) < args.batch_size:
      break
    loss = trainer.step(cols[0][:, 1:], cols[0][:, 0])
    done = step + 1
    timer.tick(step, loss, args.batch_size * ctx.num_workers)
    # the chief's weights-NNNN checkpoints (ModelCheckpoint callback of the reference, :57-60);
    # each one carries the serving signature, so TFModel can score straight from model_dir
    if ctx.is_chief and model_dir and done % args.save_steps == 0:
      checkpoint.save(model_dir, done, trainer.state_dict(), model=trainer.served_model())
  if ctx.is_chief and model_dir and done % args.save_steps:
    checkpoint.save(model_dir, done, trainer.state_dict(), model=trainer.served_model())
  timer.close(done, loss)
  if args.export_dir:
    trainer.export(args.export_dir, ctx.is_chief)
  feed.terminate()


if __name__ == "__main__":
  from tensorflowonspark_b200 import TFCluster
  from tensorflowonspark_b200._spark import SparkConf, SparkContext

  parser = argparse.ArgumentParser()
  parser.add_argument("--batch_size", type=int, default=64)
  parser.add_argument("--cluster_size", type=int, default=None)
  parser.add_argument("--epochs", type=int, default=3)
  parser.add_argument("--images_labels", required=True, help="path to the MNIST CSV directory")
  parser.add_argument("--learning_rate", type=float, default=1e-3)
  parser.add_argument("--model_dir", default="mnist_model")
  parser.add_argument("--export_dir", default="mnist_export")
  parser.add_argument("--save_steps", type=int, default=100)
  parser.add_argument("--tensorboard", action="store_true")
  args = parser.parse_args()

  conf = SparkConf().setAppName("mnist_spark")
  if args.cluster_size:
    conf.set("spark.executor.instances", str(args.cluster_size))
  sc = SparkContext(conf=conf)
  executors = int(sc.getConf().get("spark.executor.instances", "1"))
  args.cluster_size = args.cluster_size or executors

  def parse(line):
    # one C-level parse per line (8x faster than int() per pixel); rows travel as int arrays and
    # are stacked straight into the shared-memory feed ring
    import numpy as np
    return np.fromstring(line, dtype=np.int64, sep=",")

  images_labels = sc.textFile(args.images_labels).map(parse)
  args.num_examples = images_labels.count()
  print("args:", args)

  cluster = TFCluster.run(sc, main_fun, args, args.cluster_size, num_ps=0,
                          tensorboard=args.tensorboard, input_mode=TFCluster.InputMode.SPARK,
                          log_dir=args.model_dir, master_node="chief")
  cluster.train(images_labels, args.epochs)
  cluster.shutdown(grace_secs=5)
  sc.stop()
