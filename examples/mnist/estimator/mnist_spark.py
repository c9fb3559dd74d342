"""Estimator-style MNIST with InputMode.SPARK: a ``train_and_evaluate`` loop that checkpoints
every ``--save_checkpoints_steps`` steps, resumes from the newest checkpoint in ``--model_dir``,
stops the feed once ``--max_steps`` is reached (the reference's ``StopFeedHook`` ->
``tf_feed.terminate()``) and exports the model from the chief (reference:
examples/mnist/estimator/mnist_spark.py:14-22,94-118,153-155).

  python examples/mnist/estimator/mnist_spark.py --cluster_size 2 \
      --images_labels /tmp/mnist/csv/train --model_dir /tmp/mnist_model
"""
import argparse
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(HERE))))
sys.path.insert(0, os.path.dirname(HERE))


def main_fun(args, ctx):
  import numpy as np
  import mnist_common
  from tensorflowonspark_b200 import TFNode
  from tensorflowonspark_b200.utils import checkpoint
  est = mnist_common.Trainer(ctx, args.batch_size, args.learning_rate)
  model_dir = TFNode.local_path(ctx.absolute_path(args.model_dir))
  step = 0
  latest = checkpoint.latest_checkpoint(model_dir)
  if latest:  # every rank restores the same file: identical replicas, no broadcast needed
    step, state = checkpoint.load(latest)
    est.load_state_dict(state)
    print("{}:{} resumed from {} (step {})".format(ctx.job_name, ctx.task_index, latest, step))
  feed = ctx.get_data_feed(train_mode=True)
  timer = mnist_common.StepTimer(logdir=model_dir if ctx.is_chief else None)
  loss = None
  # each rank must take the same number of collective steps: stop at 90 % of the expected feed
  max_steps = args.max_steps or int(args.num_examples * args.epochs * 0.9
                                     / ctx.world_size / args.batch_size)
  while not feed.should_stop() and step < max_steps:
    cols = feed.next_batch_arrays(args.batch_size)   # [B, 785] view of the feed ring
    if not cols or len(cols[0]) < args.batch_size:
      continue
    arr = np.asarray(cols[0])
    loss = est.step(arr[:, 1:], arr[:, 0])
    step += 1
    timer.tick(step, loss, args.batch_size * ctx.world_size)
    if ctx.is_chief and step % args.save_checkpoints_steps == 0:
      checkpoint.save(model_dir, step, est.state_dict())
  timer.close(step, loss)
  if ctx.is_chief:
    checkpoint.save(model_dir, step, est.state_dict())
  if args.export_dir:   # before terminate(): the driver only grants grace_secs after the feed ends
    est.export(args.export_dir, ctx.is_chief)
  feed.terminate()   # StopFeedHook: drain what Spark still wants to push and end the feed job


if __name__ == "__main__":
  from tensorflowonspark_b200 import TFCluster
  from tensorflowonspark_b200._spark import SparkConf, SparkContext

  parser = argparse.ArgumentParser()
  parser.add_argument("--batch_size", type=int, default=64)
  parser.add_argument("--cluster_size", type=int, default=2)
  parser.add_argument("--epochs", type=int, default=1)
  parser.add_argument("--images_labels", required=True, help="CSV rows: label,pixel0..pixel783")
  parser.add_argument("--num_examples", type=int, default=60000)
  parser.add_argument("--max_steps", type=int, default=0)
  parser.add_argument("--learning_rate", type=float, default=1e-3)
  parser.add_argument("--model_dir", default="mnist_model")
  parser.add_argument("--export_dir", default="mnist_export")
  parser.add_argument("--save_checkpoints_steps", type=int, default=100)
  parser.add_argument("--tensorboard", action="store_true")
  args = parser.parse_args()
  sc = SparkContext(conf=SparkConf().setAppName("mnist_estimator").set(
      "spark.executor.instances", str(args.cluster_size)))
  def parse(line):
    import numpy as np
    return np.fromstring(line, dtype=np.int64, sep=",")   # one C-level parse per CSV line

  rows = sc.textFile(args.images_labels).map(parse)
  cluster = TFCluster.run(sc, main_fun, args, args.cluster_size, num_ps=0,
                          tensorboard=args.tensorboard, input_mode=TFCluster.InputMode.SPARK,
                          log_dir=args.model_dir, master_node="chief")
  cluster.train(rows, args.epochs)
  cluster.shutdown(grace_secs=5)
  sc.stop()
