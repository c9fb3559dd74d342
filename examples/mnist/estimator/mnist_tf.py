"""Estimator-style MNIST with InputMode.TENSORFLOW and an **evaluator side-car**
(``eval_node=True``): the chief + workers train from their own TFRecord shards and checkpoint;
the evaluator node is outside the collective, watches ``--model_dir`` and scores every new
checkpoint on the test set until the driver stops it (reference:
examples/mnist/estimator/mnist_tf.py:62-79,107-108).

  python examples/mnist/estimator/mnist_tf.py --cluster_size 3 \
      --images_labels /tmp/mnist/tfr --model_dir /tmp/mnist_model
"""
import argparse
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(HERE))))
sys.path.insert(0, os.path.dirname(HERE))


def _load(files):
  import numpy as np
  from tensorflowonspark_b200 import tfrecord
  images, labels = [], []
  for f in files:
    for rec in tfrecord.read_records(f):
      ex = tfrecord.decode_example(rec)
      images.append(ex["image"][1])
      labels.append(ex["label"][1][0])
  return np.asarray(images, dtype=np.uint8), np.asarray(labels)


def main_fun(args, ctx):
  import glob
  import time
  import numpy as np
  import mnist_common
  from tensorflowonspark_b200 import TFNode
  from tensorflowonspark_b200.utils import checkpoint
  root = TFNode.local_path(ctx.absolute_path(args.images_labels))
  model_dir = TFNode.local_path(ctx.absolute_path(args.model_dir))

  if ctx.job_name == "evaluator":
    est = mnist_common.Trainer(ctx, args.batch_size, args.learning_rate, train=False)
    images, labels = _load(sorted(glob.glob(os.path.join(root, "test", "part-*"))))
    from tensorflowonspark_b200.utils import summary
    events = summary.SummaryWriter(os.path.join(model_dir, "eval"), flush_secs=0)   # tf.estimator's eval/ dir
    seen = None
    while True:   # stopped by the driver through the control queue at shutdown
      latest = checkpoint.latest_checkpoint(model_dir)
      if latest and latest != seen:
        step, state = checkpoint.load(latest)
        est.load_state_dict(state)
        loss, acc = est.evaluate(images, labels)
        print("evaluator: step {} loss {:.4f} accuracy {:.4f}".format(step, loss, acc))
        with open(os.path.join(model_dir, "eval.log"), "a") as f:
          f.write("{} {:.6f} {:.6f}\n".format(step, loss, acc))
        events.add_scalars({"loss": loss, "accuracy": acc}, step)
        seen = latest
      time.sleep(args.eval_interval)

  est = mnist_common.Trainer(ctx, args.batch_size, args.learning_rate)
  files = sorted(glob.glob(os.path.join(root, "train", "part-*")))
  images, labels = _load(files[ctx.rank::ctx.world_size])
  # every rank must run the same number of collective steps: derive them from the global size
  steps = args.max_steps or int(args.num_examples * args.epochs * 0.9 / ctx.world_size
                                ) // args.batch_size
  rng = np.random.RandomState(ctx.rank)
  timer = mnist_common.StepTimer(logdir=model_dir if ctx.is_chief else None)
  step, loss = 0, None
  for step in range(1, steps + 1):
    idx = rng.randint(0, len(images), args.batch_size)
    loss = est.step(images[idx], labels[idx])
    timer.tick(step, loss, args.batch_size * ctx.world_size)
    if ctx.is_chief and (step % args.save_checkpoints_steps == 0 or step == steps):
      checkpoint.save(model_dir, step, est.state_dict())
  timer.close(step, loss)
  if args.export_dir:
    est.export(args.export_dir, ctx.is_chief)


if __name__ == "__main__":
  from tensorflowonspark_b200 import TFCluster
  from tensorflowonspark_b200._spark import SparkConf, SparkContext

  parser = argparse.ArgumentParser()
  parser.add_argument("--batch_size", type=int, default=64)
  parser.add_argument("--cluster_size", type=int, default=3, help="chief + evaluator + workers")
  parser.add_argument("--epochs", type=int, default=1)
  parser.add_argument("--images_labels", required=True, help="dir with train/ and test/ TFRecords")
  parser.add_argument("--max_steps", type=int, default=0)
  parser.add_argument("--num_examples", type=int, default=60000)
  parser.add_argument("--learning_rate", type=float, default=1e-3)
  parser.add_argument("--model_dir", default="mnist_model")
  parser.add_argument("--export_dir", default="mnist_export")
  parser.add_argument("--save_checkpoints_steps", type=int, default=100)
  parser.add_argument("--eval_interval", type=float, default=1.0)
  parser.add_argument("--tensorboard", action="store_true")
  args = parser.parse_args()
  sc = SparkContext(conf=SparkConf().setAppName("mnist_estimator_tf").set(
      "spark.executor.instances", str(args.cluster_size)))
  cluster = TFCluster.run(sc, main_fun, args, args.cluster_size, num_ps=0,
                          tensorboard=args.tensorboard, input_mode=TFCluster.InputMode.TENSORFLOW,
                          log_dir=args.model_dir, master_node="chief", eval_node=True)
  cluster.shutdown(grace_secs=int(3 * args.eval_interval) + 2)
  sc.stop()
