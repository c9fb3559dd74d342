"""Restart-from-checkpoint recovery, end to end: MNIST training (InputMode.TENSORFLOW, chief +
workers in one process group, checkpoint every ``--save_steps``) wrapped in
``utils.recovery.run_with_restarts``.  ``--inject`` (a ``TFOS_FAULT_INJECT`` spec such as
``raise:rank=1:step=25``) is active in the FIRST attempt only: a rank fails mid-training, the
driver sees the failed node, the job is started again on a fresh context and resumes from the
newest checkpoint instead of step 0.

  python examples/mnist/mnist_resilient.py --cluster_size 2 --images_labels /tmp/mnist/tfr \
      --model_dir /tmp/mnist_resilient --max_steps 60 --inject raise:rank=1:step=25

The reference stops at detect-and-abort (TFCluster.py:179-183) and leaves the re-submission to
the operator; resuming is TensorFlow's implicit ``model_dir`` behaviour
(examples/mnist/estimator/mnist_spark.py:94-97).
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
  from tensorflowonspark_b200.utils import checkpoint
  root = TFNode.local_path(ctx.absolute_path(args.images_labels))
  model_dir = TFNode.local_path(ctx.absolute_path(args.model_dir))
  trainer = mnist_common.Trainer(ctx, args.batch_size, args.learning_rate)
  start = 0
  latest = checkpoint.latest_checkpoint(model_dir)
  if latest:   # every rank restores the same complete file: identical replicas, no broadcast
    start, state = checkpoint.load(latest)
    trainer.load_state_dict(state)
  print("{}:{} starts at step {} ({})".format(ctx.job_name, ctx.task_index, start, trainer.desc))
  files = sorted(glob.glob(os.path.join(root, "train", "part-*")))
  cols = tfrecord.decode_batch([r for f in files[ctx.rank::ctx.world_size] for r in tfrecord.read_records(f)],
                               {"image": ("int64", 784, np.uint8), "label": ("int64", 1)})
  images, labels = cols["image"], cols["label"].reshape(-1)
  rng = np.random.RandomState(1000 * start + ctx.rank)
  for step in range(start + 1, args.max_steps + 1):
    idx = rng.randint(0, len(images), args.batch_size)
    loss = trainer.step(images[idx], labels[idx])       # (fault.maybe_inject fires in here)
    if ctx.is_chief and (step % args.save_steps == 0 or step == args.max_steps):
      checkpoint.save(model_dir, step, trainer.state_dict())
      print("saved step {} loss {:.4f}".format(step, loss))


if __name__ == "__main__":
  from tensorflowonspark_b200 import TFCluster
  from tensorflowonspark_b200._spark import SparkConf, SparkContext
  from tensorflowonspark_b200.utils import recovery

  parser = argparse.ArgumentParser()
  parser.add_argument("--batch_size", type=int, default=32)
  parser.add_argument("--cluster_size", type=int, default=2)
  parser.add_argument("--images_labels", required=True, help="dir with train/ TFRecords")
  parser.add_argument("--learning_rate", type=float, default=0.05)
  parser.add_argument("--model_dir", default="mnist_resilient")
  parser.add_argument("--max_steps", type=int, default=60)
  parser.add_argument("--save_steps", type=int, default=10)
  parser.add_argument("--max_restarts", type=int, default=2)
  parser.add_argument("--inject", default="", help="TFOS_FAULT_INJECT spec for the first attempt only")
  args = parser.parse_args()

  def make_context():
    # executors are created with the context and inherit the environment of this moment
    return SparkContext(conf=SparkConf().setAppName("mnist_resilient").set(
        "spark.executor.instances", str(args.cluster_size)))

  def job(sc, attempt):
    cluster = TFCluster.run(sc, main_fun, args, args.cluster_size, num_ps=0,
                            input_mode=TFCluster.InputMode.TENSORFLOW, master_node="chief")
    cluster.shutdown(grace_secs=1)

  if args.inject:
    os.environ["TFOS_FAULT_INJECT"] = args.inject
  attempts = recovery.run_with_restarts(
      make_context, job, max_restarts=args.max_restarts,
      on_failure=lambda attempt, exc: os.environ.pop("TFOS_FAULT_INJECT", None))
  print("job finished after {} attempt(s)".format(attempts))
