"""MNIST with InputMode.TENSORFLOW reading TFRecords straight from disk through a streaming input
pipeline - list_files -> repeat -> shuffle -> interleave -> parse -> batch - instead of loading
them up front (reference: examples/mnist/keras/mnist_tf_ds.py:41-50 builds this chain with
tf.data; :84-117 checkpoints the weights every epoch and exports from the chief).

Two record layouts, as in the reference (``--data_format``):
  tfos  records written by examples/mnist/mnist_data_setup.py (image = 784 ints, label = class
        index or a one-hot vector of 10)
  tfds  records as tensorflow_datasets writes them (image = PNG bytes, label = 1 int)

  python examples/mnist/mnist_tf_ds.py --cluster_size 2 --images_labels '/tmp/mnist/tfr/train/part-*'
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main_fun(args, ctx):
  import numpy as np
  import mnist_common
  from tensorflowonspark_b200 import TFNode, tfrecord
  from tensorflowonspark_b200.utils import checkpoint, data

  def parse_tfds(rec):
    ex = tfrecord.decode_example(rec)
    image = data.decode_png_gray(ex["image"][1][0]).reshape(-1)
    return image, np.int64(ex["label"][1][0])

  trainer = mnist_common.Trainer(ctx, args.batch_size, args.learning_rate)
  print("{}:{} using {}".format(ctx.job_name, ctx.task_index, trainer.desc))
  pattern = TFNode.local_path(ctx.absolute_path(args.images_labels))
  ds = data.TFRecordPipeline(pattern, epochs=args.epochs, shuffle_buffer=args.buffer_size,
                             seed=ctx.rank).shard(ctx.world_size, ctx.rank)
  if args.data_format == "tfds":      # PNG-encoded images: decoded record by record
    ds = ds.map(parse_tfds).batch(args.batch_size)
  else:                               # int64 pixel lists: whole batches decoded natively into arrays
    ds = ds.decode({"image": ("int64", 784, np.uint8), "label": ("int64", 1)}).batch(args.batch_size)
  model_dir = TFNode.local_path(ctx.absolute_path(args.model_dir))
  # every rank must take the same number of collective steps: the executor with the fewest records
  # would end first ("Out of Range" in the reference), so the loop is bounded up front
  steps_per_epoch = int(args.num_examples * 0.9) // (args.batch_size * ctx.world_size)
  timer = mnist_common.StepTimer(logdir=model_dir if ctx.is_chief else None)
  step, loss = 0, None
  for images, labels in ds:
    if step >= steps_per_epoch * args.epochs:
      break
    loss = trainer.step(images, labels.reshape(-1))
    step += 1
    timer.tick(step, loss, args.batch_size * ctx.world_size)
    if step % steps_per_epoch == 0 and ctx.is_chief:     # ModelCheckpoint(save_weights_only=True)
      path = checkpoint.save(model_dir, step, trainer.state_dict(), model=trainer.served_model())
      print("epoch {}: saved weights to {}".format(step // steps_per_epoch, path))
  timer.close(step, loss)
  assert step == steps_per_epoch * args.epochs, "input ran dry after {} steps".format(step)
  if args.export_dir:
    trainer.export(args.export_dir, ctx.is_chief)


if __name__ == "__main__":
  from tensorflowonspark_b200 import TFCluster
  from tensorflowonspark_b200._spark import SparkConf, SparkContext

  parser = argparse.ArgumentParser()
  parser.add_argument("--batch_size", help="number of records per batch (per worker)", type=int, default=64)
  parser.add_argument("--buffer_size", help="size of the shuffle buffer", type=int, default=10000)
  parser.add_argument("--cluster_size", type=int, default=2)
  parser.add_argument("--data_format", choices=["tfos", "tfds"], default="tfos")
  parser.add_argument("--epochs", type=int, default=3)
  parser.add_argument("--images_labels", required=True, help="glob / directory of TFRecord files")
  parser.add_argument("--num_examples", type=int, default=60000)
  parser.add_argument("--learning_rate", type=float, default=1e-3)
  parser.add_argument("--model_dir", default="mnist_model")
  parser.add_argument("--export_dir", default="mnist_export")
  parser.add_argument("--tensorboard", action="store_true")
  args = parser.parse_args()
  print("args:", args)
  sc = SparkContext(conf=SparkConf().setAppName("mnist_tf_ds").set("spark.executor.instances",
                                                                    str(args.cluster_size)))
  cluster = TFCluster.run(sc, main_fun, args, args.cluster_size, num_ps=0,
                          tensorboard=args.tensorboard, input_mode=TFCluster.InputMode.TENSORFLOW,
                          log_dir=args.model_dir, master_node="chief")
  cluster.shutdown(grace_secs=5)
  sc.stop()
