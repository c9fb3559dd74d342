"""Estimator-style MNIST as a Spark ML pipeline: ``TFEstimator.fit(df)`` trains with
InputMode.SPARK, checkpointing into ``--model_dir`` every ``save_checkpoints_steps`` steps and
stopping the feed at 90 % of the expected steps (StopFeedHook); ``--mode inference`` runs
``TFModel.transform`` from the export - or, when no export exists, from the newest checkpoint in
``--model_dir`` (reference: examples/mnist/estimator/mnist_pipeline.py:7-117 train loop,
:143-176 driver; tensorflowonspark/pipeline.py:549-555 for the model_dir fallback).

  python examples/mnist/estimator/mnist_pipeline.py --cluster_size 2 --images_labels /tmp/mnist/csv/train
  python examples/mnist/estimator/mnist_pipeline.py --mode inference --format tfr \
      --images_labels /tmp/mnist/tfr/test --output /tmp/mnist_predictions
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
  if latest:
    step, state = checkpoint.load(latest)
    est.load_state_dict(state)
  tf_feed = TFNode.DataFeed(ctx.mgr, input_mapping=args.input_mapping)
  # synchronous training: every worker must stop before any of them runs out of RDD partitions,
  # so train for 90 % of the expected number of steps
  steps = 60000 * args.epochs / args.batch_size if not getattr(args, "num_examples", 0) \
      else args.num_examples * args.epochs / args.batch_size
  max_steps_per_worker = step + int(steps / ctx.num_workers * 0.9)
  while not tf_feed.should_stop() and step < max_steps_per_worker:
    batch = tf_feed.next_batch(args.batch_size)
    if len(batch["image"]) < args.batch_size:
      continue
    est.step(np.asarray(batch["image"], dtype=np.uint8), np.asarray(batch["label"]))
    step += 1
    if ctx.is_chief and step % 100 == 0:          # RunConfig(save_checkpoints_steps=100)
      checkpoint.save(model_dir, step, est.state_dict(), model=est.served_model(),
                      signatures=est.sig)
  if ctx.is_chief:
    checkpoint.save(model_dir, step, est.state_dict(), model=est.served_model(), signatures=est.sig)
    if args.export_dir:
      print("Exporting saved_model to {}".format(args.export_dir))
  if args.export_dir:
    est.export(args.export_dir, ctx.is_chief)
  tf_feed.terminate()       # StopFeedHook.end()


if __name__ == "__main__":
  from tensorflowonspark_b200 import dfutil
  from tensorflowonspark_b200._spark import SparkConf, SparkContext, SparkSession
  from tensorflowonspark_b200.pipeline import TFEstimator, TFModel

  parser = argparse.ArgumentParser()
  parser.add_argument("--batch_size", type=int, default=64)
  parser.add_argument("--buffer_size", type=int, default=10000)
  parser.add_argument("--cluster_size", type=int, default=2)
  parser.add_argument("--epochs", type=int, default=3)
  parser.add_argument("--format", choices=["csv", "tfr"], default="csv")
  parser.add_argument("--images_labels", required=True)
  parser.add_argument("--learning_rate", type=float, default=1e-3)
  parser.add_argument("--mode", choices=["train", "inference"], default="train")
  parser.add_argument("--model_dir", default="mnist_model")
  parser.add_argument("--export_dir", default="mnist_export")
  parser.add_argument("--output", default="predictions")
  parser.add_argument("--tensorboard", action="store_true")
  args = parser.parse_args()
  print("args:", args)

  sc = SparkContext(conf=SparkConf().setAppName("mnist_estimator").set(
      "spark.executor.instances", str(args.cluster_size)))
  spark = SparkSession(sc)
  if args.format == "tfr":
    df = dfutil.loadTFRecords(sc, args.images_labels)
  else:
    def parse(ln):
      vec = [int(x) for x in ln.split(",")]
      return (vec[1:], vec[0])
    df = spark.createDataFrame(sc.textFile(args.images_labels).map(parse), ["image", "label"])
  df.show(5)

  if args.mode == "train":
    args.num_examples = df.count()
    estimator = TFEstimator(main_fun, args) \
        .setInputMapping({"image": "image", "label": "label"}) \
        .setModelDir(args.model_dir) \
        .setExportDir(args.export_dir) \
        .setClusterSize(args.cluster_size) \
        .setTensorboard(args.tensorboard) \
        .setEpochs(args.epochs) \
        .setBatchSize(args.batch_size) \
        .setMasterNode("chief") \
        .setGraceSecs(5)
    model = estimator.fit(df)
  else:
    model = TFModel(args) \
        .setInputMapping({"image": "image"}) \
        .setOutputMapping({"logits": "prediction"}) \
        .setSignatureDefKey("serving_default") \
        .setBatchSize(args.batch_size)
    if args.export_dir and os.path.exists(os.path.join(args.export_dir, "signature.json")):
      model.setExportDir(args.export_dir)
    else:                       # nothing exported: serve the newest checkpoint
      args.export_dir = None
      model = TFModel(args).setInputMapping({"image": "image"}) \
          .setOutputMapping({"logits": "prediction"}).setSignatureDefKey("serving_default") \
          .setBatchSize(args.batch_size).setModelDir(args.model_dir)
    preds = model.transform(df)
    rows = preds.collect()
    labels = [r.label for r in df.select("label").collect()]
    argmax = [max(range(len(r.prediction)), key=lambda i: r.prediction[i]) for r in rows]
    acc = sum(int(a == b) for a, b in zip(argmax, labels)) / float(len(labels))
    print("inference accuracy: {:.4f} over {} rows".format(acc, len(labels)))
    out = args.output[len("file://"):] if args.output.startswith("file://") else args.output
    os.makedirs(out, exist_ok=True)
    import json
    with open(os.path.join(out, "part-00000.json"), "w") as f:
      for r, a in zip(rows, argmax):
        f.write(json.dumps({"prediction": list(r.prediction), "argmax": a}) + "\n")
  sc.stop()
