"""Spark ML pipeline: TFEstimator.fit on a DataFrame, TFModel.transform for inference
(reference: examples/mnist/keras/mnist_pipeline.py:120-146).

  python examples/mnist/mnist_pipeline.py --cluster_size 2 --images_labels /tmp/mnist/csv/train \
      --export_dir /tmp/mnist_pipe_export
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def train_fn(args, ctx):
  import mnist_common
  step_fn, export_fn, desc = mnist_common.make_trainer(ctx, args.batch_size, args.learning_rate)
  feed = ctx.get_data_feed(input_mapping=args.input_mapping)
  steps = int(args.num_examples * args.epochs * 0.9 / (args.batch_size * ctx.num_workers))
  for step in range(steps):
    batch = feed.next_batch(args.batch_size)
    if len(batch["image"]) < args.batch_size:
      break
    step_fn(batch["image"], batch["label"])
  if args.export_dir:
    export_fn(args.export_dir, ctx.is_chief)
  feed.terminate()


if __name__ == "__main__":
  from tensorflowonspark_b200.pipeline import TFEstimator
  from tensorflowonspark_b200._spark import SparkConf, SparkContext, SparkSession

  parser = argparse.ArgumentParser()
  parser.add_argument("--batch_size", type=int, default=64)
  parser.add_argument("--cluster_size", type=int, default=2)
  parser.add_argument("--epochs", type=int, default=1)
  parser.add_argument("--images_labels", required=True)
  parser.add_argument("--learning_rate", type=float, default=1e-3)
  parser.add_argument("--export_dir", default="mnist_pipe_export")
  args = parser.parse_args()
  sc = SparkContext(conf=SparkConf().setAppName("mnist_pipeline").set(
      "spark.executor.instances", str(args.cluster_size)))
  spark = SparkSession(sc)

  def parse(line):
    v = [int(x) for x in line.split(",")]
    return (v[1:], v[0])

  df = spark.createDataFrame(sc.textFile(args.images_labels).map(parse), ["image", "label"])
  args.num_examples = df.count()
  estimator = TFEstimator(train_fn, args) \
      .setInputMapping({"image": "image", "label": "label"}) \
      .setExportDir(args.export_dir).setClusterSize(args.cluster_size) \
      .setMasterNode("chief").setEpochs(args.epochs).setBatchSize(args.batch_size).setGraceSecs(5)
  model = estimator.fit(df)
  model.setTagSet("serve").setSignatureDefKey("serving_default") \
       .setInputMapping({"image": "image"}).setOutputMapping({"prediction": "col_out"})
  preds = model.transform(df.limit(200))
  labels = [r.label for r in df.limit(200).collect()]
  got = [r.col_out for r in preds.collect()]
  acc = sum(int(a == b) for a, b in zip(labels, got)) / float(len(got))
  print("pipeline accuracy on 200 rows: {:.3f}".format(acc))
  sc.stop()
