"""Transport-only benchmark of the InputMode.SPARK data path (no GPU needed): RDD partitions ->
feeder tasks -> shared-memory ring -> DataFeed.next_batch_arrays on the node.  Rows are the
segmentation example's (uint8 image [128,128,3], uint8 mask [128,128]) = 64 KiB each, produced
without a random generator so that the producer is not the bottleneck.

  python tools/bench_feed.py --executors 2 --examples 98304
"""
import argparse
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

IMG = 128


def rows_of(n, seed):
  import numpy as np
  img = np.full((IMG, IMG, 3), seed % 251, np.uint8)
  mask = np.full((IMG, IMG), seed % 3, np.uint8)
  return [(img, mask) for _ in range(n)]


def main_fun(args, ctx):
  feed = ctx.get_data_feed(train_mode=True)
  B, seen, nbytes = args["batch"], 0, 0
  t0 = None
  while not feed.should_stop():
    cols = feed.next_batch_arrays(B)
    if not cols or len(cols[0]) == 0:
      continue
    if t0 is None:
      t0 = time.time()       # clock starts with the first batch: steady state, not start-up
      continue
    seen += len(cols[0])
    nbytes += cols[0].nbytes + cols[1].nbytes
  dt = time.time() - (t0 or time.time())
  with open(args["out"] + str(ctx.executor_id), "w") as f:
    json.dump({"rows": seen, "bytes": nbytes, "seconds": dt}, f)


if __name__ == "__main__":
  p = argparse.ArgumentParser()
  p.add_argument("--executors", type=int, default=2)
  p.add_argument("--examples", type=int, default=98304)   # long enough for the ring slots to be warm
  p.add_argument("--batch", type=int, default=64)
  a = p.parse_args()
  from tensorflowonspark_b200 import TFCluster
  from tensorflowonspark_b200._spark import SparkConf, SparkContext
  sc = SparkContext(conf=SparkConf().setAppName("bench_feed").set("spark.executor.instances",
                                                                  str(a.executors)))
  out = tempfile.mkdtemp() + "/r"
  parts = a.executors * 8
  per = a.examples // parts
  rdd = sc.parallelize(range(parts), parts).flatMap(lambda i: rows_of(per, i))
  cluster = TFCluster.run(sc, main_fun, {"batch": a.batch, "out": out}, a.executors, 0,
                          input_mode=TFCluster.InputMode.SPARK)
  t0 = time.time()
  cluster.train(rdd, 1)
  wall = time.time() - t0
  cluster.shutdown(grace_secs=1)
  res = [json.load(open(out + str(i))) for i in range(a.executors)]
  sc.stop()
  rows = sum(r["rows"] for r in res)
  mbs = sum(r["bytes"] / max(r["seconds"], 1e-9) for r in res) / 1e6
  print(json.dumps({"metric": "DataFeed transport (InputMode.SPARK, shm ring)", "executors": a.executors,
                    "rows": rows, "aggregate_MB_per_s": round(mbs, 1),
                    "rows_per_s": round(sum(r["rows"] / max(r["seconds"], 1e-9) for r in res)),
                    "train_wall_s": round(wall, 2)}))
