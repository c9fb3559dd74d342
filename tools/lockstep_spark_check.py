"""Same property as tools/lockstep_check.py, but with the communicator bootstrapped the way Spark
nodes do it (ctx.symmetric_comm() over the reservation board): a slow rank must slow every rank
down, and the replicas must stay bit-identical.

  python tools/lockstep_spark_check.py --cluster_size 2 [--input_mode spark]
"""
import argparse
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main_fun(args, ctx):
  import torch
  from tensorflowonspark_b200.models import unet
  torch.cuda.set_device(0)
  comm = ctx.symmetric_comm() if ctx.world_size > 1 else None
  net = unet.UNetTrainer(batch=8, image=128, classes=3, device="cuda:0", lr=1e-3, comm=comm)
  if comm is not None:
    comm.broadcast("weights", root=0)
    comm.broadcast("aux32", root=0)
  x, _ = net.synthetic_batch(seed=ctx.rank)
  net.set_input(x, (x[..., 0] > 127).int() + (x[..., 1] > 200).int())
  net.train_step()
  net.capture()
  torch.cuda.synchronize()
  t0 = time.time()
  for _ in range(args["steps"]):
    if ctx.rank == 1:
      time.sleep(0.05)
    net.train_step()
  torch.cuda.synchronize()
  ms = (time.time() - t0) / args["steps"] * 1e3
  digest = float(net.store.weights.float().double().sum())
  with open("{}{}".format(args["out"], ctx.rank), "w") as f:
    json.dump({"rank": ctx.rank, "world": ctx.world_size, "ms_per_step": ms, "digest": digest,
               "comm": None if comm is None else [comm.rank, comm.world]}, f)
  if args["spark"]:
    ctx.get_data_feed().terminate()


if __name__ == "__main__":
  p = argparse.ArgumentParser()
  p.add_argument("--cluster_size", type=int, default=2)
  p.add_argument("--steps", type=int, default=20)
  p.add_argument("--input_mode", default="tf")
  a = p.parse_args()
  from tensorflowonspark_b200 import TFCluster
  from tensorflowonspark_b200._spark import SparkConf, SparkContext
  conf = SparkConf().setAppName("lockstep").set("spark.executor.instances", str(a.cluster_size)) \
      .set("spark.executor.resource.gpu.amount", "1").set("spark.task.resource.gpu.amount", "1")
  sc = SparkContext(conf=conf)
  out = tempfile.mkdtemp() + "/r"
  spark = a.input_mode == "spark"
  cluster = TFCluster.run(sc, main_fun, {"out": out, "steps": a.steps, "spark": spark},
                          a.cluster_size, 0,
                          input_mode=TFCluster.InputMode.SPARK if spark else TFCluster.InputMode.TENSORFLOW,
                          master_node="chief")
  if spark:
    cluster.train(sc.parallelize(range(100), 4), 1)
  cluster.shutdown(grace_secs=10 if spark else 0)
  res = [json.load(open(out + str(r))) for r in range(a.cluster_size)]
  sc.stop()
  for r in res:
    print(r)
  ok = all(r["ms_per_step"] > 45 for r in res) and len({r["digest"] for r in res}) == 1
  print("LOCKSTEP", "OK" if ok else "BROKEN")
  sys.exit(0 if ok else 1)
