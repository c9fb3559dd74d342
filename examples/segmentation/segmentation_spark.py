"""Image segmentation U-Net on Spark executors.

Reference: examples/segmentation/segmentation_spark.py:25-193 (MobileNetV2-encoder U-Net,
128x128x3 inputs, 3 classes, batch 64, Adam, InputMode.TENSORFLOW with tfds).  Both feeding modes
are provided here: ``--input_mode tf`` (each worker synthesises / reads its own data) and
``--input_mode spark`` (BASELINE.json config #4: the images travel RDD -> shared-memory pinned
ring -> ``cudaMemcpyAsync`` on a copy stream -> device, through ``TFNode.DataFeed``).

  python examples/segmentation/segmentation_spark.py --cluster_size 8 --batch_size 64 --steps 100
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

IMG = 128


def synth_rows(n, seed):
  """Synthetic oxford-pets-like rows: (uint8 image [128,128,3], uint8 mask [128,128]) as numpy
  arrays - the feeder packs them column-wise into the shared-memory ring without touching
  individual pixels in Python."""
  import numpy as np
  rng = np.random.RandomState(seed)
  pool = min(n, 64)   # 64 distinct random images per partition, cycled: the generator must not
  imgs = rng.randint(0, 256, size=(pool, IMG, IMG, 3), dtype=np.uint8)   # be the bottleneck
  masks = (imgs[..., 0] > 127).astype(np.uint8) + (imgs[..., 1] > 200).astype(np.uint8)
  return [(imgs[i % pool], masks[i % pool]) for i in range(n)]


def main_fun(args, ctx):
  import time
  import numpy as np
  import torch
  from tensorflowonspark_b200.feed import DevicePrefetcher
  from tensorflowonspark_b200.models import unet
  from tensorflowonspark_b200.utils import fault
  torch.cuda.set_device(0)
  comm = ctx.gradient_comm() if ctx.world_size > 1 else None   # one host: P2P/NVLS kernels; several: NCCL
  B = args.batch_size
  net = unet.UNetTrainer(batch=B, image=IMG, classes=3, device="cuda:0", lr=args.learning_rate,
                         comm=comm)
  if comm is not None:
    comm.broadcast("weights", root=0)
    comm.broadcast("aux32", root=0)
  t0, seen = time.time(), 0
  if args.input_mode == "spark":
    feed = ctx.get_data_feed(train_mode=True)
    feed.pin_ring()   # ring slots become DMA sources: no staging copy between feeder and GPU
    # masks stay uint8 on the wire (4x fewer bytes than int32); set_input widens them on the GPU
    pre = DevicePrefetcher([((B, IMG, IMG, 3), torch.uint8), ((B, IMG, IMG), torch.uint8)], "cuda:0")
    steps = int(args.num_examples * args.epochs * 0.9 / (B * ctx.num_workers))
    host = [0.0, 0.0, 0.0, 0.0]   # seconds of host time: feed, staging, input hand-over, launch
    for step in range(steps):
      h0 = time.time()
      cols = feed.next_batch_arrays(B)   # [images uint8 [B,128,128,3], masks uint8 [B,128,128]]
      if not cols or len(cols[0]) < B:
        break
      if pre._used[pre._w]:
        pre.ready[pre._w].synchronize()   # (timed with "feed": back-pressure wait, see push_arrays)
      h1 = time.time()
      pre.push_arrays(cols)   # ring views -> pinned staging -> async H2D (host runs <= 2 ahead)
      h2 = time.time()
      bx, by = pre.pop()
      net.set_input(bx, by)
      pre.release()
      h3 = time.time()
      fault.maybe_inject(ctx.rank, step)   # TFOS_FAULT_INJECT (resilience tests), no-op otherwise
      loss = net.train_step()
      h4 = time.time()
      for k, d in enumerate((h1 - h0, h2 - h1, h3 - h2, h4 - h3)):
        host[k] += d
      if step == 0:
        net.capture()          # static buffers: later steps replay one CUDA graph
        torch.cuda.synchronize()
        t0, seen = time.time(), -B   # steady-state rate: the clock starts after the capture
      seen += B
      if (step + 1) % 10 == 0 and ctx.is_chief:
        torch.cuda.synchronize()
        print("step {:4d} loss {:.4f} {:.0f} images/s".format(
            step + 1, float(loss), seen * ctx.num_workers / (time.time() - t0)))
    print("rank {} ran {} steps ({} rows); host ms/step: feed {:.2f} staging {:.2f} hand-over {:.2f} "
          "launch {:.2f}; H2D tensors straight from the ring {} / staged {}".format(
              ctx.rank, step + 1, seen + B, *([1e3 * h / (step + 1) for h in host]
                                              + [pre.direct_copies, pre.staged_copies])),
          flush=True)
    feed.terminate()
  else:
    x, y = net.synthetic_batch(seed=ctx.rank)
    net.set_input(x, (x[..., 0] > 127).int() + (x[..., 1] > 200).int())
    net.train_step()
    net.capture()
    torch.cuda.synchronize()
    t0 = time.time()          # steady-state rate: the clock starts after the capture
    for step in range(args.steps):
      fault.maybe_inject(ctx.rank, step)
      loss = net.train_step()
      if (step + 1) % 10 == 0 and ctx.is_chief:
        torch.cuda.synchronize()
        print("step {:4d} loss {:.4f} {:.0f} images/s".format(
            step + 1, float(loss), (step + 1) * B * ctx.world_size / (time.time() - t0)))
  torch.cuda.synchronize()


if __name__ == "__main__":
  from tensorflowonspark_b200 import TFCluster
  from tensorflowonspark_b200._spark import SparkConf, SparkContext

  parser = argparse.ArgumentParser()
  parser.add_argument("--batch_size", type=int, default=64)
  parser.add_argument("--cluster_size", type=int, default=1)
  parser.add_argument("--epochs", type=int, default=1)
  parser.add_argument("--steps", type=int, default=50)
  parser.add_argument("--learning_rate", type=float, default=1e-3)
  parser.add_argument("--input_mode", default="tf", choices=["tf", "spark"])
  parser.add_argument("--num_examples", type=int, default=2048)
  args = parser.parse_args()
  conf = SparkConf().setAppName("segmentation_spark") \
      .set("spark.executor.instances", str(args.cluster_size)) \
      .set("spark.executor.resource.gpu.amount", "1").set("spark.task.resource.gpu.amount", "1")
  sc = SparkContext(conf=conf)
  if args.input_mode == "spark":
    parts = args.cluster_size * 4
    per = args.num_examples // parts
    rdd = sc.parallelize(range(parts), parts).flatMap(lambda i: synth_rows(per, i))
    cluster = TFCluster.run(sc, main_fun, args, args.cluster_size, num_ps=0,
                            input_mode=TFCluster.InputMode.SPARK, master_node="chief")
    cluster.train(rdd, args.epochs)
    cluster.shutdown(grace_secs=2)
  else:
    cluster = TFCluster.run(sc, main_fun, args, args.cluster_size, num_ps=0,
                            input_mode=TFCluster.InputMode.TENSORFLOW, master_node="chief")
    cluster.shutdown()
  sc.stop()
