"""TFParallel batch inference: one independent ResNet-50 replica per GPU, bf16, synthetic input
(BASELINE.json config "TFParallel batch inference ResNet-50 on 8xB200").

  python bench/inference_resnet50.py --gpus 2 --batch 256 --steps 30
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run_replica(args, ctx):
  import torch
  from tensorflowonspark_b200.models import resnet
  torch.cuda.set_device(0)
  net = resnet.ResNetTrainer(depth=50, batch=args.batch, image=224, device="cuda:0", training=False)
  x, _ = net.synthetic_batch(seed=ctx.worker_num)
  hx = x.cpu().pin_memory()
  for _ in range(5):
    net.forward_only(x)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(args.steps):
    net.forward_only(hx)            # includes the H2D copy of every batch
    top1 = net.logits.argmax(1)
  e1.record()
  torch.cuda.synchronize()
  ms = e0.elapsed_time(e1)
  return [(ctx.worker_num, ms, int(top1[0]))]


if __name__ == "__main__":
  from tensorflowonspark_b200 import TFParallel
  from tensorflowonspark_b200._spark import SparkConf, SparkContext
  p = argparse.ArgumentParser()
  p.add_argument("--gpus", type=int, default=1)
  p.add_argument("--batch", type=int, default=256)
  p.add_argument("--steps", type=int, default=30)
  args = p.parse_args()
  args.num_gpus = 1
  sc = SparkContext(conf=SparkConf().setAppName("inference_bench").set(
      "spark.executor.instances", str(args.gpus)))
  out = TFParallel.run(sc, run_replica, args, args.gpus, use_barrier=True)
  sc.stop()
  ms = max(o[1] for o in out)
  print(json.dumps({"metric": "ResNet-50 inference images/s (TFParallel, bf16, incl. H2D)",
                    "value": args.batch * args.steps * args.gpus / (ms / 1e3), "unit": "images/s",
                    "n_gpus": args.gpus, "ms_per_batch": ms / args.steps, "batch": args.batch}))
