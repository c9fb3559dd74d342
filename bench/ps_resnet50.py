"""ResNet-50 with an asynchronous parameter server: 1 'ps' node holding the fp32 parameters on its
GPU + (N-1) workers that pull with peer loads and push with remote red.add, no barrier
(BASELINE.json config "ResNet-50 async parameter-server (1 PS + 7 workers)").

  python bench/ps_resnet50.py --gpus 2 --batch 128 --steps 20
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main_fun(args, ctx):
  import time
  import torch
  from tensorflowonspark_b200.models import resnet
  torch.cuda.set_device(0)
  if ctx.job_name == "ps":
    seed_net = resnet.ResNetTrainer(depth=50, batch=8, image=32, device="cuda:0")
    _, server = ctx.start_cluster_server(params=seed_net.store.master)
    del seed_net
    server.join()
    return
  _, server = ctx.start_cluster_server()
  ps = server.ps
  net = resnet.ResNetTrainer(depth=50, batch=args.batch, image=224, device="cuda:0")
  st = net.store
  x, y = net.synthetic_batch(seed=ctx.task_index)
  net.set_input(x, y)

  def step():
    ps.pull(out_fp32=st.master, out_bf16=st.weights)          # stale-tolerant read over NVLink
    st.aux32[:st.total - st.decay_end].copy_(st.master[st.decay_end:])
    net._forward(True)
    net._loss(True)
    net._backward()
    ps.push(st.grads, lr=args.lr)                             # applied in the PS GPU's memory

  for _ in range(3):
    step()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(args.steps):
    step()
  e1.record()
  torch.cuda.synchronize()
  with open(os.path.join(args.out, "worker{}.json".format(ctx.task_index)), "w") as f:
    json.dump({"ms": e0.elapsed_time(e1), "loss": float(net.loss_sum)}, f)
  time.sleep(0.5)


if __name__ == "__main__":
  import tempfile
  from tensorflowonspark_b200 import TFCluster
  from tensorflowonspark_b200._spark import SparkConf, SparkContext
  p = argparse.ArgumentParser()
  p.add_argument("--gpus", type=int, default=2)
  p.add_argument("--batch", type=int, default=256)
  p.add_argument("--steps", type=int, default=20)
  p.add_argument("--lr", type=float, default=0.01)
  args = p.parse_args()
  args.out = tempfile.mkdtemp()
  conf = SparkConf().setAppName("ps_bench").set("spark.executor.instances", str(args.gpus)) \
      .set("spark.executor.resource.gpu.amount", "1").set("spark.task.resource.gpu.amount", "1")
  sc = SparkContext(conf=conf)
  cluster = TFCluster.run(sc, main_fun, args, args.gpus, num_ps=1,
                          input_mode=TFCluster.InputMode.TENSORFLOW)
  cluster.shutdown()
  sc.stop()
  res = [json.load(open(os.path.join(args.out, f))) for f in sorted(os.listdir(args.out))]
  ms = max(r["ms"] for r in res)
  workers = args.gpus - 1
  print(json.dumps({"metric": "ResNet-50 async-PS training images/s (1 PS + {} workers)".format(workers),
                    "value": args.batch * args.steps * workers / (ms / 1e3), "unit": "images/s",
                    "n_gpus": args.gpus, "ms_per_step": ms / args.steps,
                    "final_loss": [r["loss"] for r in res]}))
