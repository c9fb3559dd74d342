"""ResNet training launched on Spark executors, InputMode.TENSORFLOW, synthetic ImageNet-shaped
data, sync data parallelism with the fused all-reduce + momentum-SGD kernel.

Reference shape: examples/resnet/resnet_cifar_spark.py:17-22 (argv pass-through to a main_fun
that parses its own flags) + resnet_cifar_dist.py:196-257 (strategy scope, synthetic data option,
piecewise LR).  ``--model resnet50`` is BASELINE.json's headline config; ``--model resnet56`` is the
CIFAR-10 network the reference actually trains.

  python examples/resnet/resnet_spark.py --cluster_size 8 --model resnet50 --batch_size 256 \
      --train_steps 100
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def parse(argv):
  p = argparse.ArgumentParser()
  p.add_argument("--cluster_size", type=int, default=1)
  p.add_argument("--model", default="resnet50", choices=["resnet50", "resnet101", "resnet56"])
  p.add_argument("--batch_size", type=int, default=256, help="per-GPU batch")
  p.add_argument("--train_steps", type=int, default=50)
  p.add_argument("--image", type=int, default=None)
  p.add_argument("--epochs_per_step", type=float, default=0.0)
  p.add_argument("--model_dir", default=None)
  p.add_argument("--save_steps", type=int, default=0)
  p.add_argument("--no_graph", action="store_true")
  p.add_argument("--tensorboard", action="store_true")
  p.add_argument("--metrics", default=None, help="JSONL file for per-step metrics")
  p.add_argument("--metrics_port", type=int, default=None,
                 help="serve the same fields on a Prometheus /metrics endpoint (port + rank; 0: any free port)")
  return p.parse_args(argv)


def main_fun(argv, ctx):
  import time
  import torch
  from tensorflowonspark_b200.models import resnet
  from tensorflowonspark_b200.utils import checkpoint, metrics
  args = parse(argv[1:])
  from tensorflowonspark_b200.utils import fault
  torch.cuda.set_device(0)  # the node runtime made the assigned GPU device 0
  comm = ctx.gradient_comm() if ctx.world_size > 1 else None   # one host: P2P/NVLS kernels; several: NCCL
  B = args.batch_size
  if args.model == "resnet56":
    net = resnet.CifarResNetTrainer(depth=56, batch=B, device="cuda:0", comm=comm,
                                    lr=resnet.piecewise_lr(0, B * ctx.world_size))
  else:
    depth = int(args.model.replace("resnet", ""))
    net = resnet.ResNetTrainer(depth=depth, batch=B, image=args.image or 224, device="cuda:0",
                               comm=comm, lr=0.1 * B * ctx.world_size / 256.0)
  if comm is not None:
    comm.broadcast("weights", root=0)   # the chief's initial variables win
    comm.broadcast("aux32", root=0)
  start = 0
  if args.model_dir:
    start, state = checkpoint.load(ctx.absolute_path(args.model_dir))
    if state is not None:
      net.load_state_dict(state["params"])      # parameters + batch-norm running statistics
      net.optim.load_state_dict(state["optim"])
      print("resumed from step", start)
  x, y = net.synthetic_batch(seed=ctx.rank)
  net.set_input(x, y)
  net.train_step()
  if not args.no_graph:
    net.capture()
  exporter = None
  if args.metrics_port is not None and metrics.Exporter.available():
    exporter = metrics.Exporter(rank=ctx.rank, port=args.metrics_port + ctx.rank if args.metrics_port else 0)
    print("rank {} metrics at {}".format(ctx.rank, exporter.url()))
  log = metrics.StepLogger(args.metrics, rank=ctx.rank, exporter=exporter) if args.metrics else None
  events = None
  if args.model_dir and ctx.is_chief:     # TensorBoard scalars next to the checkpoints
    from tensorflowonspark_b200.utils import summary
    events = summary.SummaryWriter(ctx.absolute_path(args.model_dir))
  torch.cuda.synchronize()
  t0 = time.time()
  for step in range(start, start + args.train_steps):
    if args.epochs_per_step:
      net.set_lr(resnet.piecewise_lr(int(step * args.epochs_per_step), B * ctx.world_size))
    fault.maybe_inject(ctx.rank, step)      # TFOS_FAULT_INJECT (resilience tests), no-op otherwise
    loss = net.train_step()
    if (step + 1) % 10 == 0:
      torch.cuda.synchronize()
      dt = time.time() - t0
      rate = 10 * B * ctx.world_size / dt
      if ctx.is_chief:
        print("step {:5d} loss {:.4f}  {:.0f} images/s".format(step + 1, float(loss), rate))
      if log:
        log.log(step=step + 1, loss=float(loss), images_per_s=rate, step_ms=dt * 100)
      if events:
        events.add_scalars({"loss": float(loss), "images_per_s": rate}, step + 1)
      if exporter and not log:
        exporter.update(step=step + 1, loss=float(loss), images_per_s=rate, step_ms=dt * 100)
      t0 = time.time()
    if args.model_dir and args.save_steps and (step + 1) % args.save_steps == 0 and ctx.is_chief:
      checkpoint.save(ctx.absolute_path(args.model_dir), step + 1,
                      {"params": net.state_dict(), "optim": net.optim.state_dict()})
  torch.cuda.synchronize()
  if events:
    events.close()
  if exporter:
    exporter.close()


if __name__ == "__main__":
  from tensorflowonspark_b200 import TFCluster
  from tensorflowonspark_b200._spark import SparkConf, SparkContext

  args = parse(sys.argv[1:])
  conf = SparkConf().setAppName("resnet_spark") \
      .set("spark.executor.instances", str(args.cluster_size)) \
      .set("spark.executor.resource.gpu.amount", "1").set("spark.task.resource.gpu.amount", "1")
  sc = SparkContext(conf=conf)
  cluster = TFCluster.run(sc, main_fun, sys.argv, args.cluster_size, num_ps=0,
                          input_mode=TFCluster.InputMode.TENSORFLOW, master_node="chief",
                          tensorboard=args.tensorboard, log_dir=args.model_dir)
  cluster.shutdown()
  sc.stop()
