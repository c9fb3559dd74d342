"""One eager (un-graphed) ResNet-50 training step bracketed by cudaProfilerStart/Stop, for

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/step.csv python tools/step_profile.py
  python tools/ncu_summary.py gpurun_out/step.csv

(the per-kernel times are serialised and cold-cache: they explain a step, they are not a bench value)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tensorflowonspark_b200.models import resnet  # noqa: E402


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--batch", type=int, default=256)
  ap.add_argument("--depth", type=int, default=50)
  ap.add_argument("--image", type=int, default=224)
  args = ap.parse_args()
  dev = torch.device("cuda", 0)
  net = resnet.ResNetTrainer(depth=args.depth, batch=args.batch, image=args.image, device=dev)
  x, y = net.synthetic_batch(seed=0)
  net.set_input(x, y)
  for _ in range(2):
    net.step_kernels()
  torch.cuda.synchronize()
  torch.cuda.profiler.start()
  net.step_kernels()
  torch.cuda.synchronize()
  torch.cuda.profiler.stop()


if __name__ == "__main__":
  main()
