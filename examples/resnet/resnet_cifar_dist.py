"""ResNet-56 / CIFAR-10, one process per GPU without Spark - the analogue of the reference's
TF_CONFIG-driven multi-worker entry point (examples/resnet/resnet_cifar_dist.py): ranks come from
the torchrun environment, the control channel is torch.distributed, gradients go through the fused
all-reduce + momentum-SGD kernel over CUDA-IPC symmetric memory.

  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 \
      examples/resnet/resnet_cifar_dist.py --use_synthetic_data --train_steps 200
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import resnet_cifar_main  # noqa: E402


def main():
  import torch
  import torch.distributed as dist
  from tensorflowonspark_b200.parallel import symm
  rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
  local = int(os.environ.get("LOCAL_RANK", str(rank)))
  torch.cuda.set_device(local)
  comm = None
  if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = symm.from_torch_distributed(torch.device("cuda", local))
  resnet_cifar_main.main_fun(sys.argv, resnet_cifar_main.LocalContext(rank, world, comm))
  if world > 1:
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
  main()
