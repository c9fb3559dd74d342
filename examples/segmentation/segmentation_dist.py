"""The segmentation U-Net with one process per GPU and no Spark (reference:
examples/segmentation/segmentation_dist.py - MultiWorkerMirroredStrategy driven by TF_CONFIG):
ranks come from torchrun, Adam is fused with the gradient all-reduce over NVLink peer memory.

  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 \
      examples/segmentation/segmentation_dist.py --epochs 2
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import segmentation  # noqa: E402


def main():
  import torch
  import torch.distributed as dist
  from tensorflowonspark_b200.parallel import symm
  args = segmentation.define_flags().parse_args()
  rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
  local = int(os.environ.get("LOCAL_RANK", str(rank)))
  torch.cuda.set_device(local)
  dev = torch.device("cuda", local)
  comm = None
  if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    comm = symm.from_torch_distributed(dev)
  segmentation.train(args, rank, world, comm, dev, is_chief=rank == 0)
  if world > 1:
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
  main()
