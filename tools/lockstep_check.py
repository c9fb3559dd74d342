"""Lock-step / replica-consistency check of the fused all-reduce + optimizer under CUDA graphs:
one rank sleeps 50 ms per step (SLOW_RANK, default 1); every rank must then take >= 50 ms per
step, and the bf16 weights must stay bit-identical to rank 0's.

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 \
      tools/lockstep_check.py resnet|unet
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from tensorflowonspark_b200.parallel import symm
from tensorflowonspark_b200.models import resnet, unet
which = sys.argv[1]
comm = symm.from_torch_distributed(dev)
if which == "resnet":
  net = resnet.ResNetTrainer(depth=50, batch=16, image=64, device=dev, comm=comm, lr=0.05)
else:
  net = unet.UNetTrainer(batch=8, image=128, classes=3, device=dev, lr=1e-3, comm=comm)
comm.broadcast("weights", root=0); comm.broadcast("aux32", root=0)
x, y = net.synthetic_batch(seed=rank)
if which == "resnet":
  net.set_input(x, y)
else:
  net.set_input(x, (x[..., 0] > 127).int() + (x[..., 1] > 200).int())
net.train_step(); net.capture()
torch.cuda.synchronize(); dist.barrier()
N = 20
t0 = time.time()
for i in range(N):
  if rank == int(os.environ.get("SLOW_RANK", "1")):
    time.sleep(0.05)          # a slow rank must slow everybody down (lock-step)
  net.train_step()
torch.cuda.synchronize()
dt = (time.time() - t0) / N * 1e3
w = net.store.weights.float()
ref = w.clone(); dist.broadcast(ref, 0)
diff = float((w - ref).abs().max())
m = net.store.master.clone()
print("rank {} {}: {:.1f} ms/step (one rank sleeps 50 ms) | max |bf16 weights - rank0's| = {:.3e}".format(
    rank, which, dt, diff), flush=True)
dist.barrier(); dist.destroy_process_group()
