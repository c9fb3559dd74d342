"""Gradient communicator for worker ranks that do NOT share one host (or one NVLink domain).

The symmetric-memory communicator (parallel/symm.py) maps every peer's buffers into every rank
with CUDA IPC / cuMem fabric-less handles: that only works among the GPUs of one machine.  A
cluster that Spark spreads over several hosts - the deployment the reference is built for
(tensorflowonspark/TFSparkNode.py:373-384 hands TensorFlow a multi-host cluster spec) - needs a
transport that crosses the network.  ``GroupComm`` is that fallback with the same surface the
trainers use (``alloc`` / ``broadcast`` / ``barrier`` / ``world`` / ``rank``):

  * gradients are summed with ``torch.distributed.all_reduce`` (NCCL: NVLink inside a host,
    the network between hosts; gloo on CPU-only nodes) bucket by bucket on the communication
    stream, overlapped with the rest of backward exactly like the fused path;
  * the update is the same fused optimizer kernel in its single-rank form (scale 1 / world,
    fp32 master, bf16 weights written in the same pass), run redundantly on every rank - the
    replicas stay bit-identical because every rank applies the same reduced gradient;
  * buffers are ordinary device memory; ``broadcast`` is ``dist.broadcast`` from the chief.

``ctx.gradient_comm()`` picks: one host -> SymmComm (P2P / NVLS kernels), several -> GroupComm.
"""
import logging

import torch

logger = logging.getLogger(__name__)


class GroupComm(object):
  cross_host = True   # FusedOptimizer: all-reduce through the process group, kernel in local form

  def __init__(self, group=None, device=None):
    import torch.distributed as dist
    if not dist.is_initialized():
      raise RuntimeError("GroupComm needs an initialised torch.distributed process group "
                         "(ctx.init_process_group())")
    self.group = group if group is not None else dist.group.WORLD
    self.world = dist.get_world_size(self.group)
    self.rank = dist.get_rank(self.group)
    self.device = torch.device(device) if device is not None else torch.device("cpu")
    self.bufs = {}
    self.nvls = False
    logger.info("gradient communicator: torch.distributed %s group, rank %d/%d (cross-host path)",
                dist.get_backend(self.group), self.rank, self.world)

  # ---- the surface shared with SymmComm ---------------------------------------------------
  def alloc(self, name, numel, dtype, multicast=False):
    if name in self.bufs:
      raise ValueError("buffer {!r} already allocated".format(name))
    t = torch.zeros(int(numel), dtype=dtype, device=self.device)
    self.bufs[name] = t
    return t

  def local(self, name):
    return self.bufs[name]

  def _global(self, group_rank):
    import torch.distributed as dist
    return dist.get_global_rank(self.group, group_rank) if self.group is not dist.group.WORLD else group_rank

  def broadcast(self, name, root=0):
    """The root's copy of ``name`` wins on every rank (initial variables from the chief)."""
    import torch.distributed as dist
    dist.broadcast(self.bufs[name], src=self._global(root), group=self.group)

  def all_reduce(self, tensor):
    """In-place SUM over the group, ordered on the current stream."""
    import torch.distributed as dist
    dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=self.group)

  def barrier(self):
    import torch.distributed as dist
    if self.device.type == "cuda":
      dist.barrier(group=self.group, device_ids=[self.device.index or 0])
    else:
      dist.barrier(group=self.group)

  def close(self):
    self.bufs.clear()

  # ---- what only peer-mapped memory can do ------------------------------------------------
  def peer_ptrs(self, name):
    raise RuntimeError("peer pointers need symmetric memory: the ranks of this job do not share a "
                       "host (GroupComm); use ctx.symmetric_comm() on a single-host group")

  flag_ptrs = mc_ptr = peer_ptrs
