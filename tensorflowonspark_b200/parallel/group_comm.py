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

``ctx.gradient_comm()`` picks: one host -> SymmComm (P2P / NVLS kernels); several hosts with the
same number of GPU workers -> HierComm (below: NVLink inside the host, NCCL between hosts on the
shards); anything else -> GroupComm.
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


class HierComm(object):
  """Two-level gradient communicator for several multi-GPU hosts.

  ``local``: the SymmComm of the GPUs of THIS host (None on a one-GPU host), ``inter``: the
  torch.distributed group of the ranks that have this rank's local index on every host.  The
  all-reduce of a bucket then is (FusedOptimizer, hierarchical mode):

    1. intra-host reduce-scatter - allreduce_opt kernel, PHASE 1: every rank sums its 1 / L shard
       of the bucket over the L local peers through peer loads (NVLink) into its own buffer;
    2. inter-host all-reduce of that shard only (NCCL over the network: 1 / L of the bytes per
       rank, all L ranks of a host driving their own NICs concurrently);
    3. update + intra-host all-gather - PHASE 2: optimizer on the shard (scale 1 / global world,
       state sharded like on one host), new bf16 weights stored to every local peer.

  Network bytes per rank and step are 1 / L of what a flat all-reduce of the whole vector moves
  through each process, and the optimizer state stays sharded (ZeRO-1-like) inside the host."""

  hierarchical = True

  def __init__(self, local, inter, rank, world, device=None, local_rank=0, local_world=1):
    import torch.distributed as dist
    self.local, self.inter = local, inter
    self.rank, self.world = int(rank), int(world)
    self.local_rank = local.rank if local is not None else int(local_rank)
    self.local_world = local.world if local is not None else int(local_world)
    self.hosts = self.world // max(1, self.local_world)
    self.device = torch.device(device) if device is not None else (
        local.device if local is not None else torch.device("cpu"))
    self.nvls = False
    self.bufs = {}
    self._dist = dist
    logger.info("gradient communicator: %d host(s) x %d GPU(s), rank %d (local %d): NVLink "
                "reduce-scatter / all-gather inside the host, %s all-reduce of the shards between",
                self.hosts, self.local_world, self.rank, self.local_rank,
                dist.get_backend(inter) if inter is not None else "no")

  # ---- buffers: symmetric inside the host --------------------------------------------------
  def alloc(self, name, numel, dtype, multicast=False):
    if self.local is not None:
      t = self.local.alloc(name, numel, dtype, multicast=False)
    else:
      t = torch.zeros(int(numel), dtype=dtype, device=self.device)
    self.bufs[name] = t
    return t

  def peer_ptrs(self, name):
    return self.local.peer_ptrs(name) if self.local is not None else [self.bufs[name].data_ptr()]

  def flag_ptrs(self):
    return self.local.flag_ptrs()

  def epoch_ptr(self, slot):
    return self.local.epoch_ptr(slot)

  def counter_ptr(self, slot):
    return self.local.counter_ptr(slot)

  def mc_ptr(self, name):
    return 0

  # ---- collectives -------------------------------------------------------------------------
  def all_reduce_inter(self, tensor):
    """SUM of ``tensor`` over the hosts (the ranks sharing this local index), in place."""
    if self.inter is not None and self.hosts > 1:
      self._dist.all_reduce(tensor, op=self._dist.ReduceOp.SUM, group=self.inter)

  def broadcast(self, name, root=0):
    """Global rank ``root``'s copy of ``name`` wins everywhere: across the hosts among the ranks
    with the root's local index, then inside every host from that local index."""
    if root != 0:
      raise ValueError("HierComm.broadcast: the root is the chief (global rank 0)")
    if self.local_rank == 0 and self.inter is not None and self.hosts > 1:
      src = self._dist.get_global_rank(self.inter, 0)
      self._dist.broadcast(self.bufs[name], src=src, group=self.inter)
      if self.device.type == "cuda":
        torch.cuda.current_stream(self.device).synchronize()
    if self.local is not None:
      self.local.barrier()              # the local root's copy is complete before anyone pulls it
      self.local.broadcast(name, root=0)

  def barrier(self):
    if self.local is not None:
      self.local.barrier()
    if self.inter is not None and self.hosts > 1:
      if self.device.type == "cuda":
        self._dist.barrier(group=self.inter, device_ids=[self.device.index or 0])
      else:
        self._dist.barrier(group=self.inter)

  def close(self):
    if self.local is not None:
      self.local.close()
    self.bufs.clear()
