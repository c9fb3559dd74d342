"""Process-group bring-up from a TFNodeContext.

The node runtime (TFSparkNode._export_dist_env) derives ``MASTER_ADDR/MASTER_PORT/RANK/
WORLD_SIZE`` from the cluster spec - rank 0 is the chief/master (else worker 0) and the
rendezvous port is the port that node *reserved* during registration, which is exactly what
the reference reserved it for (tensorflowonspark/TFSparkNode.py:343-352, there to hand to
``tf.train.Server``).  NCCL is used when the node owns a GPU, gloo otherwise.
"""
import datetime
import logging
import os

logger = logging.getLogger(__name__)


def init_from_ctx(ctx, backend=None, timeout_s=1800):
  import torch
  import torch.distributed as dist
  if ctx.rank < 0:
    raise RuntimeError("{}:{} is not a worker rank".format(ctx.job_name, ctx.task_index))
  if dist.is_initialized():
    return dist.group.WORLD
  if getattr(ctx, "tmp_socket", None) is not None:
    ctx.release_port()  # the reserved port becomes the rendezvous port
  use_cuda = bool(ctx.gpus) and torch.cuda.is_available()
  backend = backend or ("nccl" if use_cuda else "gloo")
  kwargs = {}
  if use_cuda:
    torch.cuda.set_device(0)
    if backend == "nccl":
      kwargs["device_id"] = torch.device("cuda", 0)
  logger.info("init_process_group(%s) rank %d/%d at %s:%s", backend, ctx.rank, ctx.world_size,
              os.environ.get("MASTER_ADDR"), os.environ.get("MASTER_PORT"))
  dist.init_process_group(backend, rank=ctx.rank, world_size=ctx.world_size,
                          timeout=datetime.timedelta(seconds=timeout_s), **kwargs)
  return dist.group.WORLD


def new_group(ctx, ranks, backend=None):
  """torch.distributed.new_group over ``ranks`` (the default group is joined first if needed)."""
  import torch.distributed as dist
  if not dist.is_initialized():
    init_from_ctx(ctx)
  ranks = sorted(set(int(r) for r in ranks))
  if not ranks or ranks[0] < 0 or ranks[-1] >= ctx.world_size:
    raise ValueError("ranks {} outside the job's 0..{}".format(ranks, ctx.world_size - 1))
  g = dist.new_group(ranks=ranks, backend=backend)
  return g if ctx.rank in ranks else None


# communicators created so far in this process, per (cluster, scope): every member creates its
# communicators in the same order (it is a collective), so the generation number agrees across
# ranks and keeps the board tags of two communicators of one job apart
_generations = {}


def symm_from_ctx(ctx, ranks=None):
  """SymmComm whose handle exchange runs over the reservation server's key/value board.
  ``ranks`` (optional): a sub-group of worker ranks; the communicator's rank / world are then the
  position in / size of that group, and its exchanges use a board namespace of their own."""
  from .. import reservation
  from . import symm
  hosts = ctx.worker_hosts() if hasattr(ctx, "worker_hosts") else []
  span = sorted(set(hosts[r] for r in (ranks if ranks is not None else range(len(hosts)))
                    if 0 <= int(r) < len(hosts)))
  if len(span) > 1:
    raise RuntimeError(
        "symmetric memory maps peer GPUs through CUDA IPC and needs every member on one host, but "
        "the group spans {}; use ctx.gradient_comm() (NCCL between hosts) or pass ranks= of one "
        "host".format(span))
  if ranks is None:
    members = list(range(ctx.world_size))
  else:
    members = sorted(set(int(r) for r in ranks))
    if ctx.rank not in members:
      raise ValueError("rank {} is not a member of the group {}".format(ctx.rank, members))
    if members[0] < 0 or members[-1] >= ctx.world_size:
      raise ValueError("ranks {} outside the job's 0..{}".format(members, ctx.world_size - 1))
  me, size = members.index(ctx.rank), len(members)
  scope = "all" if ranks is None else "-".join(str(r) for r in members)
  client = reservation.Client(ctx.server_addr)
  counter = [0]
  gen = _generations[(ctx.cluster_id, scope)] = _generations.get((ctx.cluster_id, scope), 0) + 1

  def exchange(obj):
    # unique per communicator (generation) and per exchange (counter); the board entries are
    # consumed by the all_gather itself, so even a re-created process that restarts its
    # generation count cannot read a peer's stale (possibly freed) IPC handle
    counter[0] += 1
    tag = "symm/{}/{}/g{}/{}".format(ctx.cluster_id, scope, gen, counter[0])
    return client.all_gather(tag, me, size, obj)

  return symm.SymmComm(me, size, exchange, ctx.device)


def hier_layout(hosts, rank):
  """Two-level layout of worker ranks: ``hosts[r]`` is the host of rank r.  Returns
  ``(local_ranks, local_index, inter_groups)`` - the ranks of ``rank``'s host, its position among
  them, and for every local index k the ranks holding index k on each host (in first-seen host
  order) - or None when the hosts do not all run the same number of workers."""
  by_host = {}
  for r, h in enumerate(hosts):
    by_host.setdefault(h, []).append(r)
  sizes = set(len(v) for v in by_host.values())
  if len(sizes) != 1:
    return None
  per = sizes.pop()
  mine = by_host[hosts[rank]]
  inter = [[by_host[h][k] for h in by_host] for k in range(per)]
  return mine, mine.index(rank), inter


def gradient_comm_from_ctx(ctx):
  """The communicator a trainer's fused optimizer runs on:

    * every worker on one host with a GPU  -> SymmComm (P2P / NVLS kernels, parallel/symm.py);
    * several hosts with the same number of GPU workers each -> group_comm.HierComm (NVLink
      reduce-scatter / all-gather kernels inside the host, NCCL between the hosts on the shards);
    * anything else (uneven hosts, no GPU) -> group_comm.GroupComm (flat torch.distributed).

  ``TFOS_GRADIENT_COMM=group`` forces the flat fallback (e.g. GPUs of one host without P2P
  access), ``=hier`` insists on the two-level path (raises when the hosts are uneven)."""
  import torch
  force = os.environ.get("TFOS_GRADIENT_COMM", "auto")
  use_cuda = bool(ctx.gpus) and torch.cuda.is_available()
  if force not in ("group", "hier") and use_cuda and ctx.single_host:
    return symm_from_ctx(ctx)
  from . import group_comm
  init_from_ctx(ctx)
  layout = hier_layout(ctx.worker_hosts(), ctx.rank) if (use_cuda and force != "group") else None
  if layout is None:
    if force == "hier":
      raise RuntimeError("TFOS_GRADIENT_COMM=hier needs GPU workers, the same number on every host; "
                         "hosts: {}".format(ctx.worker_hosts()))
    return group_comm.GroupComm(device=ctx.device)
  import torch.distributed as dist
  mine, index, inter_groups = layout
  inter = None
  for k, ranks in enumerate(inter_groups):      # collective: every rank creates every group
    g = dist.new_group(ranks=ranks) if len(ranks) > 1 else None
    if k == index:
      inter = g
  local = symm_from_ctx(ctx, ranks=mine) if len(mine) > 1 else None
  return group_comm.HierComm(local, inter, ctx.rank, ctx.world_size, device=ctx.device,
                             local_rank=index, local_world=len(mine))
