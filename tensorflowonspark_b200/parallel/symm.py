"""Symmetric (peer-mapped) device memory: CUDA IPC, or cuMem VMM + NVLS multicast.

Every rank allocates the same named buffers with cudaMalloc, publishes the IPC
handles, and maps every peer's buffer with cudaIpcOpenMemHandle, so kernels can
address ``buf[rank]`` for any rank directly over NVLink.  Buffers requested with
``multicast=True`` (gradients, weights: the operands of the fused all-reduce) are
instead created with the virtual-memory-management API (csrc/vmm.cc), shared as
file descriptors (parallel/fdshare.py) and additionally bound into an NVLink-switch
multicast object: ``mc_ptr(name)`` is an address on which ``multimem.ld_reduce``
sums the same element of every rank in the switch and ``multimem.st`` writes it to
all of them (selected when every rank reports multicast
support and the job has at least 4 ranks - ``TFOS_NVLS=1`` / ``0`` force it on / off).  The handle exchange
rides on whatever control channel is available: ``torch.distributed``
(all_gather_object) under torchrun, or the reservation server's node metadata
when launched through TFCluster (the reference only
exchanges host:port through tensorflowonspark/reservation.py; here the same
rendezvous carries the memory handles).

Also owns the per-rank flag pad used by the device-side barriers in
csrc/optim_comm.cu (slot s: words [32 s, 32 s + 16) = "ready", +16.. = "done").
"""
import logging
import os

import torch

from .. import ops

logger = logging.getLogger(__name__)

FLAG_SLOTS = 64
# default of TFOS_NVLS: "auto" = multicast from 4 ranks up (verified on 2 and 8 GPUs this round:
# profiles/r2_8gpu/ - stand-alone 0.40 vs 0.67 ms for 25.6 M parameters, 0.31 vs 0.47 ms of
# communication cost per ResNet-50 step inside the captured graph)
NVLS_DEFAULT = "auto"


class SymmComm(object):

  def __init__(self, rank, world, exchange, device):
    """exchange(obj) -> [obj_0 .. obj_{world-1}] (a collective every rank calls in the same order)."""
    self.rank, self.world, self.exchange = rank, world, exchange
    self.device = torch.device(device)
    self._local = {}   # name -> (ptr, handle, tensor)
    self._peers = {}   # name -> [ptr per rank]
    self._mc = {}      # name -> multicast VA (NVLS)
    self._vmm = {}     # name -> (va, size, granularity) of VMM allocations
    self._pending = []
    self._fdsrv = None
    # a rank that waits for a dead peer gives up after TFOS_FLAG_TIMEOUT_MS (default 20 s): the
    # kernel traps, the CUDA error surfaces in the training loop and reaches the driver through
    # the node's error queue instead of hanging the job
    if self.device.type == "cuda" and os.environ.get("TFOS_FLAG_TIMEOUT_MS"):
      with torch.cuda.device(self.device):
        ops.C().set_flag_timeout_ms(float(os.environ["TFOS_FLAG_TIMEOUT_MS"]))
    self.nvls = self._probe_nvls()
    self.flags = self.alloc("__flags__", FLAG_SLOTS * 32, torch.int32)
    self.epochs = torch.zeros(FLAG_SLOTS, dtype=torch.int32, device=self.device)
    self.counters = torch.zeros(FLAG_SLOTS, dtype=torch.int32, device=self.device)

  _DT = {torch.bfloat16: ("bf16", 2), torch.float32: ("f32", 4), torch.int32: ("i32", 4),
         torch.uint8: ("u8", 1)}

  # names that take part in the fused all-reduce: allocated on the multicast path when the box
  # supports it (ParamStore.finalize / FusedOptimizer call alloc() with these names)
  MULTICAST_NAMES = ("grads", "weights", "aux32")

  def _probe_nvls(self):
    """True when EVERY rank can use VMM + multicast (a collective decision: one exchange)."""
    # TFOS_NVLS: 0 = peer-to-peer only, 1 = multicast whenever the box supports it, auto (default)
    # = multicast from 4 ranks up.  A multimem.ld_reduce makes the switch read EVERY member's copy
    # - the requester's own included - so each GPU sends world/(world-1) times the bytes of the
    # peer-to-peer pull: 2x at 2 ranks (measured: 0.50 vs 0.37 ms for 25.6 M parameters), 1.14x at
    # 8, where the 8x smaller ingress and the 8x fewer load instructions win.
    mode = os.environ.get("TFOS_NVLS", NVLS_DEFAULT)
    if self.world < 2 or mode == "0" or (mode == "auto" and self.world < 4):
      return False
    try:
      with torch.cuda.device(self.device):
        info = dict(ops.C().vmm_info(self.world))
    except Exception as e:   # driver without the VMM entry points
      info = {"vmm": False, "multicast": False, "error": str(e)}
    mine = {"ok": bool(info.get("vmm")) and bool(info.get("multicast")) and
            int(info.get("mc_granularity") or 0) > 0,
            "g": max(int(info.get("granularity") or 0), int(info.get("mc_granularity") or 0))}
    every = self.exchange(mine)
    ok = all(e["ok"] for e in every)
    self._granularity = max(e["g"] for e in every) if ok else 0
    if self.rank == 0:
      logger.info("symmetric memory: %s", "cuMem VMM + NVLS multicast (granularity {} KiB)".format(
          self._granularity >> 10) if ok else "CUDA IPC peer mappings (no multicast: {})".format(info))
    return ok

  def alloc(self, name, numel, dtype, multicast=None):
    C = ops.C()
    code, size = self._DT[dtype]
    nbytes = max(256, int(numel) * size)
    if multicast is None:
      multicast = name in self.MULTICAST_NAMES
    with torch.cuda.device(self.device):
      if multicast and self.nvls:
        if self._fdsrv is None:
          from .fdshare import FdServer
          self._fdsrv = FdServer()
        va, mapped, fd = C.vmm_alloc(nbytes, self._granularity)
        self._fdsrv.register("mem:" + name, fd)
        self._vmm[name] = (va, mapped)
        ptr, handle = va, {"vmm": True, "size": mapped, "srv": self._fdsrv.address}
      else:
        ptr, handle = C.symm_alloc(nbytes)
      t = C.tensor_from_ptr(ptr, [int(numel)], code)
    self._local[name] = (ptr, handle, t)
    self._pending.append(name)
    return t

  def _sync_handles(self):
    if not self._pending:
      return
    from .fdshare import fetch_fd
    names = list(self._pending)
    self._pending = []
    mine = {n: self._local[n][1] for n in names}
    everyone = self.exchange(mine)
    C = ops.C()
    with torch.cuda.device(self.device):
      for n in names:
        ptrs = []
        for r in range(self.world):
          h = everyone[r][n]
          if r == self.rank:
            ptrs.append(self._local[n][0])
          elif isinstance(h, dict) and h.get("vmm"):
            fd = fetch_fd(h["srv"], "mem:" + n)
            ptrs.append(C.vmm_import(fd, int(h["size"]), self._granularity))
          else:
            ptrs.append(C.symm_open(h))
        self._peers[n] = ptrs
      # NVLS: one multicast object per VMM buffer.  create (rank 0) -> import -> every rank adds
      # its device -> [all added] -> every rank binds its memory -> [all bound] -> map.
      vmm_names = [n for n in names if n in self._vmm]
      if vmm_names:
        ids = {}
        if self.rank == 0:
          for n in vmm_names:
            mc_id, fd = C.mc_create(self.world, self._vmm[n][1])
            self._fdsrv.register("mc:" + n, fd)
            ids[n] = mc_id
        roots = self.exchange({"srv": self._fdsrv.address if self.rank == 0 else None})
        if self.rank != 0:
          for n in vmm_names:
            ids[n] = C.mc_import(fetch_fd(roots[0]["srv"], "mc:" + n))
        for n in vmm_names:
          C.mc_add_device(ids[n])
        self.exchange({"added": True})
        for n in vmm_names:
          C.mc_bind(ids[n], self._vmm[n][0], self._vmm[n][1])
        self.exchange({"bound": True})
        for n in vmm_names:
          self._mc[n] = C.mc_map(ids[n], self._vmm[n][1], self._granularity)
        torch.cuda.synchronize(self.device)
        self.exchange({"mapped": True})

  def mc_ptr(self, name):
    """Multicast (NVLS) address of buffer ``name`` or 0 when it lives on the IPC path."""
    self._sync_handles()
    return int(self._mc.get(name, 0))

  def peer_ptrs(self, name):
    self._sync_handles()
    return list(self._peers[name])

  def flag_ptrs(self):
    return self.peer_ptrs("__flags__")

  def epoch_ptr(self, slot):
    return self.epochs.data_ptr() + 4 * slot

  def counter_ptr(self, slot):
    return self.counters.data_ptr() + 4 * slot

  def _ctl(self, slot):
    return {"world": self.world, "rank": self.rank, "slot": slot, "flags": self.flag_ptrs(),
            "epoch": self.epoch_ptr(slot), "block_counter": self.counter_ptr(slot)}

  def barrier(self, slot=FLAG_SLOTS - 1):
    """Device-side barrier across ranks on the current stream (no host sync)."""
    ops.K.flag_barrier(self._ctl(slot))

  def broadcast(self, name, root=0, slot=FLAG_SLOTS - 2, grid=64):
    """Every rank pulls the root's copy of buffer ``name`` over NVLink (startup variable broadcast)."""
    d = self._ctl(slot)
    d["root"] = root
    d["bufs"] = self.peer_ptrs(name)
    t = self._local[name][2]
    d["bytes"] = t.numel() * t.element_size() // 16 * 16
    d["grid"] = grid
    ops.K.bcast_pull(d)

  def close(self):
    C = ops.C()
    for n, ptrs in self._peers.items():
      for r, p in enumerate(ptrs):
        if r != self.rank:
          try:
            (C.vmm_free if n in self._vmm else C.symm_close)(p)
          except Exception:
            pass
    for va in self._mc.values():
      try:
        C.vmm_free(va)
      except Exception:
        pass
    self._peers, self._mc = {}, {}
    if self._fdsrv is not None:
      self._fdsrv.close()
      self._fdsrv = None


def from_torch_distributed(device):
  """SymmComm whose control channel is the default torch.distributed group."""
  import torch.distributed as dist
  rank, world = dist.get_rank(), dist.get_world_size()

  def exchange(obj):
    out = [None] * world
    dist.all_gather_object(out, obj)
    return out

  return SymmComm(rank, world, exchange, device)
