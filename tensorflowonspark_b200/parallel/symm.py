"""Symmetric (peer-mapped) device memory over CUDA IPC.

Every rank allocates the same named buffers with cudaMalloc, publishes the IPC
handles, and maps every peer's buffer with cudaIpcOpenMemHandle, so kernels can
address ``buf[rank]`` for any rank directly over NVLink.  The handle exchange
rides on whatever control channel is available: ``torch.distributed``
(all_gather_object) under torchrun, or the reservation server's node metadata
when launched through TFCluster (the reference only
exchanges host:port through tensorflowonspark/reservation.py; here the same
rendezvous carries the memory handles).

Also owns the per-rank flag pad used by the device-side barriers in
csrc/optim_comm.cu (slot s: words [32 s, 32 s + 16) = "ready", +16.. = "done").
"""
import torch

from .. import ops

FLAG_SLOTS = 64


class SymmComm(object):

  def __init__(self, rank, world, exchange, device):
    """exchange(obj) -> [obj_0 .. obj_{world-1}] (a collective every rank calls in the same order)."""
    self.rank, self.world, self.exchange = rank, world, exchange
    self.device = torch.device(device)
    self._local = {}   # name -> (ptr, handle, tensor)
    self._peers = {}   # name -> [ptr per rank]
    self._pending = []
    self.flags = self.alloc("__flags__", FLAG_SLOTS * 32, torch.int32)
    self.epochs = torch.zeros(FLAG_SLOTS, dtype=torch.int32, device=self.device)
    self.counters = torch.zeros(FLAG_SLOTS, dtype=torch.int32, device=self.device)

  _DT = {torch.bfloat16: ("bf16", 2), torch.float32: ("f32", 4), torch.int32: ("i32", 4),
         torch.uint8: ("u8", 1)}

  def alloc(self, name, numel, dtype):
    C = ops.C()
    code, size = self._DT[dtype]
    nbytes = max(256, int(numel) * size)
    with torch.cuda.device(self.device):
      ptr, handle = C.symm_alloc(nbytes)
      t = C.tensor_from_ptr(ptr, [int(numel)], code)
    self._local[name] = (ptr, handle, t)
    self._pending.append(name)
    return t

  def _sync_handles(self):
    if not self._pending:
      return
    names = list(self._pending)
    self._pending = []
    mine = {n: self._local[n][1] for n in names}
    everyone = self.exchange(mine)
    C = ops.C()
    with torch.cuda.device(self.device):
      for n in names:
        ptrs = []
        for r in range(self.world):
          if r == self.rank:
            ptrs.append(self._local[n][0])
          else:
            ptrs.append(C.symm_open(everyone[r][n]))
        self._peers[n] = ptrs

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
            C.symm_close(p)
          except Exception:
            pass
    self._peers = {}


def from_torch_distributed(device):
  """SymmComm whose control channel is the default torch.distributed group."""
  import torch.distributed as dist
  rank, world = dist.get_rank(), dist.get_world_size()

  def exchange(obj):
    out = [None] * world
    dist.all_gather_object(out, obj)
    return out

  return SymmComm(rank, world, exchange, device)
