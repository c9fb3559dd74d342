"""Fused gradient all-reduce + optimizer step (csrc/optim_comm.cu).

Replaces the collective the reference gets from
``tf.distribute.experimental.MultiWorkerMirroredStrategy`` (reference user code:
examples/mnist/keras/mnist_spark.py:11,55-66; TFoS itself only provides the
cluster spec, tensorflowonspark/TFSparkNode.py:373-384) *and* the Keras
optimizer update that follows it: one kernel per gradient bucket pulls the peer
shards over NVLink, averages, applies SGD / momentum / Adam to the fp32 master
shard and stores the new bf16 weights into every rank's weight buffer.
"""
import torch

from .. import ops

OPTS = {"sgd": 0, "momentum": 1, "adam": 2}


class FusedOptimizer(object):

  def __init__(self, store, comm=None, opt="momentum", lr=0.1, momentum=0.9, weight_decay=0.0,
               beta1=0.9, beta2=0.999, eps=1e-7, num_buckets=1, grid=None):
    self.store, self.comm = store, comm
    self.opt = OPTS[opt]
    dev = store.master.device
    self.world = comm.world if comm is not None else 1
    self.rank = comm.rank if comm is not None else 0
    self.hyper_host = torch.tensor(
        [lr, momentum, weight_decay, 1.0 / self.world, beta1, beta2, eps, 0.0],
        dtype=torch.float32).pin_memory() if torch.cuda.is_available() else None
    self.hyper = torch.tensor([lr, momentum, weight_decay, 1.0 / self.world, beta1, beta2, eps, 0.0],
                              dtype=torch.float32, device=dev)
    self.state1 = torch.zeros_like(store.master) if self.opt != 0 else None
    self.state2 = torch.zeros_like(store.master) if self.opt == 2 else None
    self.step_count = 0
    self.grid = grid or (148 if self.world == 1 else 64)
    # contiguous buckets, boundaries on multiples of 8 elements; bucket 0 is launched first
    n = store.total
    per = (n // num_buckets + 7) // 8 * 8
    self.buckets = []
    b = 0
    while b < n:
      e = min(n, b + per)
      self.buckets.append((b, e))
      b = e
    self._args = []
    for slot, (b, e) in enumerate(self.buckets):
      d = {
          "master": store.master.data_ptr(),
          "state1": self.state1.data_ptr() if self.state1 is not None else 0,
          "state2": self.state2.data_ptr() if self.state2 is not None else 0,
          "hyper": self.hyper.data_ptr(),
          "begin": b, "end": e, "decay_end": store.decay_end,
          "world": self.world, "rank": self.rank, "slot": slot, "opt": self.opt,
          "grid": self.grid, "zero_grads": 0,
          "aux_begin": store.decay_end,
      }
      if comm is None:
        d["grads"] = [store.grads.data_ptr()]
        d["weights"] = [store.weights.data_ptr()]
        d["aux32"] = [store.aux32.data_ptr()]
      else:
        d["grads"] = comm.peer_ptrs("grads")
        d["weights"] = comm.peer_ptrs("weights")
        d["aux32"] = comm.peer_ptrs("aux32")
        d["flags"] = comm.flag_ptrs()
        d["epoch"] = comm.epoch_ptr(slot)
        d["block_counter"] = comm.counter_ptr(slot)
      self._args.append(d)

  def set_lr(self, lr):
    self.hyper[0:1].fill_(float(lr))

  def zero_grads(self):
    self.store.grads.zero_()
    ops.count()

  def step(self, bucket=None):
    """Launch the fused kernel for one bucket (or all).  Stream order guarantees the
    bucket's gradients are complete on this rank; the kernel's own flag barrier
    covers the peers."""
    if self.opt == 2:
      self.step_count += 1
      self.hyper[7:8].add_(1.0)
    todo = range(len(self.buckets)) if bucket is None else [bucket]
    for i in todo:
      ops.K.allreduce_opt(self._args[i])

  def state_dict(self):
    sd = {"step": self.step_count}
    if self.state1 is not None:
      sd["state1"] = self.state1.detach().cpu()
    if self.state2 is not None:
      sd["state2"] = self.state2.detach().cpu()
    sd["hyper"] = self.hyper.detach().cpu()
    return sd

  def load_state_dict(self, sd):
    self.step_count = int(sd.get("step", 0))
    if self.state1 is not None and "state1" in sd:
      self.state1.copy_(sd["state1"])
    if self.state2 is not None and "state2" in sd:
      self.state2.copy_(sd["state2"])
    if "hyper" in sd:
      self.hyper.copy_(sd["hyper"])
