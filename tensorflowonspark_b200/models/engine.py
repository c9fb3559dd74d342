"""Explicit-graph training engine on the native kernels.

There is no autograd and no tracing compiler here: a model is a list of layer
objects with static device buffers, ``forward()``/``backward()`` enqueue native
kernels on the current CUDA stream in a fixed order, and the whole step can be
captured once into a CUDA graph.  Parameters live in three flat buffers
(fp32 master, bf16 compute copy, fp32 gradients) so that the data-parallel
all-reduce and the optimizer are one fused kernel over contiguous ranges
(parallel/fused_optim.py).

This is the B200-native stand-in for what the reference delegates to Keras /
tf.distribute in its example programs (e.g. examples/mnist/keras/mnist_spark.py:
11-66, examples/resnet/resnet_cifar_dist.py:196-257).
"""
import math

import torch

from .. import ops
from ..ops import igemm


def _pad8(n):
  return (n + 7) // 8 * 8


class ParamStore(object):
  """Flat parameter storage.  Decayed parameters (conv / dense weights) come first,
  then the non-decayed ones (batch-norm scale/offset, biases)."""

  def __init__(self):
    self._specs = []  # (name, shape, decay, init_fn)
    self.finalized = False

  def register(self, name, shape, decay, init):
    assert not self.finalized
    spec = {"name": name, "shape": tuple(shape), "decay": bool(decay), "init": init}
    self._specs.append(spec)
    return spec

  def finalize(self, device, alloc=None, seed=1234):
    """alloc(nbytes_name, numel, dtype) -> tensor lets the caller place the buffers in
    symmetric (peer-mapped) memory; default is ordinary device memory."""
    order = [s for s in self._specs if s["decay"]] + [s for s in self._specs if not s["decay"]]
    off = 0
    for s in order:
      s["offset"] = off
      s["numel"] = int(math.prod(s["shape"]))
      off += _pad8(s["numel"])
    self.decay_end = sum(_pad8(s["numel"]) for s in order if s["decay"])
    self.total = _pad8(off)
    self.order = order
    if alloc is None:
      alloc = lambda name, n, dt: torch.zeros(n, dtype=dt, device=device)  # noqa: E731
    self.master = alloc("master", self.total, torch.float32)  # symmetric too: broadcastable
    self.weights = alloc("weights", self.total, torch.bfloat16)
    self.grads = alloc("grads", self.total, torch.float32)
    # fp32 replica of the non-decayed tail (BN scale/offset are consumed in fp32)
    self.aux32 = alloc("aux32", max(8, self.total - self.decay_end), torch.float32)
    gen = torch.Generator(device="cpu")
    gen.manual_seed(seed)
    for s in order:
      v = s["init"](s["shape"], gen).to(torch.float32).reshape(-1)
      self.master[s["offset"]:s["offset"] + s["numel"]].copy_(v)
    self.weights.copy_(self.master)
    self.aux32[:self.total - self.decay_end].copy_(self.master[self.decay_end:])
    self.finalized = True
    self.by_name = {s["name"]: s for s in order}

  def _view(self, buf, s, base=0):
    return buf[s["offset"] - base:s["offset"] - base + s["numel"]].view(s["shape"])

  def w(self, s):
    return self._view(self.weights, s)

  def g(self, s):
    return self._view(self.grads, s)

  def m(self, s):
    return self._view(self.master, s)

  def f32(self, s):
    """fp32 value of a non-decayed parameter as seen by every rank."""
    assert not s["decay"]
    return self._view(self.aux32, s, self.decay_end)

  def state_dict(self):
    # with world > 1 the fused all-reduce + optimizer kernel keeps the fp32 master current only
    # for this rank's own shard of every bucket: FusedOptimizer registers assemble(), which pulls
    # the peers' shards over NVLink before anything is saved or exported
    assemble = getattr(self, "_assemble", None)
    if assemble is not None:
      assemble()
    return {s["name"]: self.m(s).detach().cpu().clone() for s in self.order}

  def load_state_dict(self, sd):
    for s in self.order:
      if s["name"] in sd:
        self.m(s).copy_(sd[s["name"]].to(self.master.device, torch.float32).view(s["shape"]))
    self.weights.copy_(self.master)
    self.aux32[:self.total - self.decay_end].copy_(self.master[self.decay_end:])


def he_normal(fan_in):
  std = math.sqrt(2.0 / fan_in)
  return lambda shape, gen: torch.randn(shape, generator=gen) * std


def constant(v):
  return lambda shape, gen: torch.full(shape, float(v))


def normal(std):
  return lambda shape, gen: torch.randn(shape, generator=gen) * std


# Weight gradients are needed by nobody until the optimizer runs, read only buffers that the rest
# of the backward pass never rewrites, and are bound by the L2 -> SM fabric (profiles/): issued on
# a side stream they overlap the HBM-bound batch-norm kernels of the layers below instead of
# standing in the critical path.  A trainer opts in with set_wgrad_stream(); the optimizer joins
# the stream before it consumes the gradients (FusedOptimizer.add_producer_stream).
_wgrad_stream = None


def set_wgrad_stream(stream):
  global _wgrad_stream
  _wgrad_stream = stream


def _run_wgrad(layer):
  """wgrad (+ bias gradient) of a Conv / Dense layer, on the side stream when one is set."""
  def body():
    layer.wgrad.run()
    if layer.sbias is not None:
      ops.K.colsum(layer.dy, layer.store.g(layer.sbias))
  side = _wgrad_stream
  if side is None:
    return body()
  main = torch.cuda.current_stream(side.device)
  ev = getattr(layer, "_ev_dy", None)
  if ev is None:
    ev = layer._ev_dy = torch.cuda.Event()
  ev.record(main)          # dy (and this step's zeroed gradient buffer) are final here
  side.wait_event(ev)
  with torch.cuda.stream(side):
    body()


class StatsArena(object):
  """One flat fp32 buffer holding the (sum, sumsq) accumulators of every batch norm of a trainer:
  a single memset per step replaces a clearing pass per layer."""

  def __init__(self, device, capacity=1 << 17):
    self.buf = torch.zeros(capacity, dtype=torch.float32, device=device)
    self.used = 0

  def take(self, n):
    n8 = (n + 7) // 8 * 8
    if self.used + n8 > self.buf.numel():
      raise RuntimeError("StatsArena capacity exceeded")
    out = self.buf[self.used:self.used + n]
    self.used += n8
    return out

  def zero(self):
    self.buf[:self.used].zero_()
    ops.count()


class RunningArena(object):
  """The running mean / variance of every batch norm of a trainer in ONE flat fp32 buffer
  ([mean_0 | var_0 | mean_1 | var_1 ...], each padded to 8 elements): the non-trainable model
  state is then a single tensor that checkpoints save, exports serve and the parameter server
  hosts as the tail of its vector (parallel/ps.py)."""

  def __init__(self, device, capacity=1 << 17):
    self.buf = torch.zeros(capacity, dtype=torch.float32, device=device)
    self.used = 0

  def take(self, C):
    C8 = (C + 7) // 8 * 8
    if self.used + 2 * C8 > self.buf.numel():
      raise RuntimeError("RunningArena capacity exceeded")
    mean = self.buf[self.used:self.used + C]
    var = self.buf[self.used + C8:self.used + C8 + C]
    var.fill_(1.0)
    self.used += 2 * C8
    return mean, var

  def tensor(self):
    return self.buf[:self.used]


class BatchNorm(object):
  """Training-mode batch norm over NHWC bf16 with statistics fused into the producer conv."""

  def __init__(self, store, name, C, eps=1e-5, momentum=0.1, zero_gamma=False):
    self.C, self.eps, self.momentum, self.name = C, eps, momentum, name
    self.sg = store.register(name + ".gamma", (C,), False, constant(0.0 if zero_gamma else 1.0))
    self.sb = store.register(name + ".beta", (C,), False, constant(0.0))
    self.store = store

  def build(self, device, arena=None, running=None):
    """``arena`` (StatsArena): take the (sum, sumsq) accumulators from a per-trainer arena that
    the trainer zeroes once per step; the statistics are then finalised inside the apply kernel
    (one launch less per batch norm and no clearing pass).  ``running`` (RunningArena): place
    the running statistics in the trainer's flat non-trainable-state buffer."""
    z = lambda: torch.zeros(self.C, dtype=torch.float32, device=device)  # noqa: E731
    self.arena = arena
    if arena is not None:
      self.sum, self.sumsq = arena.take(self.C), arena.take(self.C)
    else:
      self.sum, self.sumsq = z(), z()
    self.mean, self.invstd, self.scale, self.shift = z(), z(), z(), z()
    # accumulators of the backward reduction when it is fused into the epilogue of the data
    # gradient that produces this unit's dy (enable_fused_reduce): zeroed with the same memset
    self.red_g = self.red_gx = None
    self.fused_reduce = False
    if arena is not None:
      self.red_g, self.red_gx = arena.take(self.C), arena.take(self.C)
    if running is not None:
      self.running_mean, self.running_var = running.take(self.C)
    else:
      self.running_mean = z()
      self.running_var = torch.ones(self.C, dtype=torch.float32, device=device)
    self.gamma, self.beta = self.store.f32(self.sg), self.store.f32(self.sb)
    self.dgamma, self.dbeta = self.store.g(self.sg), self.store.g(self.sb)

  @property
  def stats(self):
    return (self.sum, self.sumsq)

  def forward(self, x_raw, y, residual=None, act=1, training=True, fused_stats=True, rev=False):
    """``rev``: walk the rows last-to-first.  The convolution that produced ``x_raw`` wrote its
    tiles first-to-last, so the END of the tensor is what the 126 MB L2 still holds; the apply
    kernel then finishes at the START of ``y``, which is where the next convolution begins."""
    count = x_raw.numel() // self.C
    ops.C().bn_set_row_reverse(bool(rev))
    # a residual unit's ReLU mask depends on the sum, not on x alone: keep it as one bit per
    # element so that the backward pass does not have to re-read the whole bf16 output
    # (the buffer is never released: a captured CUDA graph may hold its address)
    if training and act == 1 and (residual is not None or getattr(self, "fused_reduce", False)):
      # (a unit whose backward reduction is fused into a dgrad epilogue always keeps the bits:
      # the epilogue and the apply kernel then read 1/16 of a byte per element for the mask)
      self.mask = self.ensure_mask(x_raw.numel(), x_raw.device)
    else:
      self.mask = None
    if training and fused_stats and getattr(self, "arena", None) is not None:
      ops.K.bn_apply_finalize(x_raw, residual, y, act, self.mask, self.sum, self.sumsq, self.gamma,
                              self.beta, self.running_mean, self.running_var, self.mean,
                              self.invstd, self.scale, self.shift, float(count), self.eps,
                              self.momentum)
      return
    if training:
      if not fused_stats:
        ops.K.bn_stats(x_raw, self.sum, self.sumsq)
      ops.K.bn_finalize(self.sum, self.sumsq, self.gamma, self.beta, self.running_mean,
                        self.running_var, self.mean, self.invstd, self.scale, self.shift,
                        float(count), self.eps, self.momentum)
    else:
      ops.K.bn_inference_coeffs(self.gamma, self.beta, self.running_mean, self.running_var,
                                self.scale, self.shift, self.eps)
    ops.K.bn_apply(x_raw, residual, self.scale, self.shift, y, act, self.mask)

  def ensure_mask(self, numel, device):
    """The unit's ReLU bit mask buffer (numel / 8 bytes), allocated once."""
    buf = getattr(self, "_mask_buf", None)
    if buf is None or buf.numel() * 8 != numel:
      buf = self._mask_buf = torch.empty(numel // 8, dtype=torch.uint8, device=device)
    return buf

  def backward(self, dy, x_raw, y, dx, dres=None, relu=True, residual=False, mask=None,
               rev=(False, False)):
    """relu mask: recomputed from x_raw with the forward scale/shift when the unit has no
    residual input (saves reading y: 2 of 6-8 bytes per element); from the stored y otherwise;
    ``mask``: another unit's bit mask (a projection shortcut's BN sees the gradient of the block
    output, i.e. the ReLU that followed the *sum*)."""
    mode = 0 if not relu else (1 if (residual or dres is not None) else 2)
    ysrc = y if mode == 1 else None
    if mask is not None and relu:
      mode, ysrc = 3, mask
    elif relu and getattr(self, "mask", None) is not None:
      mode, ysrc = 3, self.mask   # bit mask written by forward(): 1/16 of y's bytes
    if getattr(self, "fused_reduce", False):
      # sum(g), sum(g x) were accumulated by the epilogue of the data gradient that wrote dy
      # (ops/igemm.conv_dgrad(bn_reduce=...)): no reduction pass, the apply kernel finishes them
      ops.C().bn_set_row_reverse(bool(rev[1]))
      ops.K.bn_bwd_apply(dy, x_raw, ysrc, self.gamma, self.mean, self.invstd, self.dgamma,
                         self.dbeta, dx, dres, mode, self.scale, self.shift, self.red_g,
                         self.red_gx)
      ops.C().bn_set_row_reverse(False)
      return
    # rev = (reduce, apply) row order: each kernel starts at the end where the previous one
    # stopped, so that what it reads first is still in L2 (ResNetTrainer assigns the directions)
    ops.C().bn_set_row_reverse(bool(rev[0]))
    ops.K.bn_bwd_reduce(dy, x_raw, ysrc, self.mean, self.invstd, self.dgamma, self.dbeta, mode,
                        self.scale, self.shift)
    ops.C().bn_set_row_reverse(bool(rev[1]))
    ops.K.bn_bwd_apply(dy, x_raw, ysrc, self.gamma, self.mean, self.invstd, self.dgamma,
                       self.dbeta, dx, dres, mode, self.scale, self.shift)
    ops.C().bn_set_row_reverse(False)


class Conv(object):
  """Convolution parameters + its three igemm plans (fprop / dgrad / wgrad)."""

  def __init__(self, store, name, cin, cout, k, stride=1, pad=None, bias=False):
    self.cin, self.cout, self.k, self.stride = cin, cout, k, stride
    self.pad = (k // 2) if pad is None else pad
    self.name = name
    self.store = store
    self.sw = store.register(name + ".w", (cout, k, k, cin), True, he_normal(k * k * cin))
    self.sbias = store.register(name + ".b", (cout,), False, constant(0.0)) if bias else None

  def out_hw(self, H, W):
    return ((H + 2 * self.pad - self.k) // self.stride + 1,
            (W + 2 * self.pad - self.k) // self.stride + 1)

  def build(self, x, y, dy=None, dx=None, stats=None, relu=False, dx_accumulate=False,
            need_dgrad=True, training=True, dx_acc_mask=None, dx_bn_reduce=None):
    """``dx_bn_reduce``: the BatchNorm whose dy this layer's data gradient produces (its input
    has the shape of ``dx``): fuse that batch norm's backward reduction into the dgrad epilogue
    when the plan allows it (every pixel of dx written by this launch)."""
    st = self.store
    self.x, self.y = x, y
    if not training:
      stats = None  # inference: running statistics are used, nothing to accumulate
    bias = st.f32(self.sbias) if self.sbias is not None else None
    self.fwd = igemm.conv_fprop(x, st.w(self.sw), y, self.stride, self.pad, bias=bias, relu=relu,
                                stats=stats)
    self._fwd_args = (x, y, bias, relu, stats)
    self.fwd_remote = None
    self.wgrad = self.dgrad = None
    if training and dy is not None:
      self.wgrad = igemm.conv_wgrad(dy, x, st.g(self.sw), self.stride, self.pad)
      if need_dgrad and dx is not None:
        red = None
        if dx_bn_reduce is not None and dx_bn_reduce[0].red_g is not None:
          bn_below, x_below, mask_below = dx_bn_reduce
          red = (x_below, mask_below, bn_below.red_g, bn_below.red_gx)
        try:
          self.dgrad = igemm.conv_dgrad(dy, st.w(self.sw), dx, self.stride, self.pad,
                                        accumulate=dx_accumulate, acc_mask=dx_acc_mask,
                                        bn_reduce=red)
          if red is not None:
            dx_bn_reduce[0].fused_reduce = True
        except (ValueError, RuntimeError) as e:
          if red is None:
            raise
          # (a strided 1x1 does not write every pixel, odd channel counts leave ragged tiles)
          self.dgrad = igemm.conv_dgrad(dy, st.w(self.sw), dx, self.stride, self.pad,
                                        accumulate=dx_accumulate, acc_mask=dx_acc_mask)
    self.dy = dy

  def forward(self):
    self.fwd.run()

  def bind_remote_weights(self, flat_weights):
    """Second forward plan whose B operand (the filter) is fetched by TMA from ``flat_weights`` -
    a peer-mapped view of ANOTHER rank's bf16 weight buffer - instead of local memory: the
    startup broadcast fused with the first convolution that consumes the weights."""
    x, y, bias, relu, stats = self._fwd_args
    w = self.store._view(flat_weights, self.sw)
    self.fwd_remote = igemm.conv_fprop(x, w, y, self.stride, self.pad, bias=bias, relu=relu,
                                       stats=stats)

  def forward_remote(self):
    self.fwd_remote.run()

  def backward(self):
    _run_wgrad(self)
    if self.dgrad is not None:
      self.dgrad.run()


class Dense(object):
  """Fully connected layer on [B, K] activations (GEMM + bias (+ReLU) epilogue)."""

  def __init__(self, store, name, cin, cout, bias=True, init=None):
    self.cin, self.cout, self.name, self.store = cin, cout, name, store
    self.sw = store.register(name + ".w", (cout, cin), True, init or he_normal(cin))
    self.sbias = store.register(name + ".b", (cout,), False, constant(0.0)) if bias else None

  def build(self, x, y, dy=None, dx=None, relu=False, training=True):
    st = self.store
    bias = st.f32(self.sbias) if self.sbias is not None else None
    self.fwd = igemm.gemm(x, st.w(self.sw), y, "nk", bias=bias, relu=relu)
    self._fwd_args = (x, y, bias, relu)
    self.fwd_remote = None
    self.wgrad = self.dgrad = None
    self.dy = dy
    if training and dy is not None:
      self.wgrad = igemm.gemm_wgrad(dy, x, st.g(self.sw))
      if dx is not None:
        self.dgrad = igemm.gemm(dy, st.w(self.sw), dx, "kn")

  def forward(self):
    self.fwd.run()

  def bind_remote_weights(self, flat_weights):
    x, y, bias, relu = self._fwd_args
    self.fwd_remote = igemm.gemm(x, self.store._view(flat_weights, self.sw), y, "nk", bias=bias,
                                 relu=relu)

  def forward_remote(self):
    self.fwd_remote.run()

  def backward(self):
    _run_wgrad(self)
    if self.dgrad is not None:
      self.dgrad.run()
