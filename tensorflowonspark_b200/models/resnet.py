"""ResNet (v1.5 bottleneck: 50/101/152; basic-block CIFAR variants 20/32/56) on the
native engine: NHWC bf16 activations, tcgen05 implicit-GEMM convolutions with
batch-norm statistics fused into their epilogues, fused BN+ReLU(+residual)
apply/backward kernels, fused all-reduce + momentum-SGD optimizer.

Workload parity: the reference trains ResNet-56/CIFAR-10 through an external
Keras model (examples/resnet/resnet_cifar_dist.py:208, momentum SGD with a
piecewise schedule :35-66); BASELINE.json names ResNet-50/ImageNet-shaped data
as the headline config.  Both are provided here.
"""
import os

import torch

from .. import ops
from ..ops import igemm
from . import engine
from .engine import BatchNorm, Conv, Dense, ParamStore, normal


class _Unit(object):
  """conv -> BN (-> +residual) -> ReLU with its forward and gradient buffers."""

  def __init__(self, store, name, cin, cout, k, stride, zero_gamma=False):
    self.conv = Conv(store, name + ".conv", cin, cout, k, stride)
    self.bn = BatchNorm(store, name + ".bn", cout, zero_gamma=zero_gamma)
    self.cout = cout


class _Block(object):
  pass


def _buf(shape, device, dtype=torch.bfloat16):
  return torch.zeros(shape, dtype=dtype, device=device)


class ResNetTrainer(object):
  """Static-buffer ResNet.  ``train_step(images_u8 | normalized, labels)`` runs
  forward, backward and the fused optimizer; ``forward_only`` serves inference."""

  CFG = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3), 152: (3, 8, 36, 3)}

  def __init__(self, depth=50, batch=256, image=224, num_classes=1000, device="cuda:0",
               lr=0.1, momentum=0.9, weight_decay=1e-4, comm=None, training=True, seed=1234,
               optimizer="momentum"):
    assert depth in self.CFG, "bottleneck depths: 50/101/152"
    self.device = torch.device(device)
    self.B, self.image, self.num_classes, self.depth = batch, image, num_classes, depth
    self.training = training
    self.comm = comm
    dev = self.device
    st = self.store = ParamStore()
    B = batch

    # ---- parameters -----------------------------------------------------
    self.stem_w = st.register("stem.conv.w", (64, 7, 64), True, self._stem_init)
    self.stem_bn = BatchNorm(st, "stem.bn", 64)
    self.blocks = []
    cin = 64
    for si, nblocks in enumerate(self.CFG[depth]):
      width = 64 * (2 ** si)
      for bi in range(nblocks):
        stride = 2 if (bi == 0 and si > 0) else 1
        b = _Block()
        name = "layer{}.{}".format(si + 1, bi)
        b.u1 = _Unit(st, name + ".u1", cin, width, 1, 1)
        b.u2 = _Unit(st, name + ".u2", width, width, 3, stride)
        b.u3 = _Unit(st, name + ".u3", width, width * 4, 1, 1, zero_gamma=True)
        b.ds = _Unit(st, name + ".ds", cin, width * 4, 1, stride) if bi == 0 else None
        b.stride, b.cin, b.width, b.name = stride, cin, width, name
        self.blocks.append(b)
        cin = width * 4
    self.fc = Dense(st, "fc", cin, num_classes, bias=True, init=normal(0.01))
    self.feat = cin

    alloc = comm.alloc if comm is not None else None
    st.finalize(dev, alloc=alloc, seed=seed)

    # ---- buffers and plans -----------------------------------------------
    OH, OW, Wp = igemm.stem_geometry(image, image)
    self.in_u8 = _buf((B, image, image, 3), dev, torch.uint8)
    self.xp = _buf((B, image, Wp, 8), dev)
    self.labels = torch.zeros(B, dtype=torch.int32, device=dev)
    self.stem_raw = _buf((B, OH, OW, 64), dev)
    self.stem_act = None   # allocated below unless the fused stem tail makes it unnecessary
    PH, PW = (OH + 2 - 3) // 2 + 1, (OW + 2 - 3) // 2 + 1
    self.pool = _buf((B, PH, PW, 64), dev)
    self.pool_idx = _buf((B, PH, PW, 64), dev, torch.uint8)
    tr = training
    # all batch-norm statistics accumulators live in one arena that is zeroed once per step;
    # finalisation is folded into the apply kernels (engine.BatchNorm.build)
    self.stats_arena = engine.StatsArena(dev) if (tr and dev.type == "cuda" and os.environ.get(
        "TFOS_BN_FUSED_FINALIZE", "1") != "0") else None
    self.running = engine.RunningArena(dev)   # all BN running statistics, one flat buffer
    self.stem_bn.build(dev, self.stats_arena, self.running)
    self.fuse_stem = (tr and self.stats_arena is not None and dev.type == "cuda"
                      and os.environ.get("TFOS_FUSE_STEM", "1") != "0")
    if not self.fuse_stem:
      self.stem_act = _buf((B, OH, OW, 64), dev)
    self.p_stem = igemm.stem_fprop(self.xp, st.w(self.stem_w), self.stem_raw,
                                   stats=self.stem_bn.stats if tr else None)
    if tr:
      self.g_stem_act = None if self.fuse_stem else _buf((B, OH, OW, 64), dev)
      self.g_stem_raw = _buf((B, OH, OW, 64), dev)
      self.p_stem_wgrad = igemm.stem_wgrad(self.g_stem_raw, self.xp, st.g(self.stem_w))

    x, H, W = self.pool, PH, PW
    prev_block = None
    gx = _buf(x.shape, dev) if tr else None  # gradient wrt the block input
    self.g_pool = gx
    for b in self.blocks:
      s = b.stride
      H2, W2 = (H - 1) // s + 1, (W - 1) // s + 1
      w, c4 = b.width, b.width * 4
      b.x = x
      b.r1, b.a1 = _buf((B, H, W, w), dev), _buf((B, H, W, w), dev)
      b.r2, b.a2 = _buf((B, H2, W2, w), dev), _buf((B, H2, W2, w), dev)
      b.r3, b.out = _buf((B, H2, W2, c4), dev), _buf((B, H2, W2, c4), dev)
      if tr:
        b.g_r1, b.g_a1 = _buf(b.r1.shape, dev), _buf(b.a1.shape, dev)
        b.g_r2, b.g_a2 = _buf(b.r2.shape, dev), _buf(b.a2.shape, dev)
        b.g_r3 = _buf(b.r3.shape, dev)
        # identity blocks share the stage gradient buffer: g_out is rewritten in place
        # into g_x (masked residual gradient + conv1 dgrad accumulated on top)
        b.g_out = _buf(b.out.shape, dev) if b.ds is not None else gx
        b.g_x = gx
      else:
        b.g_r1 = b.g_a1 = b.g_r2 = b.g_a2 = b.g_r3 = b.g_out = b.g_x = None
      for u in (b.u1, b.u2, b.u3) + ((b.ds,) if b.ds else ()):
        u.bn.build(dev, self.stats_arena, self.running)
      ident = b.ds is None
      # the block output's ReLU mask (one bit per element, written by bn3's forward): identity
      # blocks fold "mask * dY" into conv1's accumulating dgrad instead of materialising it
      omask = b.u3.bn.ensure_mask(b.out.numel(), dev) if tr else None
      # Fused batch-norm backward reductions (TFOS_BN_FUSED_REDUCE=0 disables): the data gradient
      # that WRITES a batch norm's dy also accumulates that batch norm's sum(g), sum(g x):
      #   conv3.dgrad -> g_a2 = dy of bn2        conv2.dgrad -> g_a1 = dy of bn1
      #   identity block: conv1.dgrad (accumulating) completes g_x = dy of the PREVIOUS block's bn3
      fr = tr and self.stats_arena is not None and os.environ.get("TFOS_BN_FUSED_REDUCE", "1") != "0"
      red1 = (b.u1.bn, b.r1, b.u1.bn.ensure_mask(b.a1.numel(), dev)) if fr else None
      red2 = (b.u2.bn, b.r2, b.u2.bn.ensure_mask(b.a2.numel(), dev)) if fr else None
      redp = None
      if fr and ident and prev_block is not None:
        redp = (prev_block.u3.bn, prev_block.r3, prev_block.u3.bn.ensure_mask(
            prev_block.out.numel(), dev))
      b.u1.conv.build(x, b.r1, b.g_r1, b.g_x, stats=b.u1.bn.stats, dx_accumulate=ident,
                      training=tr, dx_acc_mask=omask if ident else None, dx_bn_reduce=redp)
      b.u2.conv.build(b.a1, b.r2, b.g_r2, b.g_a1, stats=b.u2.bn.stats, training=tr,
                      dx_bn_reduce=red1)
      b.u3.conv.build(b.a2, b.r3, b.g_r3, b.g_a2, stats=b.u3.bn.stats, training=tr,
                      dx_bn_reduce=red2)
      if b.ds is not None:
        b.rd, b.idn = _buf(b.r3.shape, dev), _buf(b.r3.shape, dev)
        b.g_rd = _buf(b.r3.shape, dev) if tr else None
        b.ds.conv.build(x, b.rd, b.g_rd, b.g_x, stats=b.ds.bn.stats, dx_accumulate=True,
                        training=tr)
      prev_block = b
      x, H, W = b.out, H2, W2
      if tr:
        gx = b.g_out
    self.last = x
    self.g_last = gx
    self.avg = _buf((B, self.feat), dev)
    self.logits = _buf((B, num_classes), dev, torch.float32)
    self.ldd = (num_classes + 7) // 8 * 8
    self.dlogits = _buf((B, self.ldd), dev) if tr else None
    self.g_avg = _buf((B, self.feat), dev) if tr else None
    self.loss_sum = torch.zeros(1, dtype=torch.float32, device=dev)
    self.correct = torch.zeros(1, dtype=torch.float32, device=dev)
    dl = self.dlogits[:, :num_classes] if (tr and self.ldd == num_classes) else None
    if tr and dl is None:
      raise ValueError("num_classes must be a multiple of 8 (pad the head)")
    self.fc.build(self.avg, self.logits, dl, self.g_avg, training=tr)

    self.optim = None
    if tr:
      from ..parallel.fused_optim import FusedOptimizer
      buckets = self._comm_buckets() if (comm is not None and comm.world > 1) else None
      self.optim = FusedOptimizer(st, comm=comm, opt=optimizer, lr=lr, momentum=momentum,
                                  weight_decay=weight_decay, buckets=buckets)
    # opt-in (TFOS_WGRAD_STREAM=1): weight gradients on a side stream, concurrent with the BN
    # kernels (engine.py).  Measured on B200: 20.13 vs 20.17 ms/step - the two kernel families do
    # co-reside on the SMs but share the same L2 -> SM fabric ceiling, so nothing is gained.
    self.wg_stream = None
    if tr and dev.type == "cuda" and os.environ.get("TFOS_WGRAD_STREAM", "0") == "1":
      self.wg_stream = torch.cuda.Stream(device=dev)
      self.optim.add_producer_stream(self.wg_stream)
    self.graph = None
    self.mean = [0.485, 0.456, 0.406]
    self.std = [0.229, 0.224, 0.225]
    self._assign_directions()

  def _assign_directions(self):
    """L2 hand-over ("serpentine") schedule, TFOS_L2_SERPENTINE=0 disables it.

    Activations are 13 - 411 MB per tensor at batch 256, the L2 holds 126 MB: a kernel that walks
    a tensor front to back leaves its TAIL in L2.  Every memory-bound kernel of the chain
    therefore starts at the end where its predecessor stopped:
      forward   conv (tiles ascending) -> BN apply (rows descending) -> next conv (ascending) ...
      backward  BN reduce (d) -> BN apply (not d) -> the unit's dgrad (d) -> next unit (not d) ...
    Only the ORDER in which tiles / rows are visited changes; results are identical."""
    self.serpentine = self.device.type == "cuda" and os.environ.get("TFOS_L2_SERPENTINE", "1") == "1"
    d = False
    for b in reversed(self.blocks):
      for u in (b.u3, b.u2, b.u1):
        u.rev = (d, not d) if self.serpentine else (False, False)
        if self.serpentine and u.conv.dgrad is not None:
          u.conv.dgrad.set_reverse(d)
        d = not d
      if b.ds is not None:
        b.ds.rev = (False, True) if self.serpentine else (False, False)
    self.stem_rev = (d, not d) if self.serpentine else (False, False)

  def _comm_buckets(self):
    """Gradient buckets for overlapping the fused all-reduce with backward: (begin, end, tag)
    where tag is the index of the block after whose backward the range is final ('stem' = end
    of backward).  layer4+fc hold 2/3 of the parameters and are final a third into backward;
    what is left for the un-overlappable tail after the stem's weight gradient is only stem +
    layer1 (0.23 M of 25.6 M parameters) and the batch-norm parameters of stem..layer2."""
    st = self.store

    def first_offset(prefix, decay=True):
      return min(s["offset"] for s in st.order if s["decay"] == decay and s["name"].startswith(prefix))

    def first_block(prefix):
      return next(i for i, b in enumerate(self.blocks) if b.name.startswith(prefix))

    b2, b3, b4 = first_offset("layer2."), first_offset("layer3."), first_offset("layer4.")
    nd3 = first_offset("layer3.", decay=False)   # non-decayed tail: BN scale/offset in layer order
    return [(b4, st.decay_end, first_block("layer4.")), (b3, b4, first_block("layer3.")),
            (nd3, st.total, first_block("layer3.")), (b2, b3, first_block("layer2.")),
            (0, b2, "stem"), (st.decay_end, nd3, "stem")]

  @staticmethod
  def _stem_init(shape, gen):
    import math
    w = torch.randn((shape[0], 7, 7, 3), generator=gen) * math.sqrt(2.0 / (7 * 7 * 3))
    return igemm.pack_stem_weight(w)

  # ------------------------------------------------------------------ data
  def synthetic_batch(self, seed=0):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    x = torch.randint(0, 256, (self.B, self.image, self.image, 3), dtype=torch.uint8, generator=g)
    y = torch.randint(0, self.num_classes, (self.B,), dtype=torch.int32, generator=g)
    return x.to(self.device), y.to(self.device)

  def set_input(self, images_u8, labels=None):
    """Stage a uint8 NHWC batch (device or pinned host) into the static input buffers."""
    self.in_u8.copy_(images_u8, non_blocking=True)
    if labels is not None:
      self.labels.copy_(labels, non_blocking=True)

  # ------------------------------------------------- broadcast fused with use
  def bind_broadcast_root(self, root=0):
    """Prepare the fused startup broadcast: forward plans whose filters are read by TMA
    straight out of the ROOT rank's weight buffer over NVLink (tile by tile, as the tcgen05
    main loop consumes them), while `bcast_pull` fills the local copy on a side stream."""
    comm = self.comm
    flat = ops.C().tensor_from_ptr(comm.peer_ptrs("weights")[root], [self.store.total], "bf16")
    self._root_weights = flat
    self._root = root
    w_stem = self.store._view(flat, self.stem_w)
    self.p_stem_remote = igemm.stem_fprop(self.xp, w_stem, self.stem_raw, stats=self.stem_bn.stats)
    for b in self.blocks:
      for u in (b.u1, b.u2, b.u3) + ((b.ds,) if b.ds is not None else ()):
        u.conv.bind_remote_weights(flat)
    self.fc.bind_remote_weights(flat)
    self._side = torch.cuda.Stream(device=self.device)

  def first_step_fused_broadcast(self):
    """First training step after start-up: the root's variables win (reference semantics of
    entering strategy.scope(), examples/mnist/keras/mnist_spark.py:55-56) without a separate
    broadcast phase in front of the first forward pass."""
    comm = self.comm
    comm.broadcast("aux32", self._root)  # BN scale/offset + biases: tiny, consumed in fp32
    comm.barrier()                        # root's weights are final and visible
    main = torch.cuda.current_stream(self.device)
    self._side.wait_stream(main)
    with torch.cuda.stream(self._side):
      comm.broadcast("weights", self._root)     # local copy for the backward pass / later steps
      comm.broadcast("master", self._root, slot=60)  # fp32 masters restart from the root's values
    self._forward(True, remote=True)            # forward consumes the root's tiles over NVLink
    main.wait_stream(self._side)
    self._loss(True)
    self._backward()
    self.optim.finish()
    return self.loss_sum

  # --------------------------------------------------------------- forward
  def _forward(self, training, remote=False):
    K = ops.K
    run = (lambda c: c.forward_remote()) if remote else (lambda c: c.forward())
    if self.stats_arena is not None:
      self.stats_arena.zero()   # the fused conv epilogues accumulate from zero every pass
    K.decode_normalize(self.in_u8, self.xp, igemm.STEM_PAD, self.mean, self.std)
    (self.p_stem_remote if remote else self.p_stem).run()
    rv = self.serpentine
    if training and self.fuse_stem:
      # BN + ReLU + max pool in one pass over the raw conv output (csrc/stem_fused.cu): the
      # 411 MB stem activation is never written
      bn = self.stem_bn
      K.stem_bn_relu_pool_fwd(self.stem_raw, self.pool, self.pool_idx, bn.sum, bn.sumsq, bn.gamma,
                              bn.beta, bn.running_mean, bn.running_var, bn.mean, bn.invstd,
                              bn.scale, bn.shift, float(self.stem_raw.numel() // bn.C), bn.eps,
                              bn.momentum)
    else:
      self.stem_bn.forward(self.stem_raw, self.stem_act, None, 1, training)
      K.maxpool_fwd(self.stem_act, self.pool, self.pool_idx, 3, 2, 1)
    for b in self.blocks:
      for u, raw, act in ((b.u1, b.r1, b.a1), (b.u2, b.r2, b.a2)):
        run(u.conv)
        u.bn.forward(raw, act, None, 1, training, rev=rv)
      run(b.u3.conv)
      if b.ds is not None:
        run(b.ds.conv)
        b.ds.bn.forward(b.rd, b.idn, None, 0, training, rev=rv)
        idn = b.idn
      else:
        idn = b.x
      b.u3.bn.forward(b.r3, b.out, idn, 1, training, rev=rv)
    K.avgpool_fwd(self.last, self.avg)
    run(self.fc)

  def _loss(self, with_grad):
    self.loss_sum.zero_()
    self.correct.zero_()
    ops.K.softmax_xent(self.logits, self.labels, self.dlogits if with_grad else None,
                       self.loss_sum, self.correct, self.num_classes, 1.0 / self.B)

  # -------------------------------------------------------------- backward
  def _backward(self):
    K = ops.K
    self.optim.zero_grads()
    engine.set_wgrad_stream(self.wg_stream)
    try:
      self._backward_layers(K)
    finally:
      engine.set_wgrad_stream(None)

  def _backward_layers(self, K):
    self.fc.backward()
    K.avgpool_bwd(self.g_avg, self.g_last)
    for bi in reversed(range(len(self.blocks))):
      b = self.blocks[bi]
      # out = relu(bn3(r3) + idn): the masked gradient goes to both branches; the mask is applied
      # where g_out is consumed (here, in conv1's accumulate epilogue, in the shortcut's BN)
      b.u3.bn.backward(b.g_out, b.r3, b.out, b.g_r3, relu=True, residual=True, rev=b.u3.rev)
      b.u3.conv.backward()
      b.u2.bn.backward(b.g_a2, b.r2, b.a2, b.g_r2, relu=True, rev=b.u2.rev)
      b.u2.conv.backward()
      b.u1.bn.backward(b.g_a1, b.r1, b.a1, b.g_r1, relu=True, rev=b.u1.rev)
      b.u1.conv.backward()  # writes (ds) or accumulates (identity) into g_x
      if b.ds is not None:
        b.ds.bn.backward(b.g_out, b.rd, None, b.g_rd, relu=True, mask=b.u3.bn.mask, rev=b.ds.rev)
        b.ds.conv.backward()  # accumulates into g_x
      self.optim.launch(bi)   # buckets that became final start their all-reduce now
    if self.fuse_stem:
      # max-pool backward recomputed inside the BN reduction and the BN apply: the gradient of
      # the stem activation (411 MB) is never written or read
      bn = self.stem_bn
      K.stem_pool_bn_bwd(self.g_pool, self.pool_idx, self.stem_raw, bn.gamma, bn.mean, bn.invstd,
                         bn.scale, bn.shift, bn.dgamma, bn.dbeta, self.g_stem_raw)
      ops.count()   # two launches
    else:
      K.maxpool_bwd(self.g_pool, self.pool_idx, self.g_stem_act, 3, 2, 1)
      self.stem_bn.backward(self.g_stem_act, self.stem_raw, self.stem_act, self.g_stem_raw,
                            relu=True, rev=self.stem_rev)
    self.p_stem_wgrad.run()

  # ------------------------------------------------------------------ API
  def step_kernels(self):
    """Forward + loss + backward + optimizer on the static buffers (capturable)."""
    self._forward(True)
    self._loss(True)
    self._backward()
    self.optim.finish()

  def capture(self):
    """Capture one training step into a CUDA graph (launch-bound inner loop -> one launch)."""
    s = torch.cuda.Stream(device=self.device)
    s.wait_stream(torch.cuda.current_stream(self.device))
    with torch.cuda.stream(s):
      for _ in range(2):
        self.step_kernels()
    torch.cuda.current_stream(self.device).wait_stream(s)
    torch.cuda.synchronize(self.device)
    g = torch.cuda.CUDAGraph()
    # cross-host communicator: its collectives stay outside the graph (FusedOptimizer.finish()
    # defers them), but the process group's watchdog thread may still poll events of the warm-up
    # collectives - thread-local capture keeps that from invalidating the capture
    mode = "thread_local" if getattr(self.optim, "group_mode", False) else "global"
    with torch.cuda.graph(g, capture_error_mode=mode):
      self.step_kernels()
    self.graph = g
    return g

  def train_step(self, images_u8=None, labels=None):
    if images_u8 is not None:
      self.set_input(images_u8, labels)
    if self.graph is not None:
      self.graph.replay()
      self.optim.after_replay()   # cross-host communicator: all-reduce + update outside the graph
    else:
      self.step_kernels()
    return self.loss_sum  # device scalar: mean loss of the step

  def forward_only(self, images_u8=None):
    if images_u8 is not None:
      self.set_input(images_u8)
    self._forward(False)
    return self.logits

  # ------------------------------------------ inference with folded batch norm
  def build_folded_inference(self):
    """Serving path: every batch norm is folded into the convolution in front of it
    (w' = w * gamma / sqrt(var + eps) per output channel, bias = beta - mean * that scale, from
    the fp32 masters and the running statistics), ReLU and the residual add run in the conv
    epilogues (bias / ReLU / accumulate), identity blocks accumulate IN PLACE into their input.
    The forward pass is then 53 tcgen05 convolutions + 2 pools + the dense head: none of the 53
    batch-norm passes (2/3 of the inference-time HBM traffic) remains.  Call again after the
    parameters or running statistics change (``fold_batch_norms``)."""
    assert not self.training, "folded inference is built on a training=False trainer"
    st, dev = self.store, self.device
    units = [("stem", self.stem_w, self.stem_bn)]
    for b in self.blocks:
      for u in (b.u1, b.u2, b.u3) + ((b.ds,) if b.ds is not None else ()):
        units.append((u, u.conv.sw, u.bn))
    self._fold_units = units
    self.wf = torch.zeros_like(st.weights)        # folded bf16 filters, same layout as weights
    self.bf = torch.zeros(sum((bn.C + 7) // 8 * 8 for _, _, bn in units), dtype=torch.float32,
                          device=dev)
    bias, off = {}, 0
    for key, _, bn in units:
      bias[id(bn)] = self.bf[off:off + bn.C]
      off += (bn.C + 7) // 8 * 8
    wv = lambda spec: st._view(self.wf, spec)  # noqa: E731
    plans = [igemm.stem_fprop(self.xp, wv(self.stem_w), self.stem_act, bias=bias[id(self.stem_bn)],
                              relu=True)]
    cur = self.pool
    for b in self.blocks:
      s_ = b.stride
      plans.append(igemm.conv_fprop(cur, wv(b.u1.conv.sw), b.a1, 1, 0, bias=bias[id(b.u1.bn)],
                                    relu=True))
      plans.append(igemm.conv_fprop(b.a1, wv(b.u2.conv.sw), b.a2, s_, 1, bias=bias[id(b.u2.bn)],
                                    relu=True))
      if b.ds is not None:
        plans.append(igemm.conv_fprop(cur, wv(b.ds.conv.sw), b.out, s_, 0,
                                      bias=bias[id(b.ds.bn)]))
        target = b.out
      else:
        target = cur          # out = relu(conv3 + x): accumulate into x itself
      plans.append(igemm.conv_fprop(b.a2, wv(b.u3.conv.sw), target, 1, 0, bias=bias[id(b.u3.bn)],
                                    relu=True, accumulate=True))
      cur = target
    self._f_plans, self._f_last = plans, cur
    self.fold_batch_norms()

  def fold_batch_norms(self):
    st = self.store
    off = 0
    for key, spec, bn in self._fold_units:
      scale = st.m(bn.sg) * torch.rsqrt(bn.running_var + bn.eps)
      w = st.m(spec).float()
      st._view(self.wf, spec).copy_(w * scale.view(-1, *([1] * (w.dim() - 1))))
      self.bf[off:off + bn.C].copy_(st.m(bn.sb) - bn.running_mean * scale)
      off += (bn.C + 7) // 8 * 8

  def forward_folded(self):
    """logits of the staged batch through the folded network (capturable)."""
    K = ops.K
    K.decode_normalize(self.in_u8, self.xp, igemm.STEM_PAD, self.mean, self.std)
    self._f_plans[0].run()
    K.maxpool_fwd(self.stem_act, self.pool, self.pool_idx, 3, 2, 1)
    for p in self._f_plans[1:]:
      p.run()
    K.avgpool_fwd(self._f_last, self.avg)
    self.fc.forward()
    return self.logits

  def set_lr(self, lr):
    self.optim.set_lr(lr)

  # ------------------------------------------------- checkpoints / serving
  export_builder = "tensorflowonspark_b200.models.resnet:build_served"
  export_signatures = {"serving_default": {
      "inputs": {"image": "image"}, "outputs": {"logits": "logits", "prediction": "prediction"},
      "input_dtypes": {"image": "uint8"}}}

  @property
  def export_builder_args(self):
    return {"depth": self.depth, "image": self.image, "num_classes": self.num_classes}

  def state_dict(self):
    """Parameters (fp32 masters, assembled across ranks when sharded) + the batch-norm running
    statistics: everything inference or a resumed run needs."""
    sd = dict(self.store.state_dict())
    sd["__running__"] = self.running.tensor().detach().cpu().clone()
    return sd

  def load_state_dict(self, sd):
    sd = dict(sd)
    running = sd.pop("__running__", None)
    self.store.load_state_dict(sd)
    if running is not None:
      self.running.tensor().copy_(running.to(self.device))


class ServedResNet(object):
  """Inference callable behind ``pipeline.TFModel`` / ``TFParallel``: one bf16 replica on the
  executor's GPU.  ``submit(inputs)`` stages a uint8 NHWC batch in page-locked memory, copies it
  on the copy stream and enqueues the forward pass + an async read-back of the results;
  ``collect()`` returns the oldest submitted batch - so the H2D copy of batch i+1 and the host's
  row handling overlap the kernels of batch i (pipeline._run_model drives it one batch ahead)."""

  def __init__(self, state, depth=50, image=224, num_classes=1000, batch=256, device="cuda:0"):
    from ..feed import DevicePrefetcher
    torch.cuda.set_device(torch.device(device))
    self.net = ResNetTrainer(depth=depth, batch=batch, image=image, num_classes=num_classes,
                             device=device, training=False)
    self.net.load_state_dict(state)
    self.folded = os.environ.get("TFOS_SERVE_FOLD_BN", "1") == "1"
    if self.folded:
      self.net.build_folded_inference()
    self._fwd = self.net.forward_folded if self.folded else (lambda: self.net._forward(False))
    self.B, self.image, self.V = batch, image, num_classes
    dev = self.net.device
    self.feeder = DevicePrefetcher([((batch, image, image, 3), torch.uint8)], dev, depth=2)
    self.h_logits = [torch.empty(batch, num_classes, dtype=torch.float32).pin_memory()
                     for _ in range(2)]
    self.h_pred = [torch.empty(batch, dtype=torch.int64).pin_memory() for _ in range(2)]
    self.done = [torch.cuda.Event() for _ in range(2)]
    self._pending = []
    self._k = 0
    self.pred_dev = torch.zeros(batch, dtype=torch.int64, device=dev)
    self.graph, self._batches = None, 0
    self.use_graph = os.environ.get("TFOS_SERVE_GRAPH", "1") == "1"
    self.copy_threads = max(1, int(os.environ.get("TFOS_SERVE_COPY_THREADS", "4")))
    if self.copy_threads > 1:
      from concurrent.futures import ThreadPoolExecutor
      self._pool = ThreadPoolExecutor(max_workers=self.copy_threads, thread_name_prefix="tfos-stage")

  def submit(self, inputs):
    import numpy as np
    x = inputs["image"] if isinstance(inputs, dict) else inputs
    x = np.asarray(x)
    n = len(x)
    if n > self.B:
      raise ValueError("batch of {} rows exceeds the served batch size {}".format(n, self.B))
    x = x.reshape(n, self.image, self.image, 3)
    if n < self.B:   # ragged tail: pad (the padded rows' outputs are dropped)
      x = np.concatenate([x, np.zeros((self.B - n,) + x.shape[1:], dtype=x.dtype)])
    self.feeder.push_arrays([x])
    self._enqueue(n)

  def submit_rows(self, columns):
    """Same as :meth:`submit` for a column of raw row cells (bytes / memoryview / ndarray, one
    uint8 HWC image each): every row is copied ONCE, straight into the page-locked staging
    buffer the H2D copy reads - no intermediate batch array (pipeline._run_model uses this when a
    served model offers it)."""
    import numpy as np
    rows = columns["image"] if isinstance(columns, dict) else columns
    n = len(rows)
    if n > self.B:
      raise ValueError("batch of {} rows exceeds the served batch size {}".format(n, self.B))
    (host,) = self.feeder.acquire_host()
    hv = host.numpy().reshape(self.B, -1)

    def fill(lo, hi):
      for r in range(lo, hi):
        row = rows[r]
        hv[r] = np.frombuffer(row, dtype=np.uint8) if not isinstance(row, np.ndarray) \
            else row.reshape(-1)

    # 38.5 MB per 256-image batch: one thread copies ~5 GB/s, which is slower than the GPU
    # consumes it (8.5 vs 6.5 ms per batch); numpy releases the GIL during the copies, so a few
    # threads fill disjoint row ranges of the staging buffer in parallel
    k = self.copy_threads
    if k > 1 and n >= 4 * k:
      step = (n + k - 1) // k
      list(self._pool.map(lambda i: fill(i * step, min(n, (i + 1) * step)), range(k)))
    else:
      fill(0, n)
    if n < self.B:
      hv[n:] = 0
    self.feeder.push_host()
    self._enqueue(n)

  def _capture(self):
    """The whole forward pass (~165 launches) + arg-max as ONE CUDA graph: the host thread that
    also fills the staging buffers spends microseconds, not a millisecond, per batch on launches."""
    dev = self.net.device
    cur = torch.cuda.current_stream(dev)
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(cur)
    with torch.cuda.stream(side):
      self._fwd()
      self.pred_dev.copy_(self.net.logits[:, :self.V].argmax(1))
    cur.wait_stream(side)
    torch.cuda.synchronize(dev)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
      self._fwd()
      self.pred_dev.copy_(self.net.logits[:, :self.V].argmax(1))
    self.graph = g

  def _enqueue(self, n):
    (dx,) = self.feeder.pop()
    self.net.set_input(dx)
    self.feeder.release()
    if self.graph is None and self.use_graph and self._batches >= 1:
      self._capture()
    if self.graph is not None:
      self.graph.replay()
    else:
      self._fwd()
      self.pred_dev.copy_(self.net.logits[:, :self.V].argmax(1))
    self._batches += 1
    logits = self.net.logits
    k = self._k
    self._k ^= 1
    self.h_logits[k].copy_(logits[:, :self.V], non_blocking=True)
    self.h_pred[k].copy_(self.pred_dev, non_blocking=True)
    self.done[k].record(torch.cuda.current_stream(self.net.device))
    self._pending.append((k, n))

  def collect(self):
    k, n = self._pending.pop(0)
    self.done[k].synchronize()
    return {"logits": self.h_logits[k][:n].numpy().copy(),
            "prediction": self.h_pred[k][:n].numpy().copy()}

  def pending(self):
    return len(self._pending)

  def __call__(self, **inputs):
    self.submit(inputs)
    return self.collect()


def build_served(state, depth=50, image=224, num_classes=1000, batch=None, **_):
  """Export builder (utils/checkpoint.py): rebuilds the replica on the executor's GPU."""
  batch = int(batch or os.environ.get("TFOS_SERVE_BATCH", "256"))
  return ServedResNet(state, depth=depth, image=image, num_classes=num_classes, batch=batch)


class CifarResNetTrainer(object):
  """ResNet-20/32/56 for 32x32 inputs (basic blocks, 16/32/64 channels) - the network the
  reference example actually trains (examples/resnet/resnet_cifar_dist.py:208, ``resnet56``,
  momentum SGD, batch 128, weight decay 2e-4)."""

  def __init__(self, depth=56, batch=128, num_classes=10, device="cuda:0", lr=0.1, momentum=0.9,
               weight_decay=2e-4, comm=None, seed=1234):
    assert (depth - 2) % 6 == 0
    n = (depth - 2) // 6
    self.device = dev = torch.device(device)
    self.B, self.image, self.num_classes = batch, 32, num_classes
    self.VP = (num_classes + 7) // 8 * 8
    B = batch
    st = self.store = ParamStore()
    self.stem = _Unit(st, "stem", 8, 16, 3, 1)
    self.blocks = []
    cin = 16
    for si, width in enumerate((16, 32, 64)):
      for bi in range(n):
        stride = 2 if (bi == 0 and si > 0) else 1
        b = _Block()
        name = "stage{}.{}".format(si + 1, bi)
        b.u1 = _Unit(st, name + ".u1", cin, width, 3, stride)
        b.u2 = _Unit(st, name + ".u2", width, width, 3, 1, zero_gamma=True)
        b.ds = _Unit(st, name + ".ds", cin, width, 1, stride) if stride != 1 else None
        b.stride, b.width = stride, width
        self.blocks.append(b)
        cin = width
    self.fc = Dense(st, "fc", 64, self.VP, bias=True, init=normal(0.01))
    st.finalize(dev, alloc=comm.alloc if comm is not None else None, seed=seed)

    self.in_u8 = _buf((B, 32, 32, 3), dev, torch.uint8)
    self.x8 = _buf((B, 32, 32, 8), dev)
    self.labels = torch.zeros(B, dtype=torch.int32, device=dev)
    self.stem_raw, self.stem_act = _buf((B, 32, 32, 16), dev), _buf((B, 32, 32, 16), dev)
    self.g_stem_raw = _buf((B, 32, 32, 16), dev)
    self.running = engine.RunningArena(dev)
    self.stem.bn.build(dev, None, self.running)
    gx = _buf((B, 32, 32, 16), dev)
    self.g_stem_act = gx
    self.stem.conv.build(self.x8, self.stem_raw, self.g_stem_raw, None, stats=self.stem.bn.stats,
                         need_dgrad=False)
    x, H = self.stem_act, 32
    for b in self.blocks:
      H2, w = (H - 1) // b.stride + 1, b.width
      b.x = x
      b.r1, b.a1 = _buf((B, H2, H2, w), dev), _buf((B, H2, H2, w), dev)
      b.r2, b.out = _buf((B, H2, H2, w), dev), _buf((B, H2, H2, w), dev)
      b.g_r1, b.g_a1, b.g_r2 = _buf(b.r1.shape, dev), _buf(b.a1.shape, dev), _buf(b.r2.shape, dev)
      b.g_out = _buf(b.out.shape, dev) if b.ds is not None else gx
      b.g_x = gx
      for u in (b.u1, b.u2) + ((b.ds,) if b.ds else ()):
        u.bn.build(dev, None, self.running)
      ident = b.ds is None
      b.u1.conv.build(x, b.r1, b.g_r1, b.g_x, stats=b.u1.bn.stats, dx_accumulate=ident)
      b.u2.conv.build(b.a1, b.r2, b.g_r2, b.g_a1, stats=b.u2.bn.stats)
      if b.ds is not None:
        b.rd, b.idn, b.g_rd = _buf(b.r2.shape, dev), _buf(b.r2.shape, dev), _buf(b.r2.shape, dev)
        b.ds.conv.build(x, b.rd, b.g_rd, b.g_x, stats=b.ds.bn.stats, dx_accumulate=True)
      x, H, gx = b.out, H2, b.g_out
    self.last, self.g_last = x, gx
    self.avg, self.g_avg = _buf((B, 64), dev), _buf((B, 64), dev)
    self.logits = _buf((B, self.VP), dev, torch.float32)
    self.dlogits = _buf((B, self.VP), dev)
    self.loss_sum = torch.zeros(1, dtype=torch.float32, device=dev)
    self.correct = torch.zeros(1, dtype=torch.float32, device=dev)
    self.fc.build(self.avg, self.logits, self.dlogits, self.g_avg)
    from ..parallel.fused_optim import FusedOptimizer
    self.optim = FusedOptimizer(st, comm=comm, opt="momentum", lr=lr, momentum=momentum,
                                weight_decay=weight_decay)
    self.graph = None
    self.mean, self.std = [0.4914, 0.4822, 0.4465], [0.2470, 0.2435, 0.2616]

  synthetic_batch = ResNetTrainer.synthetic_batch
  state_dict = ResNetTrainer.state_dict
  load_state_dict = ResNetTrainer.load_state_dict
  set_input = ResNetTrainer.set_input
  capture = ResNetTrainer.capture
  train_step = ResNetTrainer.train_step
  set_lr = ResNetTrainer.set_lr

  def step_kernels(self):
    K = ops.K
    K.decode_normalize(self.in_u8, self.x8, 0, self.mean, self.std)
    self.stem.conv.forward()
    self.stem.bn.forward(self.stem_raw, self.stem_act, None, 1, True)
    for b in self.blocks:
      b.u1.conv.forward()
      b.u1.bn.forward(b.r1, b.a1, None, 1, True)
      b.u2.conv.forward()
      if b.ds is not None:
        b.ds.conv.forward()
        b.ds.bn.forward(b.rd, b.idn, None, 0, True)
      b.u2.bn.forward(b.r2, b.out, b.idn if b.ds is not None else b.x, 1, True)
    K.avgpool_fwd(self.last, self.avg)
    self.fc.forward()
    self.loss_sum.zero_()
    self.correct.zero_()
    K.softmax_xent(self.logits, self.labels, self.dlogits, self.loss_sum, self.correct,
                   self.num_classes, 1.0 / self.B)
    self.optim.zero_grads()
    self.fc.backward()
    K.avgpool_bwd(self.g_avg, self.g_last)
    for b in reversed(self.blocks):
      b.u2.bn.backward(b.g_out, b.r2, b.out, b.g_r2, dres=b.g_out, relu=True)
      b.u2.conv.backward()
      b.u1.bn.backward(b.g_a1, b.r1, b.a1, b.g_r1, relu=True)
      b.u1.conv.backward()
      if b.ds is not None:
        b.ds.bn.backward(b.g_out, b.rd, None, b.g_rd, relu=False)
        b.ds.conv.backward()
    self.stem.bn.backward(self.g_stem_act, self.stem_raw, self.stem_act, self.g_stem_raw, relu=True)
    self.stem.conv.backward()
    self.optim.step()


def piecewise_lr(epoch, batch_size, base=0.1, boundaries=(91, 136, 182), factors=(0.1, 0.01, 0.001)):
  """LR schedule of the reference ResNet example (resnet_cifar_dist.py:35-66):
  0.1 * bs/128, multiplied by 0.1 / 0.01 / 0.001 from epochs 91 / 136 / 182."""
  lr = base * batch_size / 128.0
  for b, f in zip(boundaries, factors):
    if epoch >= b:
      lr = base * batch_size / 128.0 * f
  return lr
