"""Multi-GPU checks of the fused collectives (run under torchrun, one rank per GPU):

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/gpu_check_multi.py

* symmetric memory: peer writes become visible after the device-side flag barrier;
* fused all-reduce + {sgd, momentum, adam}: parity with torch.distributed.all_reduce followed by
  the same update in plain PyTorch;
* broadcast: every rank ends up with the root's buffer;
* bandwidth of the fused kernel's peer traffic against the measured NVLink peer-copy figure.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def rel(a, b):
  return float((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-6))


def main():
  os.environ.setdefault("TFOS_NVLS", "1")   # exercise the multicast path at any world size
  rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
  local = int(os.environ.get("LOCAL_RANK", rank))
  torch.cuda.set_device(local)
  dev = torch.device("cuda", local)
  dist.init_process_group("nccl", device_id=dev)
  from tensorflowonspark_b200 import ops
  from tensorflowonspark_b200.parallel import symm
  comm = symm.from_torch_distributed(dev)
  ok = True

  def report(name, err, tol):
    nonlocal ok
    good = err == err and err < tol
    ok &= good
    if rank == 0 or not good:
      print("CHECK[r{}] {:40s} err={:.3e} tol={:.1e} {}".format(rank, name, err, tol,
                                                                "OK" if good else "FAIL"), flush=True)

  # ---- broadcast + barrier
  n = 1 << 20
  buf = comm.alloc("bc", n, torch.float32)
  buf.fill_(float(rank + 1))
  torch.cuda.synchronize()
  comm.broadcast("bc", root=0)
  torch.cuda.synchronize()
  report("bcast_pull", float((buf - 1.0).abs().max()), 1e-9)
  for it in range(20):  # barrier stress: peers must observe each other's writes every round
    buf.fill_(float(it * world + rank))
    comm.barrier()
    peer = ops.C().tensor_from_ptr(comm.peer_ptrs("bc")[(rank + 1) % world], [n], "f32")
    want = float(it * world + (rank + 1) % world)
    bad = float((peer[:1024] - want).abs().max())
    comm.barrier()
    if bad != 0.0:
      report("flag_barrier visibility round {}".format(it), bad, 1e-9)
      break
  else:
    report("flag_barrier visibility x20", 0.0, 1e-9)

  # ---- fused all-reduce + optimizer
  N = (25_557_032 + 7) // 8 * 8 if os.environ.get("TFOS_FULL", "1") == "1" else 1 << 20
  decay_end = N // 2 // 8 * 8
  if rank == 0:
    print("NVLS multicast substrate:", "available (cuMem VMM + cuMulticast)" if comm.nvls
          else "not available on this box / disabled: peer-to-peer path only", flush=True)
  variants = [(0, "sgd", False), (1, "momentum", False), (2, "adam", False)]
  if comm.nvls:   # the same kernels through multimem.ld_reduce / multimem.st
    variants += [(1, "momentum", True), (2, "adam", True)]
  for vi, (opt, name, nvls) in enumerate(variants):
    if nvls:
      name += "[nvls]"
    torch.manual_seed(1234)
    w0 = torch.randn(N, device=dev)
    torch.manual_seed(100 + rank)
    g_local = torch.randn(N, device=dev)
    grads = comm.alloc("g%d" % vi, N, torch.float32, multicast=nvls)
    weights = comm.alloc("w%d" % vi, N, torch.bfloat16, multicast=nvls)
    aux = comm.alloc("a%d" % vi, N - decay_end, torch.float32, multicast=nvls)
    grads.copy_(g_local)
    master = w0.clone()
    # Adam is checked in a well-conditioned state (v = 1, step 10): on the very first step the
    # update is lr * g / |g|, so an element whose 8 gradients nearly cancel flips sign with the
    # summation order and no two correct implementations agree
    s1, s2 = torch.zeros(N, device=dev), torch.full((N,), 1.0 if opt == 2 else 0.0, device=dev)
    hyper = torch.tensor([0.1, 0.9, 1e-2, 1.0 / world, 0.9, 0.999, 1e-7, 10.0 if opt == 2 else 1.0],
                         device=dev)
    d = {"master": master.data_ptr(), "state1": s1.data_ptr(), "state2": s2.data_ptr(),
         "hyper": hyper.data_ptr(), "begin": 0, "end": N, "decay_end": decay_end, "world": world,
         "rank": rank, "slot": vi, "opt": opt, "grid": 64,
         "grads": comm.peer_ptrs("g%d" % vi), "weights": comm.peer_ptrs("w%d" % vi),
         "aux32": comm.peer_ptrs("a%d" % vi), "aux_begin": decay_end, "flags": comm.flag_ptrs(),
         "epoch": comm.epoch_ptr(vi), "block_counter": comm.counter_ptr(vi)}
    if nvls:
      d.update(grads_mc=comm.mc_ptr("g%d" % vi), weights_mc=comm.mc_ptr("w%d" % vi),
               aux32_mc=comm.mc_ptr("a%d" % vi))
      assert d["grads_mc"] and d["weights_mc"] and d["aux32_mc"]
    torch.cuda.synchronize()
    dist.barrier()
    ops.K.allreduce_opt(d)
    torch.cuda.synchronize()
    # reference: NCCL all-reduce + the same update in PyTorch
    g = g_local.clone()
    dist.all_reduce(g)
    g /= world
    g[:decay_end] += 1e-2 * w0[:decay_end]
    if opt == 2:
      m, v = 0.1 * g, 0.999 + 0.001 * g * g
      c1, c2 = 1.0 - 0.9 ** 10, 1.0 - 0.999 ** 10
      ref = w0 - 0.1 * (m / c1) / (torch.sqrt(v / c2) + 1e-7)
    else:
      ref = w0 - 0.1 * g
    report("allreduce_{} bf16 weights (all shards)".format(name), rel(weights, ref), 1e-2)
    report("allreduce_{} fp32 aux replica".format(name), rel(aux, ref[decay_end:]),
           1e-4 if opt == 2 else 1e-5)
    chunk = ((N + world - 1) // world + 7) // 8 * 8
    lo, hi = min(N, chunk * rank), min(N, chunk * (rank + 1))
    # Adam divides by sqrt(v): the order in which 8 peers' gradients are summed shows up at a
    # few 1e-4 on elements whose gradient nearly cancels (plus __powf under --use_fast_math)
    report("allreduce_{} master shard".format(name), rel(master[lo:hi], ref[lo:hi]),
           1e-4 if opt == 2 else 1e-5)   # (__powf under --use_fast_math in the bias correction)
    # timing (momentum only): device-timed, max over ranks
    if opt == 1:
      grads.copy_(g_local)
      torch.cuda.synchronize()
      dist.barrier()
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      iters = 10
      e0.record()
      for _ in range(iters):
        ops.K.allreduce_opt(d)
      e1.record()
      torch.cuda.synchronize()
      t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
      ms = float(t)
      link_bytes = (world - 1) / world * N * 4 + (world - 1) / world * N * 2  # pulled grads + pushed weights
      if rank == 0:
        print("TIMING allreduce_{} N={} world={} {:.3f} ms  {:.1f} GB/s per GPU over NVLink "
              "({:.0f}% of the 770 GB/s measured peer-copy rate)".format(
                  name, N, world, ms, link_bytes / ms / 1e6, 100 * link_bytes / ms / 1e6 / 770))
      # NCCL baseline: all-reduce + unfused update
      gg = g_local.clone()
      torch.cuda.synchronize()
      dist.barrier()
      e0.record()
      for _ in range(iters):
        dist.all_reduce(gg)
        s1.mul_(0.9).add_(gg, alpha=1.0 / world)
        master.add_(s1, alpha=-0.1)
        weights.copy_(master)
      e1.record()
      torch.cuda.synchronize()
      t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
      if rank == 0:
        print("TIMING nccl all_reduce + unfused momentum update (baseline) {:.3f} ms".format(float(t)))
  # ---- startup broadcast fused with the first forward pass (filters read over NVLink by TMA)
  from tensorflowonspark_b200.models import resnet
  comm2 = symm.from_torch_distributed(dev)
  net = resnet.ResNetTrainer(depth=50, batch=8, image=64, device=dev, comm=comm2, lr=0.0,
                             weight_decay=0.0)
  if rank == 0:  # only the root holds the "real" variables
    net.store.weights.mul_(1.25)
    net.store.master.mul_(1.25)
  else:
    net.store.weights.zero_()   # a rank that consumed its own copy would produce all-zero logits
  x, y = net.synthetic_batch(seed=7)  # same batch on every rank
  net.set_input(x, y)
  net.bind_broadcast_root(0)
  torch.cuda.synchronize()
  dist.barrier()
  net.first_step_fused_broadcast()
  torch.cuda.synchronize()
  mine = net.logits.clone()
  ref_logits = mine.clone()
  dist.broadcast(ref_logits, 0)
  # not bit-equal: batch-norm statistics are accumulated with atomics (order varies per run)
  report("fused bcast+fwd: logits match root's", rel(mine, ref_logits), 2e-2)
  report("fused bcast+fwd: logits are non-trivial", 1.0 / (float(mine.abs().max()) + 1e-9), 1e3)
  w_ref = net.store.weights.float().clone()
  dist.broadcast(w_ref, 0)
  report("fused bcast: local weight copy", rel(net.store.weights, w_ref), 1e-6)
  m_ref = net.store.master.clone()
  dist.broadcast(m_ref, 0)
  report("fused bcast: fp32 masters", rel(net.store.master, m_ref), 1e-6)
  # ---- sharded optimizer state: checkpoint taken by ONE rank must hold every rank's shards
  comm3 = symm.from_torch_distributed(dev)
  tr = resnet.ResNetTrainer(depth=50, batch=8, image=64, device=dev, comm=comm3, lr=0.05,
                            weight_decay=1e-4)
  comm3.broadcast("weights", root=0)
  comm3.broadcast("aux32", root=0)
  comm3.broadcast("master", root=0, slot=60)
  x, y = tr.synthetic_batch(seed=11 + rank)
  for _ in range(3):
    tr.train_step(x, y)
  torch.cuda.synchronize()
  dist.barrier()
  sd = tr.store.state_dict() if rank == 0 else None       # chief only, like the examples
  osd = tr.optim.state_dict() if rank == 0 else None
  torch.cuda.synchronize()
  dist.barrier()
  if rank == 0:
    full = torch.cat([sd[sp["name"]].reshape(-1) for sp in tr.store.order]).to(dev)
    wts = torch.cat([tr.store.w(sp).reshape(-1).float() for sp in tr.store.order])
    # every parameter moved away from its initial value and the saved fp32 master rounds to the
    # bf16 weights that all ranks compute with - on EVERY shard, not only the chief's
    report("checkpoint: assembled master == live bf16 weights", rel(full.bfloat16(), wts), 1e-6)
    mom = osd["state1"].to(dev)
    for r in range(world):
      lo, hi = tr.optim.shard_bounds(0, r)
      report("checkpoint: momentum of rank {}'s shard is non-zero".format(r),
             1.0 / (float(mom[lo:hi].abs().max()) + 1e-12), 1e6)
    # resume: a fresh trainer loaded from the checkpoint computes with identical weights
    fresh = resnet.ResNetTrainer(depth=50, batch=8, image=64, device=dev, lr=0.05, seed=99)
    fresh.store.load_state_dict(sd)
    w2 = torch.cat([fresh.store.w(sp).reshape(-1).float() for sp in fresh.store.order])
    report("checkpoint: save -> load -> weights equal", rel(w2, wts), 1e-6)
  dist.barrier()
  flag = torch.tensor([1 if ok else 0], device=dev)
  dist.all_reduce(flag, op=dist.ReduceOp.MIN)   # every rank must have passed every check
  ok = bool(int(flag))
  if rank == 0:
    print("MULTI SUMMARY:", "ALL OK" if ok else "FAILURES")
  dist.destroy_process_group()
  sys.exit(0 if ok else 1)


if __name__ == "__main__":
  main()
