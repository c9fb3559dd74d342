"""Two-level (several hosts) all-reduce + optimizer, checked on the GPUs of ONE box
(run under torchrun, one rank per GPU):

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/gpu_check_hier.py

A. "one host x W GPUs" through the hierarchical code path: the fused kernel split into its PHASE 1
   (NVLink reduce-scatter into the rank's own shard) and PHASE 2 (update + all-gather) halves;
B. "W hosts x one GPU": NCCL all-reduce between the hosts + the PHASE 2 half on a world of one;
C. with 4+ GPUs, "W/2 hosts x 2 GPUs": both levels at once (sub-group symmetric memory, one NCCL
   group per local index).
Each against torch.distributed.all_reduce + the same momentum update in plain PyTorch, two steps
(the second one exercises the sharded momentum state), then FusedOptimizer.assemble().
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def rel(a, b):
  return float((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-6))


def main():
  rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
  local = int(os.environ.get("LOCAL_RANK", rank))
  torch.cuda.set_device(local)
  dev = torch.device("cuda", local)
  dist.init_process_group("nccl", device_id=dev)
  from tensorflowonspark_b200.models import engine
  from tensorflowonspark_b200.parallel import group_comm, symm
  from tensorflowonspark_b200.parallel.fused_optim import FusedOptimizer
  ok = True

  def report(name, err, tol):
    nonlocal ok
    good = err == err and err < tol
    ok &= good
    if rank == 0 or not good:
      print("CHECK[r{}] {:58s} err={:.3e} tol={:.1e} {}".format(rank, name, err, tol,
                                                                "OK" if good else "FAIL"), flush=True)

  def run(tag, comm):
    st = engine.ParamStore()
    st.register("w1", (1024, 1024), True, engine.normal(0.1))
    st.register("w2", (512, 2048), True, engine.normal(0.1))
    st.register("gamma", (4096,), False, engine.constant(1.0))
    st.finalize(dev, alloc=comm.alloc, seed=7)
    n = st.total
    cut = (n // 3) // 8 * 8
    opt = FusedOptimizer(st, comm=comm, opt="momentum", lr=0.1, momentum=0.9, weight_decay=1e-2,
                         buckets=[(0, cut, "a"), (cut, n, "b")])
    assert opt.hier_mode and opt.gworld == world
    w = st.master.clone().double()
    m = torch.zeros_like(w)
    decay = (torch.arange(n, device=dev) < st.decay_end).double()
    for step in range(2):
      torch.manual_seed(1000 * step + rank)
      g_local = torch.randn(n, device=dev)
      opt.zero_grads()
      st.grads.copy_(g_local)
      torch.cuda.synchronize()
      dist.barrier()
      opt.finish()                      # bucketed: launch() on the comm stream, then the join
      torch.cuda.synchronize()
      g = g_local.clone()
      dist.all_reduce(g)
      g = g.double() / world + 1e-2 * w * decay
      m = 0.9 * m + g
      w = w - 0.1 * m
      report("{} step {}: bf16 weights (every shard)".format(tag, step), rel(st.weights, w), 1e-2)
      report("{} step {}: fp32 aux replica".format(tag, step), rel(st.aux32[:n - st.decay_end], w[st.decay_end:]), 1e-5)
    dist.barrier()
    opt.assemble()                      # pull the local peers' master / momentum shards
    torch.cuda.synchronize()
    report("{}: assembled fp32 master".format(tag), rel(st.master, w), 1e-5)
    report("{}: assembled momentum".format(tag), rel(opt.state1, m), 1e-5)
    dist.barrier()

  # A: one host, all GPUs local - PHASE 1 + PHASE 2 kernels, no network step
  sym = symm.from_torch_distributed(dev)
  run("A 1 host x {} GPUs (phase 1 + phase 2)".format(world), group_comm.HierComm(sym, None, rank, world))
  # B: every GPU its own host - NCCL between them, PHASE 2 on a world of one
  run("B {} hosts x 1 GPU (NCCL + phase 2)".format(world),
      group_comm.HierComm(None, dist.group.WORLD, rank, world, device=dev))
  # C: two GPUs per host
  if world >= 4 and world % 2 == 0:
    host, idx = rank // 2, rank % 2
    pairs = [dist.new_group([2 * h, 2 * h + 1]) for h in range(world // 2)]
    inters = [dist.new_group(list(range(k, world, 2))) for k in range(2)]

    def exchange(obj):
      out = [None, None]
      dist.all_gather_object(out, obj, group=pairs[host])
      return out
    run("C {} hosts x 2 GPUs (phase 1, NCCL, phase 2)".format(world // 2),
        group_comm.HierComm(symm.SymmComm(idx, 2, exchange, dev), inters[idx], rank, world))
  t = torch.tensor([1.0 if ok else 0.0], device=dev)
  dist.all_reduce(t, op=dist.ReduceOp.MIN)
  if rank == 0:
    print("HIER CHECK", "PASSED" if float(t) == 1.0 else "FAILED", flush=True)
  dist.destroy_process_group()
  sys.exit(0 if float(t) == 1.0 else 1)


if __name__ == "__main__":
  main()
