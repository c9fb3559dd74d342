"""Cross-host gradient path (parallel/group_comm.py + FusedOptimizer group mode).

The reference hands TensorFlow a multi-host cluster spec (TFSparkNode.py:373-384) and lets
MultiWorkerMirroredStrategy cross the network; here ranks that do not share a host fall back from
peer-mapped memory to a torch.distributed all-reduce followed by the fused optimizer kernel in
its single-rank form.  On CPU the kernel is replaced by a twin with the same update rule
(csrc/optim_comm.cu opt_update), so what is tested is the orchestration: bucket all-reduce,
1 / world scaling, replicated state, identical replicas, and the deferred update around a
captured step."""
import multiprocessing
import os
import socket

import pytest
import torch

from tensorflowonspark_b200 import TFSparkNode


def _free_port():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  port = s.getsockname()[1]
  s.close()
  return port


class _KernelTwin(object):
  """allreduce_opt for world == 1 on host tensors: momentum SGD with L2 on the decayed head."""

  def __init__(self, store, optim):
    self.store, self.optim, self.calls = store, optim, 0

  def allreduce_opt(self, d):
    st, o = self.store, self.optim
    assert d["world"] == 1 and d["grads"] == [st.grads.data_ptr()]
    assert d.get("phase", 0) == (2 if o.hier_mode else 0)   # one-GPU hosts: no reduce-scatter half
    self.calls += 1
    h = o.hyper
    b, e = d["begin"], d["end"]
    g = st.grads[b:e] * h[3]
    w = st.master[b:e]
    decay = (torch.arange(b, e) < d["decay_end"]).float()
    g = g + h[2] * w * decay
    o.state1[b:e] = h[1] * o.state1[b:e] + g
    st.master[b:e] = w - h[0] * o.state1[b:e]
    st.weights[b:e] = st.master[b:e].to(torch.bfloat16)
    lo = max(b, st.decay_end)
    if e > lo:
      st.aux32[lo - st.decay_end:e - st.decay_end] = st.master[lo:e]


def _rank_main(rank, world, port, q, mode="group"):
  try:
    import torch.distributed as dist
    from tensorflowonspark_b200 import ops
    from tensorflowonspark_b200.models import engine
    from tensorflowonspark_b200.parallel import group_comm
    from tensorflowonspark_b200.parallel.fused_optim import FusedOptimizer
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{}".format(port), rank=rank,
                            world_size=world)
    if mode == "group":
      comm = group_comm.GroupComm(device="cpu")
    else:       # two "hosts" with one worker each: inter-host all-reduce + the PHASE 2 half only
      comm = group_comm.HierComm(None, dist.group.WORLD, rank, world, device="cpu")
      assert (comm.hosts, comm.local_world, comm.local_rank) == (2, 1, 0)
    st = engine.ParamStore()
    st.register("w1", (16, 8), True, engine.normal(0.1))
    st.register("w2", (8, 8), True, engine.normal(0.1))
    st.register("gamma", (8,), False, engine.constant(1.0))
    st.finalize(torch.device("cpu"), alloc=comm.alloc, seed=7 + rank)   # ranks start DIFFERENT
    for name in ("master", "weights", "aux32"):
      comm.broadcast(name, root=0)                                      # ... the chief's values win
    n = st.total
    opt = FusedOptimizer(st, comm=comm, opt="momentum", lr=0.1, momentum=0.9, weight_decay=1e-2,
                         buckets=[(0, 128, "a"), (128, n, "b")])
    twin = _KernelTwin(st, opt)
    ops.K.allreduce_opt = twin.allreduce_opt
    assert (opt.group_mode or opt.hier_mode) and opt.world == 1 and opt.gworld == world
    assert abs(float(opt.hyper[3]) - 1.0 / world) < 1e-7
    w0 = st.master.clone()
    gens = [torch.Generator().manual_seed(100 + r) for r in range(world)]
    ref_w, ref_m = w0.clone(), torch.zeros(n)
    decay = (torch.arange(n) < st.decay_end).float()
    for step in range(3):
      grads = [torch.randn(n, generator=g) for g in gens]               # every rank knows them all
      opt.zero_grads()
      st.grads.copy_(grads[rank])
      if step == 1:
        # a captured step: nothing may run inside the "capture", the update follows the replay
        opt._capturing_group = lambda: True
        opt.finish()
        assert opt.deferred and twin.calls == 2 * step
        opt._capturing_group = lambda: False
        opt.after_replay()
      else:
        opt.finish()
      g = sum(grads) / world + 1e-2 * ref_w * decay
      ref_m = 0.9 * ref_m + g
      ref_w = ref_w - 0.1 * ref_m
      assert twin.calls == 2 * (step + 1)
    err = float((st.master - ref_w).abs().max())
    everyone = [torch.zeros(n) for _ in range(world)]
    dist.all_gather(everyone, st.master)
    same = all(torch.equal(everyone[0], t) for t in everyone)
    opt.assemble()                                                      # no-op: state is replicated
    sd = opt.state_dict()
    if mode == "group":
      with pytest.raises(RuntimeError):
        comm.peer_ptrs("weights")
    comm.barrier()
    q.put((rank, err, same, float((sd["state1"] - ref_m).abs().max())))
    dist.destroy_process_group()
  except Exception:
    import traceback
    q.put((rank, traceback.format_exc(), False, None))


@pytest.mark.parametrize("mode", ["group", "hier"])
def test_group_mode_allreduce_update_keeps_replicas_identical(mode):
  world, port = 2, _free_port()
  mp = multiprocessing.get_context("spawn")
  q = mp.Queue()
  procs = [mp.Process(target=_rank_main, args=(r, world, port, q, mode)) for r in range(world)]
  for p in procs:
    p.start()
  out = [q.get(timeout=120) for _ in procs]
  for p in procs:
    p.join(30)
  for rank, err, same, merr in out:
    assert not isinstance(err, str), err
    assert err < 1e-5 and merr < 1e-5 and same, (rank, err, same, merr)


def _ctx(spec, job, index):
  ctx = TFSparkNode.TFNodeContext(executor_id=index, job_name=job, task_index=index, cluster_spec=spec)
  os.environ.pop("RANK", None)
  TFSparkNode._export_dist_env(ctx, [])
  return ctx


def test_ctx_reports_hosts_and_refuses_symmetric_memory_across_hosts(monkeypatch):
  for k in ("MASTER_ADDR", "MASTER_PORT", "WORLD_SIZE", "RANK", "LOCAL_RANK", "TF_CONFIG", "TFOS_CONFIG"):
    monkeypatch.setenv(k, os.environ.get(k, ""))     # restored afterwards
  two = {"chief": ["10.0.0.1:4000"], "worker": ["10.0.0.1:4001", "10.0.0.2:4000", "10.0.0.2:4001"]}
  ctx = _ctx(two, "worker", 1)
  assert ctx.rank == 2 and ctx.local_rank == 0
  assert ctx.worker_hosts() == ["10.0.0.1", "10.0.0.1", "10.0.0.2", "10.0.0.2"]
  assert not ctx.single_host
  with pytest.raises(RuntimeError, match="one host"):
    ctx.symmetric_comm()
  with pytest.raises(RuntimeError, match="one host"):
    ctx.symmetric_comm(ranks=[1, 2])
  one = {"chief": ["10.0.0.1:4000"], "worker": ["10.0.0.1:4001"]}
  assert _ctx(one, "chief", 0).single_host


def test_hier_layout_groups_ranks_by_host_and_local_index():
  from tensorflowonspark_b200.parallel.process_group import hier_layout
  assert hier_layout(["a", "a", "b", "b"], 2) == ([2, 3], 0, [[0, 2], [1, 3]])
  assert hier_layout(["a", "b", "a", "b"], 3) == ([1, 3], 1, [[0, 1], [2, 3]])    # interleaved hosts
  assert hier_layout(["a", "a", "b"], 0) is None                                 # uneven: flat fallback
  assert hier_layout(["a", "a"], 1) == ([0, 1], 1, [[0], [1]])
