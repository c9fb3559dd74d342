"""Cluster life-cycle on 2 executor processes (scenarios of reference tests/test_TFCluster.py:16-121
plus ps / evaluator / epochs / terminate paths the reference never tested)."""
import time

import pytest

from tensorflowonspark_b200 import TFCluster, TFNode


def test_independent_nodes(sc):
  def fn(args, ctx):
    assert args["x"] + args["y"] == 3

  cluster = TFCluster.run(sc, fn, {"x": 1, "y": 2}, 2, 0)
  cluster.shutdown()


def test_inputmode_spark_inference_roundtrip(sc):
  def fn(args, ctx):
    import numpy as np
    feed = TFNode.DataFeed(ctx.mgr, False)
    while not feed.should_stop():
      batch = feed.next_batch(10)
      if len(batch) > 0:
        feed.batch_results(np.square(np.array(batch)).tolist())

  rdd = sc.parallelize([[x] for x in range(1000)], 10)
  cluster = TFCluster.run(sc, fn, {}, 2, 0, input_mode=TFCluster.InputMode.SPARK)
  out = cluster.inference(rdd)
  assert out.map(lambda r: r[0]).sum() == sum(x * x for x in range(1000))
  cluster.shutdown()


def test_train_epochs_and_terminate(sc):
  def fn(args, ctx):
    feed = ctx.get_data_feed(train_mode=True)
    seen = 0
    while not feed.should_stop() and seen < 300:
      seen += len(feed.next_batch(50))
    with open(args["out"] + str(ctx.executor_id), "w") as f:
      f.write(str(seen))
    feed.terminate()

  import tempfile
  out = tempfile.mkdtemp() + "/seen"
  cluster = TFCluster.run(sc, fn, {"out": out}, 2, 0, input_mode=TFCluster.InputMode.SPARK)
  cluster.train(sc.parallelize(range(200), 4), num_epochs=5)  # 1000 rows offered, 600 consumed
  cluster.shutdown()
  assert [int(open(out + str(i)).read()) for i in range(2)] == [300, 300]


def test_exception_during_feed_reaches_driver(sc):
  def fn(args, ctx):
    feed = TFNode.DataFeed(ctx.mgr, False)
    while not feed.should_stop():
      batch = feed.next_batch(10)
      if len(batch) > 0:
        feed.batch_results(batch)
        raise Exception("FAKE exception during feeding")

  rdd = sc.parallelize([[x] for x in range(1000)], 10)
  with pytest.raises(Exception, match="FAKE exception during feeding|Timeout"):
    cluster = TFCluster.run(sc, fn, {}, 2, 0, input_mode=TFCluster.InputMode.SPARK)
    cluster.inference(rdd, feed_timeout=2).count()
  try:
    cluster.shutdown()
  except Exception:
    pass


def test_late_exception_caught_by_shutdown(sc):
  def fn(args, ctx):
    feed = TFNode.DataFeed(ctx.mgr, False)
    while not feed.should_stop():
      batch = feed.next_batch(10)
      if len(batch) > 0:
        feed.batch_results(batch)
    time.sleep(1)
    raise Exception("FAKE exception after feeding")

  rdd = sc.parallelize([[x] for x in range(100)], 4)
  cluster = TFCluster.run(sc, fn, {}, 2, 0, input_mode=TFCluster.InputMode.SPARK)
  cluster.inference(rdd).count()
  with pytest.raises(Exception, match="FAKE exception after feeding"):
    cluster.shutdown(grace_secs=3)


def test_port_released(sc):
  def fn(args, ctx):
    assert ctx.tmp_socket is None

  TFCluster.run(sc, fn, {}, 2, 0, input_mode=TFCluster.InputMode.TENSORFLOW,
                master_node="chief").shutdown()


def test_port_unreleased(sc):
  def fn(args, ctx):
    import socket
    assert ctx.tmp_socket is not None
    port = ctx.tmp_socket.getsockname()[1]
    s = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
    try:
      s.bind(("0.0.0.0", port))
      raise AssertionError("reserved port could be bound twice")
    except socket.error:
      pass
    ctx.release_port()
    assert ctx.tmp_socket is None

  TFCluster.run(sc, fn, {}, 2, 0, input_mode=TFCluster.InputMode.TENSORFLOW, master_node="chief",
                release_port=False).shutdown()


def test_roles_ps_and_chief(sc):
  import tempfile
  d = tempfile.mkdtemp()

  def fn(args, ctx):
    with open("{}/{}-{}".format(args["d"], ctx.job_name, ctx.task_index), "w") as f:
      f.write("{} {}".format(ctx.rank, sorted(ctx.cluster_spec)))
    if ctx.job_name == "ps":
      time.sleep(600)  # parked until the driver stops it through the control queue

  cluster = TFCluster.run(sc, fn, {"d": d}, 2, 1, input_mode=TFCluster.InputMode.TENSORFLOW,
                          master_node="chief")
  assert sorted(n["job_name"] for n in cluster.cluster_info) == ["chief", "ps"]
  cluster.shutdown()
  import os
  assert sorted(os.listdir(d)) == ["chief-0", "ps-0"]
  assert open(d + "/chief-0").read().startswith("0 ")
  assert open(d + "/ps-0").read().startswith("-1 ")


def test_size_and_mode_validation(sc):
  with pytest.raises(Exception, match="InputMode.TENSORFLOW"):
    TFCluster.run(sc, lambda a, c: None, {}, 2, 0, input_mode=TFCluster.InputMode.SPARK,
                  eval_node=True)
  with pytest.raises(AssertionError):
    TFCluster.run(sc, lambda a, c: None, {}, 2, 2)


def test_argv_list_becomes_sys_argv(sc):
  def fn(argv, ctx):
    import sys
    assert sys.argv == ["prog", "--flag", "1"] and argv == sys.argv

  TFCluster.run(sc, fn, ["prog", "--flag", "1"], 2, 0).shutdown()


def test_subgroups_process_group_and_board_namespaces(sc, monkeypatch):
  """ctx.new_group (torch.distributed sub-group) and the sub-group flavour of the symmetric-memory
  bootstrap (rank / world remapped, board namespace per group) - the hooks for TP/SP/CP groups."""
  import tempfile
  d = tempfile.mkdtemp()

  def fn(args, ctx):
    import json
    import torch
    import torch.distributed as dist
    from tensorflowonspark_b200.parallel import process_group, symm
    ctx.init_process_group(backend="gloo")
    g = ctx.new_group([0, 1])
    t = torch.tensor([float(ctx.rank + 1)])
    dist.all_reduce(t, group=g)
    solo = ctx.new_group([1])              # collective call on both ranks, member on one
    seen = {}

    class FakeComm(object):                # the CUDA-IPC part needs a GPU: capture the bootstrap
      def __init__(self, rank, world, exchange, device):
        seen.update(rank=rank, world=world, gathered=exchange({"r": ctx.rank}))

    real, symm.SymmComm = symm.SymmComm, FakeComm
    try:
      process_group.symm_from_ctx(ctx, ranks=[0, 1])
      pair = dict(seen)
      process_group.symm_from_ctx(ctx, ranks=[ctx.rank])
      alone = dict(seen)
    finally:
      symm.SymmComm = real
    with open("{}/{}".format(args["d"], ctx.rank), "w") as f:
      json.dump({"sum": float(t), "solo_member": solo is not None, "pair": pair, "alone": alone}, f)

  cluster = TFCluster.run(sc, fn, {"d": d}, 2, 0, input_mode=TFCluster.InputMode.TENSORFLOW,
                          master_node="chief")
  cluster.shutdown()
  import json
  res = [json.load(open("{}/{}".format(d, r))) for r in range(2)]
  assert [r["sum"] for r in res] == [3.0, 3.0]
  assert [r["solo_member"] for r in res] == [False, True]
  for r, rec in enumerate(res):
    assert rec["pair"]["rank"] == r and rec["pair"]["world"] == 2
    assert [x["r"] for x in rec["pair"]["gathered"]] == [0, 1]
    assert rec["alone"] == {"rank": 0, "world": 1, "gathered": [{"r": r}]}


def test_gradient_comm_falls_back_to_the_process_group_without_gpus(sc):
  """ctx.gradient_comm(): no GPU (or several hosts) -> GroupComm over torch.distributed; its
  alloc / broadcast / all_reduce are what a trainer and FusedOptimizer's group mode call."""
  import tempfile
  d = tempfile.mkdtemp()

  def fn(args, ctx):
    import json
    import torch
    comm = ctx.gradient_comm()
    w = comm.alloc("weights", 16, torch.float32)
    w.fill_(float(ctx.rank + 5))
    comm.broadcast("weights", root=0)          # the chief's initial values win
    g = torch.full((16,), float(ctx.rank + 1))
    comm.all_reduce(g)
    comm.barrier()
    with open("{}/{}".format(args["d"], ctx.rank), "w") as f:
      json.dump({"kind": type(comm).__name__, "world": comm.world, "rank": comm.rank,
                 "single_host": ctx.single_host, "w": float(w[3]), "g": float(g[7])}, f)

  cluster = TFCluster.run(sc, fn, {"d": d}, 2, 0, input_mode=TFCluster.InputMode.TENSORFLOW,
                          master_node="chief")
  cluster.shutdown()
  import json
  res = [json.load(open("{}/{}".format(d, r))) for r in range(2)]
  for r, rec in enumerate(res):
    assert rec == {"kind": "GroupComm", "world": 2, "rank": r, "single_host": True, "w": 5.0, "g": 3.0}


def test_gradient_comm_builds_the_two_level_communicator_for_workers_on_different_hosts(sc):
  """ctx.gradient_comm() on a cluster whose workers registered from two hosts, one GPU each: the
  selection logic, the per-local-index process groups and HierComm's collectives, with the GPU
  and the second host faked (gloo on CPU; the kernels' side is tools/gpu_check_hier.py)."""
  import tempfile
  d = tempfile.mkdtemp()

  def fn(args, ctx):
    import json
    import torch
    from tensorflowonspark_b200 import TFSparkNode
    ctx.init_process_group(backend="gloo")                    # (a GPU node would get NCCL)
    ctx.gpus = [0]
    ctx.worker_hosts = lambda: ["host-a", "host-b"]           # one worker per "host"
    real_avail, real_dev = torch.cuda.is_available, TFSparkNode.TFNodeContext.device
    torch.cuda.is_available = lambda: True
    TFSparkNode.TFNodeContext.device = property(lambda self: torch.device("cpu"))
    try:
      assert not ctx.single_host
      comm = ctx.gradient_comm()
    finally:
      torch.cuda.is_available, TFSparkNode.TFNodeContext.device = real_avail, real_dev
    w = comm.alloc("weights", 8, torch.float32)
    w.fill_(float(ctx.rank + 5))
    comm.broadcast("weights", root=0)
    g = torch.full((8,), float(ctx.rank + 1))
    comm.all_reduce_inter(g)
    comm.barrier()
    with open("{}/{}".format(args["d"], ctx.rank), "w") as f:
      json.dump({"kind": type(comm).__name__, "hosts": comm.hosts, "local_world": comm.local_world,
                 "local_rank": comm.local_rank, "world": comm.world, "w": float(w[0]), "g": float(g[0])}, f)

  cluster = TFCluster.run(sc, fn, {"d": d}, 2, 0, input_mode=TFCluster.InputMode.TENSORFLOW,
                          master_node="chief")
  cluster.shutdown()
  import json
  for r in range(2):
    rec = json.load(open("{}/{}".format(d, r)))
    assert rec == {"kind": "HierComm", "hosts": 2, "local_world": 1, "local_rank": 0, "world": 2,
                   "w": 5.0, "g": 3.0}
