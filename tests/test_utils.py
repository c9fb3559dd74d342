"""Checkpoint / export / metrics helpers and host utilities."""
import os

import pytest
import torch

from tensorflowonspark_b200 import gpu_info, util
from tensorflowonspark_b200.utils import checkpoint, metrics


def test_checkpoint_save_latest_resume_prune(tmp_path):
  d = str(tmp_path / "ckpt")
  assert checkpoint.latest_checkpoint(d) is None and checkpoint.load(d) == (0, None)
  for step in range(1, 9):
    checkpoint.save(d, step * 10, {"w": torch.full((3,), float(step))}, keep=3)
  files = sorted(f for f in os.listdir(d) if f.startswith("ckpt-"))
  assert files == ["ckpt-00000060.pt", "ckpt-00000070.pt", "ckpt-00000080.pt"]
  step, state = checkpoint.load(d)
  assert step == 80 and float(state["w"][0]) == 8.0
  # a torn write (temp file left behind) must not be picked up
  open(os.path.join(d, ".tmp-garbage.pt"), "wb").write(b"xx")
  assert checkpoint.latest_checkpoint(d).endswith("ckpt-00000080.pt")


def test_export_and_load_with_builder(tmp_path):
  from tensorflowonspark_b200.models import simple
  m = simple.Linear(2, 1)
  d = checkpoint.export_model(m, "file://" + str(tmp_path / "exp"), tag_set="serve",
                              signatures={"serving_default": {"inputs": {"x": "x"}}})
  served, sig = checkpoint.load_model(d, "serve")
  out = served(x=[[1.0, 2.0]])["y"]
  assert torch.allclose(out.cpu(), m(torch.tensor([[1.0, 2.0]])))
  assert sig["signatures"]["serving_default"]["inputs"] == {"x": "x"}
  with pytest.raises(ValueError):
    checkpoint.load_model(d, "train")


def test_metrics_reduce_max(tmp_path):
  prefix = str(tmp_path / "m.jsonl")
  for rank, ms in ((0, 10.0), (1, 12.5)):
    log = metrics.StepLogger(prefix, rank)
    log.log(step=1, step_ms=ms)
    log.log(step=2, step_ms=ms + 1)
    log.close()
  assert metrics.reduce_max(prefix) == {1: 12.5, 2: 13.5}


def test_executor_id_file(tmp_path, monkeypatch):
  monkeypatch.chdir(tmp_path)
  with pytest.raises(Exception, match="No executor_id file"):
    util.read_executor_id()
  util.write_executor_id(7)
  assert util.read_executor_id() == 7
  assert util.find_in_path("/nonexistent:" + str(tmp_path), "executor_id") == str(tmp_path / "executor_id")
  assert util.find_in_path("/nonexistent", "executor_id") is False
  assert util.get_ip_address().count(".") == 3


def test_gpu_placement_rules(monkeypatch):
  inv = ([(i, "GPU-%d" % i) for i in range(8)], {"GPU-1"})
  monkeypatch.setattr(gpu_info, "_inventory", lambda: inv)
  assert gpu_info.get_gpus(1, 0) == "0"
  assert gpu_info.get_gpus(2, 1) == "3,4"          # contiguous slice of the *free* list
  assert gpu_info.get_gpus(2, 3, format=gpu_info.AS_LIST) == ["7", "0"]  # wraps around
  assert len(set(gpu_info.get_gpus(3, -1).split(","))) == 3
  monkeypatch.setattr(gpu_info, "MAX_RETRIES", 0)
  with pytest.raises(Exception, match="Unable to find 8 free"):
    gpu_info.get_gpus(8, 0)


def test_fault_spec_parsing_and_actions():
  import time
  from tensorflowonspark_b200.utils import fault
  assert fault.parse("raise:rank=1:step=5;delay:rank=0:step=2:secs=0.5") == [
      {"action": "raise", "rank": 1, "step": 5}, {"action": "delay", "rank": 0, "step": 2, "secs": 0.5}]
  assert fault.maybe_inject(0, 5, "raise:rank=1:step=5") is None
  with pytest.raises(fault.InjectedFault):
    fault.maybe_inject(1, 5, "raise:rank=1:step=5")
  t0 = time.time()
  fault.maybe_inject(0, 2, "delay:rank=0:step=2:secs=0.2")
  assert time.time() - t0 >= 0.2
  assert fault.maybe_inject(0, 1, "drop_feed:rank=0:step=1") == "drop_feed"


def test_injected_fault_surfaces_on_driver(sc, monkeypatch):
  """A rank that dies mid-training must fail the job (never hang): kill and raise variants."""
  from tensorflowonspark_b200 import TFCluster

  def fn(args, ctx):
    from tensorflowonspark_b200.utils import fault
    feed = ctx.get_data_feed()
    step = 0
    while not feed.should_stop():
      feed.next_batch(10)
      fault.maybe_inject(ctx.executor_id, step, args["spec"])
      step += 1

  rdd = sc.parallelize(range(400), 4)
  cluster = TFCluster.run(sc, fn, {"spec": "raise:rank=1:step=3"}, 2, 0,
                          input_mode=TFCluster.InputMode.SPARK)
  with pytest.raises(Exception, match="injected fault|Timeout"):
    cluster.train(rdd, 1, feed_timeout=5)
    cluster.shutdown(grace_secs=1)
  try:
    cluster.shutdown()
  except Exception:
    pass


def test_usable_cpus_and_thread_limit(monkeypatch):
  from tensorflowonspark_b200 import util
  n = util.usable_cpus()
  assert 1 <= n <= (os.cpu_count() or 1)
  monkeypatch.delenv("OMP_NUM_THREADS", raising=False)
  monkeypatch.delenv("MKL_NUM_THREADS", raising=False)
  t = util.limit_intra_op_threads(8)
  assert 1 <= t <= 8 and os.environ["OMP_NUM_THREADS"] == str(t)
  monkeypatch.setenv("OMP_NUM_THREADS", "3")
  assert util.limit_intra_op_threads(1) == 3       # an explicit setting wins


def test_fused_optimizer_shards_partition_every_bucket():
  """The host-side shard arithmetic used to assemble checkpoints must mirror the kernel's:
  shards of a bucket are disjoint, ordered, 8-aligned and cover [begin, end) exactly."""
  from tensorflowonspark_b200.parallel.fused_optim import FusedOptimizer
  for world in (1, 2, 3, 8):
    opt = FusedOptimizer.__new__(FusedOptimizer)
    opt.world = world
    opt.buckets = [(0, 1000, None), (1000, 25_557_040, None), (25_557_040, 25_557_048, None)]
    for i, (b, e, _) in enumerate(opt.buckets):
      pos = b
      for r in range(world):
        lo, hi = opt.shard_bounds(i, r)
        assert lo == pos and lo <= hi <= e and (lo - b) % 8 == 0
        pos = hi
      assert pos == e


def test_fdshare_hands_a_descriptor_to_another_process(tmp_path):
  """parallel/fdshare.py: a descriptor registered by one process is received (SCM_RIGHTS) and is
  usable in another - the transport of cuMem / multicast handles between ranks."""
  import multiprocessing as mp
  import os
  from tensorflowonspark_b200.parallel import fdshare
  path = tmp_path / "payload.bin"
  path.write_bytes(b"nvls-handle-stand-in")
  srv = fdshare.FdServer()
  fd = os.open(str(path), os.O_RDONLY)
  srv.register("mem:grads", fd)

  def child(addr, q):
    try:
      got = fdshare.fetch_fd(addr, "mem:grads")
      q.put(os.read(got, 64))
      try:
        fdshare.fetch_fd(addr, "no-such-key")
        q.put(b"missing key did not raise")
      except KeyError:
        q.put(b"KeyError")
    except Exception as e:  # pragma: no cover
      q.put(repr(e).encode())

  ctx = mp.get_context("fork")
  q = ctx.Queue()
  p = ctx.Process(target=child, args=(srv.address, q))
  p.start()
  assert q.get(timeout=30) == b"nvls-handle-stand-in"
  assert q.get(timeout=30) == b"KeyError"
  p.join(30)
  srv.close()


def test_cluster_launcher_dry_run_plans_the_same_verbs_as_spark_ec2():
  """scripts/cluster_launch.py (counterpart of the reference's scripts/spark_ec2.py): the
  command plan of every verb, without touching a host."""
  import importlib.util
  import os
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  spec = importlib.util.spec_from_file_location("cluster_launch", os.path.join(root, "scripts", "cluster_launch.py"))
  mod = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mod)
  base = ["--hosts", "n0,n1", "--spark-home", "/opt/spark", "--dry-run"]
  rc, r = mod.main(base + ["launch"])
  flat = [" ".join(c) for c in r.log]
  assert rc == 0 and sum("rsync" in c for c in flat) == 2
  assert any("start-master.sh" in c and "n0" in c for c in flat)
  assert sum("start-worker.sh" in c for c in flat) == 2
  assert all("spark://n0:7077" in c for c in flat if "start-worker" in c)
  rc, r = mod.main(base + ["destroy"])
  flat = [" ".join(c) for c in r.log]
  assert any("stop-master.sh" in c for c in flat) and sum("rm -rf" in c for c in flat) == 2
  assert mod.master_url(["a", "b"]) == "spark://a:7077"
  for verb in ("stop", "start", "get-master", "reboot-slaves"):
    assert mod.main(base + [verb])[0] == 0


def test_resnet50_comm_buckets_tile_the_parameter_vector():
  """ResNetTrainer._comm_buckets (overlap of the fused all-reduce with backward): 8-aligned,
  disjoint, covering [0, total); the un-overlappable tail is stem + layer1 only."""
  import torch
  from tensorflowonspark_b200.models import resnet
  from tensorflowonspark_b200.models.engine import BatchNorm, Dense, ParamStore, normal
  from tensorflowonspark_b200.models.resnet import _Block, _Unit

  class Shell(resnet.ResNetTrainer):
    def __init__(self):
      pass

  t, st = Shell(), ParamStore()
  t.store, t.blocks = st, []
  st.register("stem.conv.w", (64, 7, 64), True, None)
  BatchNorm(st, "stem.bn", 64)
  cin = 64
  for si, nblocks in enumerate((3, 4, 6, 3)):
    width = 64 * 2 ** si
    for bi in range(nblocks):
      b, name = _Block(), "layer{}.{}".format(si + 1, bi)
      b.u1, b.u2 = _Unit(st, name + ".u1", cin, width, 1, 1), _Unit(st, name + ".u2", width, width, 3, 1)
      b.u3 = _Unit(st, name + ".u3", width, width * 4, 1, 1)
      b.ds = _Unit(st, name + ".ds", cin, width * 4, 1, 1) if bi == 0 else None
      b.name = name
      t.blocks.append(b)
      cin = width * 4
  Dense(st, "fc", cin, 1000, bias=True, init=normal(0.01))
  for s in st._specs:
    s["init"] = lambda shape, gen: torch.zeros(shape)
  st.finalize(torch.device("cpu"))
  buckets = t._comm_buckets()
  cov = sorted((b, e) for b, e, _ in buckets)
  assert all(b % 8 == 0 and e % 8 == 0 for b, e in cov)
  assert cov[0][0] == 0 and cov[-1][1] == st.total
  assert all(cov[i][1] == cov[i + 1][0] for i in range(len(cov) - 1))
  tail = sum(e - b for b, e, tag in buckets if tag == "stem")
  assert tail < 0.011 * st.total, tail      # ~1 % of the parameters are left for the tail


def test_run_with_restarts_retries_on_a_fresh_context_and_gives_up_eventually():
  from tensorflowonspark_b200.utils import recovery

  class Ctx(object):
    made = []

    def __init__(self):
      self.stopped = self.cancelled = 0
      Ctx.made.append(self)

    def stop(self):
      self.stopped += 1

    def cancelAllJobs(self):
      self.cancelled += 1

  seen, hooks = [], []

  def job(sc, attempt):
    seen.append((sc, attempt))
    if attempt == 0:
      raise SystemExit(1)          # how TFCluster.shutdown leaves a failed application
    if attempt == 1:
      raise RuntimeError("node lost")

  n = recovery.run_with_restarts(Ctx, job, max_restarts=3, backoff_s=0.0,
                                 on_failure=lambda a, e: hooks.append((a, type(e).__name__)))
  assert n == 3 and [a for _, a in seen] == [0, 1, 2]
  assert len({id(sc) for sc, _ in seen}) == 3                      # a new context per attempt
  assert hooks == [(0, "SystemExit"), (1, "RuntimeError")]
  assert all(c.stopped >= 1 for c in Ctx.made) and Ctx.made[0].cancelled == 1 and Ctx.made[2].cancelled == 0

  with pytest.raises(recovery.JobFailed) as info:
    recovery.run_with_restarts(Ctx, lambda sc, a: (_ for _ in ()).throw(ValueError(a)), max_restarts=1, backoff_s=0.0)
  assert [type(c).__name__ for c in info.value.causes] == ["ValueError", "ValueError"]
  with pytest.raises(KeyboardInterrupt):                           # never swallowed
    recovery.run_with_restarts(Ctx, lambda sc, a: (_ for _ in ()).throw(KeyboardInterrupt()), backoff_s=0.0)


def test_checkpoints_exports_and_event_files_on_a_remote_filesystem():
  """model_dir / export_dir / log_dir may be URIs of any filesystem pyarrow resolves (hdfs://,
  s3:// ... - what ``ctx.absolute_path`` yields when the cluster's defaultFS is not local, and
  what lets another host resume or serve what the chief wrote).  pyarrow's in-memory ``mock://``
  filesystem stands in for the remote store."""
  pytest.importorskip("pyarrow")
  from tensorflowonspark_b200.models import simple
  from tensorflowonspark_b200.utils import fs, summary
  root = "mock:///jobs/run1"
  md, ed = root + "/model", root + "/export"
  assert not fs.is_local(md) and fs.is_local("file:///tmp/x") and fs.is_local("/tmp/x")
  assert checkpoint.latest_checkpoint(md) is None and checkpoint.load(md) == (0, None)
  model = simple.Linear(2, 1)
  for step in (10, 20, 30, 40):
    path = checkpoint.save(md, step, {"w": torch.full((4,), float(step))}, keep=2, model=model)
  assert path == md + "/ckpt-00000040.pt" and checkpoint.latest_checkpoint(md) == path
  names = sorted(fs.listdir(md))
  assert names == ["checkpoint", "ckpt-00000030.pt", "ckpt-00000040.pt", "signature.json"]   # pruned, no temp files
  step, state = checkpoint.load(md)
  assert step == 40 and torch.equal(state["w"], torch.full((4,), 40.0))
  assert checkpoint.load(md + "/ckpt-00000030.pt")[0] == 30
  # export -> load_model, and serving the newest checkpoint of model_dir without an export
  checkpoint.save(md, 50, model.state_dict(), model=model)
  checkpoint.export_model(model, ed, signatures={"serving_default": {"inputs": {"x": "x"}, "outputs": {"y": "y"}}})
  served, sig = checkpoint.load_model(ed, "serve")
  x = torch.tensor([[1.0, 2.0]])
  assert torch.allclose(served(x=x.numpy())["y"].cpu(), model(x).detach())
  served2, _ = checkpoint.load_model_dir(md)
  assert torch.allclose(served2(x=x.numpy())["y"].cpu(), model(x).detach())
  with pytest.raises(ValueError):
    checkpoint.load_model(ed, "other_tag")
  # TensorBoard events: spooled locally, mirrored to the remote log directory at every flush
  w = summary.SummaryWriter(md, flush_secs=1e9, max_queue=2)
  w.add_scalar("loss", 1.0, 1)
  w.add_scalar("loss", 0.5, 2)                 # second event reaches max_queue -> flush -> upload
  (remote,) = [n for n in fs.listdir(md) if n.startswith("events.out.tfevents.")]
  local_copy = os.path.join(os.path.dirname(w.path), "check")
  with fs.open_read(md + "/" + remote) as src, open(local_copy, "wb") as dst:
    dst.write(src.read())
  assert [e["scalars"].get("loss") for e in summary.read_events(local_copy)] == [None, 1.0, 0.5]
  w.close()
  with pytest.raises(IOError, match="client library"):
    checkpoint.save("nosuchscheme://host/dir", 1, {})


def test_prometheus_exporter_serves_what_the_step_logger_logs(tmp_path):
  import urllib.request
  if not metrics.Exporter.available():
    pytest.skip("prometheus_client not installed")
  exp = metrics.Exporter(rank=3)
  other = metrics.Exporter(rank=4)              # private registries: two exporters in one process
  log = metrics.StepLogger(str(tmp_path / "m.jsonl"), rank=3, exporter=exp)
  log.log(step=7, loss=0.25, images_per_s=13428.0, note="text fields are skipped", ok=True)
  log.log(step=8, loss=0.125, **{"h2d MB/s": 1400.5})
  body = urllib.request.urlopen(exp.url(), timeout=5).read().decode()
  assert 'tfos_loss{rank="3"} 0.125' in body and 'tfos_step{rank="3"} 8.0' in body
  assert 'tfos_images_per_s{rank="3"} 13428.0' in body and 'tfos_h2d_MB_s{rank="3"} 1400.5' in body
  assert "tfos_last_log_timestamp_seconds" in body and "tfos_note" not in body and "tfos_ok" not in body
  assert "tfos_loss" not in urllib.request.urlopen(other.url(), timeout=5).read().decode()
  assert [r["step"] for r in metrics.read(log.path)] == [7, 8]      # the JSONL side is unchanged
  log.close()
  exp.close()
  other.close()


def test_package_cli_info_env_and_stop_streaming(capsys, monkeypatch):
  """python -m tensorflowonspark_b200: info / env / stop-streaming (reference
  examples/utils/stop_streaming.py sends STOP to the reservation server of a streaming job)."""
  import json
  from tensorflowonspark_b200 import __main__ as cli, reservation
  assert cli.main(["info"]) == 0
  info = json.loads(capsys.readouterr().out)
  assert info["torch"] == torch.__version__ and "extension" in info and isinstance(info["gpus"], (list, str))
  monkeypatch.setenv("TFOS_NVLS", "0")
  assert cli.main(["env"]) == 0 and "TFOS_NVLS=0" in capsys.readouterr().out
  server = reservation.Server(1)
  host, port = server.start()
  assert not server.done
  assert cli.main(["stop-streaming", host, str(port)]) == 0
  deadline = __import__("time").time() + 5
  while not server.done and __import__("time").time() < deadline:
    __import__("time").sleep(0.05)
  assert server.done
  server.stop()
  assert cli.main(["stop-streaming", "onlyhost"]) == 2 and cli.main(["no-such-command"]) == 2
  assert cli.main([]) == 0 and "stop-streaming HOST PORT" in capsys.readouterr().out
