"""Life-cycle paths the reference never tested (SURVEY.md section 4 "Gaps"): evaluator side-car,
ps nodes on the driver, DStream feeding + streaming shutdown, TensorBoard launch / teardown,
reservation timeout, the shutdown watchdog.  All on 2 executor processes, CPU only."""
import os
import stat
import tempfile
import time

import pytest

from tensorflowonspark_b200 import TFCluster, reservation

# a stand-in "tensorboard" executable: must be on PATH before the executors fork (module fixture)
_FAKE_DIR = tempfile.mkdtemp(prefix="tfos_fake_tb_")
_FAKE_TB = os.path.join(_FAKE_DIR, "tensorboard")
with open(_FAKE_TB, "w") as _f:
  _f.write("#!/bin/sh\necho \"$@\" > {}/args.txt\nexec sleep 600\n".format(_FAKE_DIR))
os.chmod(_FAKE_TB, os.stat(_FAKE_TB).st_mode | stat.S_IEXEC)
os.environ["PATH"] = _FAKE_DIR + os.pathsep + os.environ.get("PATH", "")
os.environ["TFOS_HEARTBEAT_TIMEOUT"] = "3"   # likewise inherited by the executors at fork time


@pytest.fixture(scope="module", autouse=True)
def _heartbeat_env_only_for_this_module():
  yield
  os.environ.pop("TFOS_HEARTBEAT_TIMEOUT", None)   # later modules fork their executors without it


def _alive(pid):
  try:
    os.kill(pid, 0)
    return True
  except OSError:
    return False


def test_evaluator_sidecar(sc):
  d = tempfile.mkdtemp()

  def fn(args, ctx):
    with open("{}/{}-{}".format(args["d"], ctx.job_name, ctx.task_index), "w") as f:
      f.write("{} {}".format(ctx.rank, ctx.world_size))
    if ctx.job_name == "evaluator":
      time.sleep(600)  # an evaluator polls checkpoints until the driver stops it

  cluster = TFCluster.run(sc, fn, {"d": d}, 2, 0, input_mode=TFCluster.InputMode.TENSORFLOW,
                          master_node="chief", eval_node=True)
  assert sorted(n["job_name"] for n in cluster.cluster_info) == ["chief", "evaluator"]
  t0 = time.time()
  cluster.shutdown()
  assert time.time() - t0 < 60  # the parked evaluator was stopped through its control queue
  assert sorted(os.listdir(d)) == ["chief-0", "evaluator-0"]
  # the evaluator is outside the training world (no collective rank)
  assert open(d + "/chief-0").read() == "0 1"
  assert open(d + "/evaluator-0").read().startswith("-1")


def test_driver_ps_nodes(sc):
  d = tempfile.mkdtemp()

  def fn(args, ctx):
    with open("{}/{}-{}".format(args["d"], ctx.job_name, ctx.task_index), "w") as f:
      f.write(str(os.getpid()))
    if ctx.job_name == "ps":
      time.sleep(600)

  cluster = TFCluster.run(sc, fn, {"d": d}, 3, 1, input_mode=TFCluster.InputMode.TENSORFLOW,
                          driver_ps_nodes=True)
  roles = sorted(n["job_name"] for n in cluster.cluster_info)
  assert roles == ["ps", "worker", "worker"]
  cluster.shutdown()
  assert sorted(os.listdir(d)) == ["ps-0", "worker-0", "worker-1"]
  # the ps node's user function ran on the driver host, outside the executor processes
  pids = {name: int(open(os.path.join(d, name)).read()) for name in os.listdir(d)}
  assert pids["ps-0"] not in (pids["worker-0"], pids["worker-1"])
  with pytest.raises(Exception, match="InputMode.TENSORFLOW"):
    TFCluster.run(sc, fn, {"d": d}, 2, 1, input_mode=TFCluster.InputMode.SPARK,
                  driver_ps_nodes=True)


def test_dstream_feed_and_streaming_shutdown(sc):
  from tensorflowonspark_b200.sparklite.streaming import StreamingContext
  out = tempfile.mkdtemp() + "/seen"

  def fn(args, ctx):
    feed = ctx.get_data_feed(train_mode=True)
    seen = 0
    while not feed.should_stop():
      rows = feed.next_batch(16)
      seen += len(rows)
      with open(args["out"] + str(ctx.executor_id), "w") as f:
        f.write(str(seen))
      if seen >= 40:
        feed.terminate()   # tells the reservation server to stop -> ends the stream

  ssc = StreamingContext(sc, 0.2)
  stream = ssc.queueStream([sc.parallelize(range(100), 2) for _ in range(3)])
  cluster = TFCluster.run(sc, fn, {"out": out}, 2, 0, input_mode=TFCluster.InputMode.SPARK)
  cluster.train(stream, feed_timeout=60)
  ssc.start()
  t0 = time.time()
  cluster.shutdown(ssc)
  assert time.time() - t0 < 90
  counts = [int(open(out + str(i)).read()) for i in range(2)]
  assert all(c >= 40 for c in counts), counts


def test_tensorboard_launch_and_teardown(sc):
  logdir = tempfile.mkdtemp()
  cluster = TFCluster.run(sc, lambda a, c: None, {}, 2, 0, tensorboard=True, log_dir=logdir)
  url = cluster.tensorboard_url()
  assert url is not None and url.startswith("http://")
  tb = [n for n in cluster.cluster_info if n.get("tb_pid")]
  assert len(tb) == 1 and tb[0]["job_name"] in ("worker", "chief") and tb[0]["task_index"] == 0
  pid, port = tb[0]["tb_pid"], tb[0]["tb_port"]
  assert url.endswith(":{}".format(port)) and _alive(pid)
  deadline = time.time() + 5
  while not os.path.exists(_FAKE_DIR + "/args.txt") and time.time() < deadline:
    time.sleep(0.1)
  args = open(_FAKE_DIR + "/args.txt").read()
  assert "--logdir=" + logdir in args and "--port={}".format(port) in args
  cluster.shutdown()
  deadline = time.time() + 10
  while _alive(pid) and time.time() < deadline:
    time.sleep(0.2)
  # a zombie still answers kill(0): reap-insensitive check through /proc
  state = ""
  try:
    state = open("/proc/{}/stat".format(pid)).read().split(") ")[1][0]
  except (IOError, OSError):
    pass
  assert state in ("", "Z"), "TensorBoard (pid {}) survived shutdown".format(pid)


def test_reservation_timeout_reports_missing_nodes():
  server = reservation.Server(3)
  addr = server.start()
  client = reservation.Client(addr)
  client.register({"executor_id": 0, "host": "h", "job_name": "worker", "task_index": 0})
  t0 = time.time()
  with pytest.raises(Exception, match="(?i)timed out|timeout"):
    server.await_reservations(timeout=2)
  assert 1.5 < time.time() - t0 < 15
  client.request_stop()
  client.close()
  server.stop()


def test_cluster_run_gives_up_when_nodes_are_missing(sc):
  # 3 nodes requested on a 2-executor engine whose third task can never register in time
  def fn(args, ctx):
    time.sleep(0.1)

  with pytest.raises(Exception):
    TFCluster.run(sc, fn, {}, 2, 0, reservation_timeout=0)


def test_silent_node_death_is_detected_by_heartbeat(sc):
  """A node process that vanishes without raising (OOM kill, CUDA trap taking the process down)
  never writes to its error queue; the feeder must notice the missing heartbeats long before
  feed_timeout."""
  def fn(args, ctx):
    feed = ctx.get_data_feed(train_mode=True)
    feed.next_batch(4)
    os._exit(17)          # no exception, no error-queue entry, heartbeat thread dies with us

  cluster = TFCluster.run(sc, fn, {}, 2, 0, input_mode=TFCluster.InputMode.SPARK)
  t0 = time.time()
  with pytest.raises(Exception, match="heartbeat"):
    cluster.train(sc.parallelize(range(1000), 2), feed_timeout=120)
  assert time.time() - t0 < 60
  try:
    cluster.shutdown(grace_secs=0, timeout=30)
  except BaseException:
    pass


def test_shutdown_watchdog_fires(sc):
  """shutdown(timeout=N): a node that never finishes must not hang the driver forever."""
  def fn(args, ctx):
    time.sleep(120)   # foreground worker in InputMode.TENSORFLOW: shutdown waits for it

  cluster = TFCluster.run(sc, fn, {}, 2, 0, input_mode=TFCluster.InputMode.TENSORFLOW)
  t0 = time.time()
  with pytest.raises((Exception, SystemExit)) as ei:
    cluster.shutdown(timeout=3)
  assert time.time() - t0 < 60
  assert "imeout" in str(ei.value) or isinstance(ei.value, SystemExit)
