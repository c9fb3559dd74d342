"""sparklite engine semantics TensorFlowOnSpark depends on (reference tests/test.py:39-43 smoke test
plus the scheduling quirks listed in SURVEY.md section 7.3 item 4)."""
import os
import time

import pytest

from tensorflowonspark_b200.sparklite import SparkJobError, TaskContext
from tensorflowonspark_b200.sparklite.streaming import StreamingContext


def test_sum(sc):
  assert sc.parallelize(range(1000)).sum() == 499500


def test_executors_are_processes_with_own_cwd(sc):
  info = sc.parallelize(range(8), 8).map(lambda x: (os.getpid(), os.getcwd())).collect()
  pids = set(p for p, _ in info)
  assert len(pids) == 2 and os.getpid() not in pids
  assert len(set(d for _, d in info)) == 2


def test_lazy_and_lineage(sc):
  hits = []
  rdd = sc.parallelize(range(10), 2).map(lambda x: x * 2).filter(lambda x: x % 4 == 0)
  assert hits == []
  assert rdd.collect() == [0, 4, 8, 12, 16]
  assert sc.union([rdd, rdd]).count() == 10
  assert rdd.mapPartitions(lambda it: [sum(it)]).collect() == [12, 28]
  assert sc.parallelize(range(5), 2).zipWithIndex().collect() == [(i, i) for i in range(5)]


def test_worker_reuse_state_persists(sc):
  def bump(it):
    import builtins
    builtins._tfos_counter = getattr(builtins, "_tfos_counter", 0) + 1
    return [builtins._tfos_counter]
  sc.parallelize(range(2), 2).mapPartitions(bump).collect()
  second = sc.parallelize(range(2), 2).mapPartitions(bump).collect()
  assert all(v >= 2 for v in second)


def test_task_error_reaches_driver(sc):
  with pytest.raises(SparkJobError) as e:
    sc.parallelize(range(4), 2).map(lambda x: 1 / 0).collect()
  assert "ZeroDivisionError" in str(e.value)
  assert sc.parallelize(range(4), 2).count() == 4  # context still usable


def test_barrier(sc):
  def f(it):
    ctx = TaskContext.get()
    return [len(ctx.getTaskInfos())]
  assert sc.parallelize(range(2), 2).barrier().mapPartitions(f).collect() == [2, 2]
  with pytest.raises(SparkJobError):
    sc.parallelize(range(3), 3).barrier().mapPartitions(f).collect()


def test_status_tracker_counts_active_tasks(sc):
  import threading
  t = threading.Thread(target=lambda: sc.parallelize(range(2), 2).foreach(lambda x: time.sleep(1.0)))
  t.start()
  time.sleep(0.3)
  st = sc.statusTracker()
  stages = st.getActiveStageIds()
  assert len(st.getActiveJobsIds()) == 1 and st.getStageInfo(stages[0]).numActiveTasks == 2
  t.join()
  assert sc.statusTracker().getActiveJobsIds() == []


def test_text_roundtrip(sc, tmp_path):
  out = str(tmp_path / "txt")
  sc.parallelize(range(10), 2).saveAsTextFile(out)
  assert sorted(int(x) for x in sc.textFile(out).collect()) == list(range(10))


def test_dataframe(sc, spark):
  df = spark.createDataFrame([(1, "a", [1.0, 2.0]), (2, "b", [3.0])], ["x", "y", "z"])
  assert df.dtypes == [("x", "bigint"), ("y", "string"), ("z", "array<double>")]
  rows = df.select("z", "x").rdd.collect()
  assert rows[0].x == 1 and rows[1].z == [3.0]
  assert df.count() == 2 and df.columns == ["x", "y", "z"]


def test_streaming_queue(sc):
  ssc = StreamingContext(sc, 0.2)
  got = []
  stream = ssc.queueStream([sc.parallelize([1, 2]), sc.parallelize([3])])
  stream.foreachRDD(lambda rdd: got.extend(rdd.collect()))
  ssc.start()
  assert not ssc.awaitTerminationOrTimeout(1.0)
  ssc.stop(stopSparkContext=False, stopGraceFully=True)
  assert sorted(got) == [1, 2, 3]


def test_executors_do_not_outlive_a_killed_driver(tmp_path):
  """kill -9 of the driver must not leave executor processes (and their node children) behind."""
  import os
  import signal
  import subprocess
  import sys
  import time
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  pidfile = str(tmp_path / "pids")
  code = (
      "import os, sys, time\n"
      "sys.path.insert(0, {root!r})\n"
      "from tensorflowonspark_b200.sparklite import SparkContext\n"
      "sc = SparkContext('local[2]', 'orphan-test')\n"
      "pids = sc.parallelize(range(2), 2).map(lambda _: os.getpid()).collect()\n"
      "open({pidfile!r}, 'w').write(' '.join(map(str, pids)))\n"
      "time.sleep(600)\n").format(root=root, pidfile=pidfile)
  drv = subprocess.Popen([sys.executable, "-c", code])
  deadline = time.time() + 60
  while not os.path.exists(pidfile) and time.time() < deadline:
    time.sleep(0.2)
  time.sleep(0.3)
  pids = [int(p) for p in open(pidfile).read().split()]
  assert len(set(pids)) == 2
  os.kill(drv.pid, signal.SIGKILL)
  drv.wait()

  def alive(pid):
    try:
      os.kill(pid, 0)
      return open("/proc/{}/stat".format(pid)).read().split(") ")[1][0] != "Z"
    except (OSError, IOError):
      return False

  deadline = time.time() + 15
  while any(alive(p) for p in pids) and time.time() < deadline:
    time.sleep(0.3)
  assert not any(alive(p) for p in pids), "executors survived their driver: {}".format(pids)
