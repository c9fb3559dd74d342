"""Asynchronous parameter-server path on CPU executors (shared-memory twin of the GPU kernels):
1 ps + 1 worker through TFCluster, the role life-cycle the reference never tested
(SURVEY.md section 4 "Gaps")."""
import numpy as np

from tensorflowonspark_b200 import TFCluster


def _fn(args, ctx):
  import numpy as np
  cluster_spec, server = ctx.start_cluster_server(params=np.zeros(16, dtype=np.float32))
  assert sorted(cluster_spec) == ["ps", "worker"]
  if ctx.job_name == "ps":
    server.join()
    return
  ps = server.ps
  target = np.arange(16, dtype=np.float32)
  for _ in range(200):                      # least squares on w by asynchronous SGD
    w = ps.pull()
    ps.push(2.0 * (w - target), lr=0.05)
  ps.push_sparse(np.ones((2, 4), np.float32), np.array([0, 3]), width=4, lr=1.0)
  np.save(args["out"], ps.pull())


def test_async_ps_cpu(sc, tmp_path):
  out = str(tmp_path / "w.npy")
  cluster = TFCluster.run(sc, _fn, {"out": out}, 2, 1, input_mode=TFCluster.InputMode.TENSORFLOW)
  cluster.shutdown()
  w = np.load(out)
  want = np.arange(16, dtype=np.float32)
  want[0:4] -= 1.0
  want[12:16] -= 1.0
  assert np.allclose(w, want, atol=1e-3), w


def test_sparse_rows_straddling_a_server_boundary_are_split_not_dropped():
  """with num_ps > 1 a table row cut by a slice boundary must still be trained."""
  from tensorflowonspark_b200 import reservation
  from tensorflowonspark_b200.parallel import ps

  srv = reservation.Server(1)
  addr = srv.start()

  class Ctx(object):
    def __init__(self, job, idx):
      self.job_name, self.task_index = job, idx
      self.cluster_spec = {"ps": ["a:1", "b:2"], "worker": ["c:3"]}
      self.cluster_id, self.server_addr, self.gpus = "straddle", addr, []

  numel, width, base = 40, 5, 3          # slices [0, 24) and [24, 40): row 4 = [23, 28) straddles
  servers = [ps.PSServer(Ctx("ps", i), numel) for i in range(2)]
  assert (servers[0].lo, servers[0].hi, servers[1].lo) == (0, 24, 24)
  client = ps.PSClient(Ctx("worker", 0))
  rows = np.arange(3 * width, dtype=np.float32).reshape(3, width) + 1.0
  client.push_sparse(rows, np.array([4, 0, 6]), width=width, base=base, lr=1.0)
  got = client.pull()
  want = np.zeros(numel, np.float32)
  for row, r in zip(rows, (4, 0, 6)):
    want[base + r * width:base + (r + 1) * width] -= row
  assert np.array_equal(got, want), (got, want)
  client.close()
  for s_ in servers:
    s_.close()
  srv.stop()
