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
