"""Parameter server over TCP (parallel/ps_net.py): the transport for ps and workers on different
hosts.  The reference's ps is a network service (tf.train.Server / gRPC, TFNode.py:126-132); the
peer-mapped server of parallel/ps.py only spans one machine.  Same contracts as test_ps.py, over
sockets: plain hogwild push / pull / sparse rows split across servers, slot mode with the
optimizer state on the servers, bf16 + fp32 model pulls, and transport selection from the
cluster spec - plus a whole TFCluster run with the transport forced."""
import numpy as np
import pytest

from tensorflowonspark_b200 import TFCluster


def _cluster(tag, spec=None):
  from tensorflowonspark_b200 import reservation
  srv = reservation.Server(1)
  addr = srv.start()
  spec = spec or {"ps": ["10.0.0.1:1", "10.0.0.2:2"], "chief": ["10.0.0.3:3"], "worker": ["10.0.0.4:4"]}

  class Ctx(object):
    def __init__(self, job, idx):
      self.job_name, self.task_index = job, idx
      self.cluster_spec = spec
      self.cluster_id, self.server_addr, self.gpus = "net-" + tag, addr, []
  return srv, Ctx


def test_transport_follows_the_hosts_of_the_cluster_spec(monkeypatch):
  from tensorflowonspark_b200.parallel import ps
  srv, Ctx = _cluster("t")
  assert ps.transport(Ctx("worker", 0)) == "tcp"                      # four hosts
  one = type("C", (), {"cluster_spec": {"ps": ["h:1"], "worker": ["h:2", "h:3"]}})()
  assert ps.transport(one) == "ipc"
  monkeypatch.setenv("TFOS_PS_TRANSPORT", "tcp")
  assert ps.transport(one) == "tcp"
  srv.stop()


def test_plain_mode_pull_push_and_sparse_rows_split_across_servers():
  from tensorflowonspark_b200.parallel import ps, ps_net
  srv, Ctx = _cluster("plain")
  numel, width, base = 40, 5, 3          # slices [0, 24) and [24, 40): row 4 = [23, 28) straddles
  servers = [ps.attach(Ctx("ps", i), params=numel) for i in range(2)]
  assert all(isinstance(s, ps_net.NetPSServer) for s in servers)
  assert (servers[0].lo, servers[0].hi, servers[1].lo) == (0, 24, 24)
  client = ps.attach(Ctx("worker", 0))
  assert isinstance(client, ps_net.NetPSClient) and isinstance(ps.PSClient(Ctx("chief", 0)), ps_net.NetPSClient)
  target = np.arange(numel, dtype=np.float32)
  for _ in range(150):                   # least squares by asynchronous SGD through the sockets
    w = client.pull()
    client.push(2.0 * (w - target), lr=0.05)
  assert np.allclose(client.pull(), target, atol=1e-3)
  before = client.pull()
  rows = np.arange(3 * width, dtype=np.float32).reshape(3, width) + 1.0
  client.push_sparse(rows, np.array([4, 0, 6]), width=width, base=base, lr=1.0)
  want = before.copy()
  for row, r in zip(rows, (4, 0, 6)):
    want[base + r * width:base + (r + 1) * width] -= row
  assert np.allclose(client.pull(), want, atol=1e-6)
  assert np.allclose(np.concatenate([s.values() for s in servers]), want, atol=1e-6)
  import torch
  out32, out16 = torch.zeros(numel), torch.zeros(numel, dtype=torch.bfloat16)
  client.pull(out_fp32=out32, out_bf16=out16)
  assert np.allclose(out32.numpy(), want, atol=1e-6) and np.allclose(out16.float().numpy(), want, rtol=1e-2)
  client.close()
  for s in servers:
    s.close()
  srv.stop()


@pytest.mark.parametrize("opt", ["sgd", "momentum", "adam"])
def test_slot_mode_optimizer_on_the_servers_matches_a_sequential_reference(opt):
  from tensorflowonspark_b200.parallel import ps
  numel, decay_end, ema_begin = 48, 24, 40
  lr, mom, wd = 0.1, 0.9, 0.01
  srv, Ctx = _cluster("slots-" + opt)
  init = np.linspace(-1, 1, numel).astype(np.float32)
  servers = [ps.attach(Ctx("ps", i), params=init, optimizer=opt, lr=lr, momentum=mom, weight_decay=wd,
                       decay_end=decay_end, ema_begin=ema_begin) for i in range(2)]
  clients = [ps.attach(Ctx("chief", 0)), ps.attach(Ctx("worker", 0))]
  assert [c.client_id for c in clients] == [0, 1] and clients[0].slot_mode
  w = init.astype(np.float64).copy()
  m, v, t = np.zeros(numel), np.zeros(numel), 0
  rng = np.random.RandomState(0)
  for step in range(7):
    for c in clients:
      g = rng.randn(numel).astype(np.float32)
      c.push_grads(g)                      # returns before the acknowledgement ...
      assert c.applies() == [step * 2 + clients.index(c) + 1] * 2    # ... which applies() collects
      gg = g.astype(np.float64).copy()
      tr = slice(0, ema_begin)
      w[ema_begin:] -= gg[ema_begin:]
      gt = gg[tr].copy()
      gt[:decay_end] += wd * w[:decay_end]
      if opt == "momentum":
        m[tr] = mom * m[tr] + gt
        gt = m[tr]
      elif opt == "adam":
        t += 1
        m[tr] = 0.9 * m[tr] + 0.1 * gt
        v[tr] = 0.999 * v[tr] + 0.001 * gt * gt
        gt = (m[tr] / (1 - 0.9 ** t)) / (np.sqrt(v[tr] / (1 - 0.999 ** t)) + 1e-7)
      w[tr] -= lr * gt
      got = c.pull_model()
      assert np.allclose(got, w, rtol=2e-4, atol=2e-5), (opt, step, np.abs(got - w).max())
  assert sum(s.poll_once() for s in servers) == 28 and sum(s.poll_once() for s in servers) == 0
  clients[1].set_lr(0.0)                   # hyper-parameters live on the servers
  clients[0].push_grads(np.ones(numel, np.float32))
  after = clients[0].pull_model()
  if opt != "adam":
    assert np.allclose(after[:ema_begin], w[:ema_begin], rtol=2e-4, atol=2e-5)
  for c in clients:
    c.close()
  for s in servers:
    s.close()
  srv.stop()


def test_model_pull_ships_bf16_weights_fp32_tail_and_running_statistics():
  """The tensor-filling pull a native trainer uses (PSWorker.pull): bf16 for the decayed weights,
  fp32 for batch-norm scale/offset + biases (aux32) and for the running statistics; the next
  push sends the running statistics as (pulled - current) deltas."""
  import torch
  from tensorflowonspark_b200.parallel import ps
  total, decay_end, nrun = 32, 16, 8
  numel = total + nrun
  srv, Ctx = _cluster("model")
  init = np.concatenate([np.linspace(-2, 2, total), np.linspace(10, 11, nrun)]).astype(np.float32)
  servers = [ps.attach(Ctx("ps", i), params=init, optimizer="sgd", lr=0.5, decay_end=decay_end,
                       ema_begin=total) for i in range(2)]           # slices [0, 24) and [24, 40)
  c = ps.attach(Ctx("worker", 0))
  weights = torch.zeros(total, dtype=torch.bfloat16)
  aux32, running = torch.zeros(total - decay_end), torch.zeros(nrun)
  c.pull_model(weights, aux32, running, decay_end=decay_end, total=total)
  assert torch.equal(weights[:decay_end], torch.from_numpy(init[:decay_end]).to(torch.bfloat16))
  assert np.array_equal(aux32.numpy(), init[decay_end:total])          # fp32, exact
  assert np.array_equal(running.numpy(), init[total:])
  assert torch.equal(weights[decay_end:], torch.from_numpy(init[decay_end:total]).to(torch.bfloat16))
  grads = torch.ones(total)
  new_running = running + 0.25                                         # the step moved the statistics
  c.push_grads(grads, new_running, total=total)
  got = c.pull()
  assert np.allclose(got[:total], init[:total] - 0.5, atol=1e-6)
  assert np.allclose(got[total:], init[total:] + 0.25, atol=1e-6)      # w -= (pulled - current)
  c.close()
  for s in servers:
    s.close()
  srv.stop()


def _fn(args, ctx):
  import numpy as np
  from tensorflowonspark_b200.parallel import ps_net
  cluster_spec, server = ctx.start_cluster_server(params=np.zeros(16, dtype=np.float32),
                                                     transport="tcp")   # one host: forced
  if ctx.job_name == "ps":
    assert isinstance(server.ps, ps_net.NetPSServer)
    server.join()
    return
  ps = server.ps
  assert isinstance(ps, ps_net.NetPSClient)
  target = np.arange(16, dtype=np.float32)
  for _ in range(200):
    w = ps.pull()
    ps.push(2.0 * (w - target), lr=0.05)
  np.save(args["out"], ps.pull())


def test_async_ps_over_tcp_through_tfcluster(sc, tmp_path):
  out = str(tmp_path / "w.npy")
  cluster = TFCluster.run(sc, _fn, {"out": out}, 2, 1, input_mode=TFCluster.InputMode.TENSORFLOW)
  cluster.shutdown()
  assert np.allclose(np.load(out), np.arange(16, dtype=np.float32), atol=1e-3)
