"""Rendezvous server/client (scenarios of reference tests/test_reservation.py:12-128)."""
import os
import threading
from unittest import mock

import pytest

from tensorflowonspark_b200.reservation import Client, Reservations, Server


def test_reservations_counting():
  r = Reservations(2)
  assert not r.done() and r.remaining() == 2
  r.add({"node": 1})
  assert not r.done() and r.remaining() == 1
  r.add({"node": 2})
  assert r.done() and r.remaining() == 0 and len(r.get()) == 2


def test_server_client_roundtrip():
  s = Server(1)
  addr = s.start()
  c = Client(addr)
  assert c.register({"node": 1, "authkey": b"\x00\x01", "addr": ("h", 5)}) == "OK"
  got = c.await_reservations()
  assert len(got) == 1 and got[0]["node"] == 1 and got[0]["authkey"] == b"\x00\x01"
  assert c.get_reservations() == got
  assert not s.done
  c.request_stop()
  assert s.done
  c.close()


def test_board_all_gather():
  s = Server(2)
  addr = s.start()
  out = {}

  def rank(r):
    c = Client(addr)
    out[r] = c.all_gather("h", r, 2, {"handle": bytes([r]) * 64})
    c.close()

  ts = [threading.Thread(target=rank, args=(r,)) for r in range(2)]
  [t.start() for t in ts]
  [t.join() for t in ts]
  assert out[0] == out[1] and out[0][1]["handle"] == b"\x01" * 64
  s.stop()


def test_server_host_and_port_env():
  with mock.patch.dict(os.environ, {"TFOS_SERVER_HOST": "my_host", "TFOS_SERVER_PORT": "9999"}):
    s = Server(1)
    assert s.get_server_ip() == "my_host" and s.get_server_ports() == [9999]
  with mock.patch.dict(os.environ, {"TFOS_SERVER_PORT": "9997-9999"}):
    assert Server(1).get_server_ports() == [9997, 9998, 9999]


def test_port_range_exhaustion():
  with mock.patch.dict(os.environ, {"TFOS_SERVER_PORT": "38997-38998"}):
    s1, s2, s3 = Server(1), Server(1), Server(1)
    a1, a2 = s1.start(), s2.start()
    assert {a1[1], a2[1]} == {38997, 38998}
    with pytest.raises(Exception):
      s3.start()
    s1.stop()
    s2.stop()


def test_concurrent_clients():
  n = 4
  s = Server(n)
  addr = s.start()

  def reg(i):
    c = Client(addr)
    c.register({"node": i})
    c.await_reservations()
    c.close()

  ts = [threading.Thread(target=reg, args=(i,)) for i in range(n)]
  [t.start() for t in ts]
  [t.join() for t in ts]
  assert sorted(r["node"] for r in s.reservations.get()) == list(range(n))
  s.stop()


def test_await_timeout():
  s = Server(2)
  s.start()
  with pytest.raises(Exception):
    s.await_reservations(timeout=1)
  s.stop()


def test_all_gather_consumes_its_keys_so_a_tag_can_be_reused():
  """a second communicator must never read the first one's (stale) entries."""
  n = 3
  s = Server(1)
  addr = s.start()
  out = {}

  def run(rank, rnd):
    c = Client(addr)
    out[(rnd, rank)] = c.all_gather("symm/job/all/g1/1", rank, n, {"handle": "r{}-round{}".format(rank, rnd)})
    c.close()

  for rnd in (1, 2):   # same tag twice, e.g. a re-created process that restarted its counters
    ts = [threading.Thread(target=run, args=(r, rnd)) for r in range(n)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for r in range(n):
      assert [d["handle"] for d in out[(rnd, r)]] == ["r{}-round{}".format(k, rnd) for k in range(n)]
    assert s._board == {} and s._fetches == {}
  s.stop()


def test_two_communicators_of_one_job_use_distinct_board_tags():
  from tensorflowonspark_b200.parallel import process_group, symm
  tags = []

  class FakeClient(object):
    def __init__(self, addr):
      pass

    def all_gather(self, tag, me, size, obj):
      tags.append(tag)
      return [obj]

  class Ctx(object):
    world_size, rank, cluster_id, server_addr, device = 1, 0, "job42", ("127.0.0.1", 1), "cpu"

  made = []
  with mock.patch("tensorflowonspark_b200.reservation.Client", FakeClient), \
      mock.patch.object(symm, "SymmComm", lambda me, size, exchange, dev: made.append(exchange)):
    process_group.symm_from_ctx(Ctx())
    process_group.symm_from_ctx(Ctx())
  made[0]({"a": 1})
  made[1]({"a": 1})
  made[0]({"a": 2})
  assert len(set(tags)) == 3 and all(t.startswith("symm/job42/all/g") for t in tags)
