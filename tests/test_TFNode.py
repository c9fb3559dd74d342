"""TFNode helpers (scenarios of reference tests/test_TFNode.py:8-58) + ring/chunk feed paths."""
import getpass
import os

from tensorflowonspark_b200 import TFManager, TFNode, marker, shmring


def test_hdfs_path():
  cwd, user = os.getcwd(), getpass.getuser()
  fs = ["file://", "hdfs://", "viewfs://"]
  paths = {
      "hdfs://foo/bar": ["hdfs://foo/bar"] * 3,
      "viewfs://foo/bar": ["viewfs://foo/bar"] * 3,
      "file://foo/bar": ["file://foo/bar"] * 3,
      "/foo/bar": ["file:///foo/bar", "hdfs:///foo/bar", "viewfs:///foo/bar"],
      "foo/bar": ["file://{}/foo/bar".format(cwd), "hdfs:///user/{}/foo/bar".format(user),
                  "viewfs:///user/{}/foo/bar".format(user)],
  }
  for i, f in enumerate(fs):
    ctx = type("MockContext", (), {"defaultFS": f, "working_dir": cwd})
    for path, expected in paths.items():
      assert TFNode.hdfs_path(ctx, path) == expected[i]


def test_datafeed_row_items():
  mgr = TFManager.start(b"abc", ["input", "output"], "local")
  q = mgr.get_queue("input")
  for i in range(10):
    q.put(i)
  q.put(None)
  feed = TFNode.DataFeed(mgr)
  assert len(feed.next_batch(2)) == 2
  assert len(feed.next_batch(4)) == 4
  assert not feed.should_stop()
  assert len(feed.next_batch(10)) == 4  # short batch at end of feed
  assert feed.should_stop()
  mgr.shutdown()


def test_datafeed_chunks_mapping_and_partitions():
  mgr = TFManager.start(b"abc", ["input", "output"], "local")
  q = mgr.get_queue("input")
  q.put(marker.Rows([[i, i * 2] for i in range(5)]))
  q.put(marker.EndPartition())
  q.put(marker.Rows([[i, i * 2] for i in range(5, 8)]))
  q.put(marker.EndPartition())
  q.put(None)
  feed = TFNode.DataFeed(mgr, train_mode=False, input_mapping={"a": "x", "b": "y"})
  b = feed.next_batch(4)
  assert b == {"x": [0, 1, 2, 3], "y": [0, 2, 4, 6]}
  assert feed.next_batch(4) == {"x": [4], "y": [8]}  # stops at the partition boundary
  feed.batch_results([1, 2, 3])
  assert mgr.get_queue("output").get().rows == [1, 2, 3]
  assert feed.next_batch(4) == {"x": [5, 6, 7], "y": [10, 12, 14]}
  assert feed.next_batch(4) == {"x": [], "y": []} and feed.should_stop()
  mgr.shutdown()


def test_datafeed_ring_blocks():
  mgr = TFManager.start(b"abc", ["input", "output"], "local")
  name, ring = shmring.create(4, 1 << 20)
  mgr.set("ring", {"name": name, "nslots": 4, "slot_bytes": 1 << 20})
  rows = [[float(i), [i, i + 1, i + 2]] for i in range(100)]
  blk = shmring.pack_rows(ring, rows)
  assert isinstance(blk, marker.RingBlock) and blk.nrows == 100
  assert shmring.pack_rows(ring, [["a", 1]]) is None  # non-numeric rows use the queue path
  q = mgr.get_queue("input")
  q.put(blk)
  q.put(None)
  feed = TFNode.DataFeed(mgr)
  got = feed.next_batch(64) + feed.next_batch(64)
  assert got == rows and feed.should_stop()
  # the consumed slot was released: the 4-slot ring accepts 4 more blocks without blocking
  for _ in range(4):
    assert shmring.pack_rows(ring, rows[:3], timeout=2.0) is not None
  mgr.shutdown()


def test_terminate_drains_queue():
  mgr = TFManager.start(b"abc", ["input", "output"], "local")
  q = mgr.get_queue("input")
  for i in range(5):
    q.put(i)
  feed = TFNode.DataFeed(mgr)
  feed.next_batch(2)
  import tensorflowonspark_b200.TFNode as tfn
  orig = tfn._queue_mod.Empty
  feed.queue_in.get  # noqa: B018
  # shorten the idle timeout for the test
  real_get = feed.queue_in.get
  feed.queue_in.get = lambda block=True, timeout=5: real_get(block, 0.3)
  feed.terminate()
  assert str(mgr.get("state")).strip("'") == "terminating"
  q.join()  # every queued item was task_done()'d
  assert orig is tfn._queue_mod.Empty
  mgr.shutdown()


def test_next_batch_arrays_from_ring_and_rows():
  import numpy as np
  mgr = TFManager.start(b"abc", ["input", "output"], "local")
  name, ring = shmring.create(4, 1 << 20)
  mgr.set("ring", {"name": name, "nslots": 4, "slot_bytes": 1 << 20})
  rows = [(np.full((4, 4, 3), i, np.uint8), np.full((4, 4), i % 3, np.uint8)) for i in range(10)]
  q = mgr.get_queue("input")
  q.put(shmring.pack_rows(ring, rows[:6]))
  q.put(marker.Rows([(r[0].tolist(), r[1].tolist()) for r in rows[6:]]))   # mixed transports
  q.put(None)
  feed = TFNode.DataFeed(mgr)
  a = feed.next_batch_arrays(4)
  assert a[0].shape == (4, 4, 4, 3) and a[0].dtype == np.uint8 and a[1].shape == (4, 4, 4)
  a2 = feed.next_batch_arrays(4)          # spans the ring block and the python-rows chunk
  assert a2[0].shape == (4, 4, 4, 3)
  assert [int(x[0, 0, 0]) for x in a[0]] + [int(x[0, 0, 0]) for x in a2[0]] == list(range(8))
  b = feed.next_batch_arrays(8)
  assert b[0].shape[0] == 2 and int(b[1][1, 0, 0]) == 9 % 3 and feed.should_stop()
  mgr.shutdown()


def test_array_views_stay_valid_while_feeder_reuses_slots():
  """next_batch_arrays hands out zero-copy views of a ring slot; the slot must not go back to
  the feeders before DataFeed.HOLD_CALLS further calls - a feeder that keeps writing must not
  be able to scribble over a batch the consumer still holds."""
  import numpy as np
  mgr = TFManager.start(b"abc", ["input", "output"], "local")
  name, ring = shmring.create(2, 1 << 16)   # 2 small slots: reuse pressure is immediate
  mgr.set("ring", {"name": name, "nslots": 2, "slot_bytes": 1 << 16})
  q = mgr.get_queue("input")

  def block(v):
    return shmring.pack_rows(ring, [(np.full((64,), v, np.int32),) for _ in range(8)], timeout=0.5)

  q.put(block(1))
  q.put(block(2))
  feed = TFNode.DataFeed(mgr)
  a = feed.next_batch_arrays(8)[0]          # all of block 1 (view into slot 0)
  assert a.base is not None and int(a[0, 0]) == 1
  b = feed.next_batch_arrays(8)[0]          # block 2; block 1 is exhausted but still held
  with __import__("pytest").raises(RuntimeError):
    block(3)                                # both slots busy: the feeder has to wait
  assert int(a.sum()) == 8 * 64 and int(b[0, 0]) == 2
  q.put(None)
  assert feed.next_batch_arrays(8) == [] and feed.should_stop()   # 2 calls after block 1 ended
  assert block(4) is not None               # ... its slot is writable again
  mgr.shutdown()


def test_scalar_rows_travel_as_one_matrix():
  import numpy as np
  mgr = TFManager.start(b"abc", ["input", "output"], "local")
  name, ring = shmring.create(4, 1 << 20)
  mgr.set("ring", {"name": name, "nslots": 4, "slot_bytes": 1 << 20})
  rows = [[i] + list(range(10)) for i in range(20)]           # CSV-like: label, features
  blk = shmring.pack_rows(ring, rows)
  assert len(blk.layout) == 1 and blk.nrows == 20
  q = mgr.get_queue("input")
  q.put(blk)
  q.put(shmring.pack_rows(ring, [[0.5, 1.5], [2.5, 3.5]]))
  q.put(None)
  feed = TFNode.DataFeed(mgr)
  m = feed.next_batch_arrays(20)
  assert len(m) == 1 and m[0].shape == (20, 11) and m[0].dtype.kind == "i"
  assert m[0][:, 0].tolist() == list(range(20))
  assert feed.next_batch(5) == [[0.5, 1.5], [2.5, 3.5]]       # python-row view of the same path
  # mixed int/float rows keep their per-column types (no silent upcast of the label)
  mixed = shmring.pack_rows(ring, [[1, 0.5], [2, 1.5]])
  assert len(mixed.layout) == 2
  mgr.shutdown()
