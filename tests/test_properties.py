"""Property tests (hypothesis): the Example codec and the shared-memory ring round-trip arbitrary
well-formed inputs, and the native and pure-python codecs agree byte for byte."""
import numpy as np
import pytest

hyp = pytest.importorskip("hypothesis")
from hypothesis import given, settings, strategies as st  # noqa: E402

from tensorflowonspark_b200 import shmring, tfrecord  # noqa: E402

_names = st.text(alphabet="abcdefghijklmnopqrstuvwxyz_0123456789", min_size=1, max_size=12)
_feature = st.one_of(
    st.tuples(st.just("int64"), st.lists(st.integers(-2 ** 63, 2 ** 63 - 1), max_size=8)),
    st.tuples(st.just("float"), st.lists(st.floats(width=32, allow_nan=False, allow_infinity=False),
                                         max_size=8)),
    st.tuples(st.just("bytes"), st.lists(st.binary(max_size=16), max_size=4)),
)


@settings(max_examples=150, deadline=None)
@given(st.dictionaries(_names, _feature, max_size=6))
def test_example_codec_roundtrip_and_native_python_agreement(features):
  enc_py = tfrecord._py_encode(features)
  assert tfrecord.encode_example(features) == enc_py          # native (if built) == python
  for decode in (tfrecord.decode_example, tfrecord._py_decode):
    got = decode(enc_py)
    assert set(got) == set(features)
    for k, (kind, vals) in features.items():
      gk, gv = got[k]
      if not vals:
        assert list(gv) == []          # an empty list carries no type on the wire
        continue
      assert gk == kind
      if kind == "float":
        assert np.allclose(np.asarray(gv, np.float32), np.asarray(vals, np.float32), rtol=0, atol=0)
      else:
        assert list(gv) == list(vals)


@settings(max_examples=60, deadline=None)
@given(st.lists(st.binary(max_size=64), max_size=10))
def test_record_framing_roundtrip(tmp_path_factory, records):
  path = str(tmp_path_factory.mktemp("tfr") / "part-00000")
  tfrecord.write_records(path, records)
  assert list(tfrecord.read_records(path)) == records


_dtypes = st.sampled_from([np.uint8, np.int32, np.int64, np.float32, np.float64])


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 40), st.lists(st.tuples(_dtypes, st.lists(st.integers(1, 5), max_size=3)),
                                    min_size=1, max_size=3), st.integers(0, 2 ** 31 - 1))
def test_ring_pack_unpack_roundtrip(nrows, colspecs, seed):
  rng = np.random.RandomState(seed)
  rows = [tuple((rng.randint(0, 100, size=shape)).astype(dt) for dt, shape in colspecs)
          for _ in range(nrows)]
  name, ring = shmring.create(2, 1 << 20)
  try:
    blk = shmring.pack_rows(ring, rows, timeout=1.0)
    assert blk is not None and blk.nrows == nrows
    cols = shmring.unpack_columns(ring, blk)
    for j, (dt, shape) in enumerate(colspecs):
      assert cols[j].dtype == np.dtype(dt) and cols[j].shape == (nrows,) + tuple(shape)
      assert np.array_equal(cols[j], np.stack([r[j] for r in rows]))
    ring.release_read(blk.pos)
  finally:
    shmring._unlink(name)


@settings(max_examples=200, deadline=None)
@given(st.integers(1, 120), st.integers(1, 120), st.integers(1, 64),
       st.sampled_from([(1, False, 128, 1), (16, True, 128, 1), (16, True, 64, 64)]))
def test_choose_box_covers_the_tensor_within_the_mma_limits(OW, OH, N, mode):
  """The pixel-box planner of the implicit-GEMM kernels (pure python, no GPU): the box fits the
  MMA's 128 (or 64) rows, respects the K granularity of the weight-gradient kernels, and the
  tile grid covers every pixel."""
  from tensorflowonspark_b200.ops import igemm
  multiple_of, allow_pad, max_rows, min_rows = mode
  try:
    bw, bh, bn, tw, th, tn = igemm.choose_box(OW, OH, N, multiple_of=multiple_of,
                                              allow_pad=allow_pad, max_rows=max_rows,
                                              min_rows=min_rows)
  except ValueError:
    assert not allow_pad          # without padding some extents have no 16-row-multiple box
    return
  rows = bw * bh * bn
  assert 1 <= rows <= max_rows and rows % multiple_of == 0 and rows >= min_rows
  assert tw * bw >= OW and th * bh >= OH and tn * bn >= N          # full coverage
  assert (tw - 1) * bw < OW and (th - 1) * bh < OH and (tn - 1) * bn < N   # no empty tiles
  if not allow_pad:
    assert bw <= OW and bh <= OH
  if bn > 1:
    assert tw == 1 and th == 1    # images are batched into a box only when one box holds a whole image


def test_stem_geometry_matches_a_7x7_stride2_pad3_convolution():
  from tensorflowonspark_b200.ops import igemm
  for hw in (32, 64, 65, 224, 225):
    OH, OW, Wp = igemm.stem_geometry(hw, hw)
    assert OH == (hw + 6 - 7) // 2 + 1 == OW
    assert Wp % 2 == 0 and Wp >= hw + 4 and Wp >= 2 * (OW - 1) + 8   # last window stays in the row


_spec_entry = st.one_of(
    st.tuples(st.just("int64"), st.integers(1, 6), st.sampled_from(["int64", "int32", "uint8"])),
    st.tuples(st.just("float"), st.integers(1, 6), st.just("float32")),
    st.tuples(st.just("bytes"), st.integers(1, 9), st.just("uint8")),
)


@settings(max_examples=80, deadline=None)
@given(st.dictionaries(_names, _spec_entry, min_size=1, max_size=5), st.integers(1, 7),
       st.randoms(use_true_random=False))
def test_decode_batch_agrees_with_the_per_record_decoder_for_any_spec(spec, n, rnd):
  """tfrecord.decode_batch (native, one pass into dense arrays) == stacking decode_example."""
  recs, want = [], {k: [] for k in spec}
  for _ in range(n):
    feats = {}
    for name, (kind, length, dt) in spec.items():
      if kind == "int64":
        hi = {"int64": 2 ** 62, "int32": 2 ** 31 - 1, "uint8": 255}[dt]
        lo = {"int64": -2 ** 62, "int32": -2 ** 31, "uint8": 0}[dt]
        vals = [rnd.randint(lo, hi) for _ in range(length)]
        feats[name] = ("int64", vals)
      elif kind == "float":
        vals = [float(np.float32(rnd.uniform(-1e6, 1e6))) for _ in range(length)]
        feats[name] = ("float", vals)
      else:
        vals = bytes(rnd.randrange(256) for _ in range(length))
        feats[name] = ("bytes", [vals])
        vals = list(vals)
      want[name].append(vals)
    feats["extra_" + str(len(recs))] = ("int64", [1, 2, 3])        # features nobody asked for
    recs.append(tfrecord.encode_example(feats))
  out = tfrecord.decode_batch(recs, {k: v for k, v in spec.items()}, threads=rnd.choice([1, 2]))
  for name, (kind, length, dt) in spec.items():
    assert out[name].shape == (n, length) and out[name].dtype == np.dtype(dt)
    assert np.array_equal(out[name], np.asarray(want[name], dtype=dt))


@settings(max_examples=25, deadline=None)
@given(st.lists(st.tuples(st.sampled_from(["push", "sparse", "pull"]), st.integers(0, 2 ** 31)),
                min_size=1, max_size=12))
def test_tcp_parameter_server_matches_a_numpy_model_for_any_request_sequence(ops):
  """parallel/ps_net.py: whatever order pushes, row-sparse pushes and pulls arrive in, two server
  slices behind sockets hold what a flat numpy vector would."""
  from tensorflowonspark_b200 import reservation
  from tensorflowonspark_b200.parallel import ps_net
  srv = reservation.Server(1)
  addr = srv.start()
  cid = "prop-{}".format(abs(hash(tuple(ops))) % 10 ** 9)

  class Ctx(object):
    def __init__(self, job, idx):
      self.job_name, self.task_index = job, idx
      self.cluster_spec = {"ps": ["10.0.0.1:1", "10.0.0.2:2"], "worker": ["10.0.0.3:3"]}
      self.cluster_id, self.server_addr, self.gpus = cid, addr, []

  numel, width, base = 56, 5, 3
  servers = [ps_net.NetPSServer(Ctx("ps", i), numel) for i in range(2)]
  client = ps_net.NetPSClient(Ctx("worker", 0))
  model = np.zeros(numel, np.float32)
  try:
    for op, seed in ops:
      rng = np.random.RandomState(seed)
      if op == "push":
        g, lr = rng.randn(numel).astype(np.float32), float(rng.rand())
        client.push(g, lr=lr, scale=0.5)
        model -= np.float32(lr * 0.5) * g
      elif op == "sparse":
        rows = rng.randn(3, width).astype(np.float32)
        idx = rng.randint(0, (numel - base) // width, 3)
        client.push_sparse(rows, idx, width=width, base=base, lr=0.25)
        for row, r in zip(rows, idx):
          model[base + r * width:base + (r + 1) * width] -= np.float32(0.25) * row
      else:
        assert np.allclose(client.pull(), model, atol=1e-5)
    assert np.allclose(client.pull(), model, atol=1e-5)
    assert np.allclose(np.concatenate([s_.values() for s_ in servers]), model, atol=1e-5)
  finally:
    client.close()
    for s_ in servers:
      s_.close()
    srv.stop()
