"""TFRecord framing / Example codec (native + pure-Python twins, protobuf as oracle) and the
DataFrame round-trip (scenarios of reference tests/test_dfutil.py:30-73, DFUtilTest.scala:29-132,
SimpleTypeParserTest.scala:9-15)."""
import struct

import pytest

from tensorflowonspark_b200 import _build, dfutil, tfrecord


def test_crc32c_vectors():
  # RFC 3720 test vectors
  assert tfrecord.crc32c(b"123456789") == 0xe3069283
  assert tfrecord.crc32c(b"\x00" * 32) == 0x8a9136aa
  assert tfrecord.crc32c(b"\xff" * 32) == 0x62a8ab43
  assert tfrecord._py_crc32c(bytes(range(32))) == 0x46dd794e == tfrecord.crc32c(bytes(range(32)))


def test_framing_and_corruption(tmp_path):
  p = str(tmp_path / "x.tfrecord")
  recs = [b"", b"a", b"hello world" * 100, bytes(range(256))]
  tfrecord.write_records(p, recs)
  assert [bytes(r) for r in tfrecord.read_records(p)] == recs
  raw = open(p, "rb").read()
  assert struct.unpack("<Q", raw[:8])[0] == 0
  bad = bytearray(raw)
  bad[-10] ^= 0xff
  open(p, "wb").write(bytes(bad))
  with pytest.raises(Exception, match="CRC"):
    tfrecord.read_records(p)


FEATS = {"i": ("int64", [1, -2, 1 << 40]), "f": ("float", [1.5, -2.25]), "b": ("bytes", [b"xy", b""]),
         "e": ("int64", [])}


def test_example_python_and_native_agree():
  py = tfrecord._py_encode(FEATS)
  assert tfrecord._py_decode(py) == {k: (v[0], list(v[1])) for k, v in FEATS.items()}
  if _build.available():
    nat = tfrecord.encode_example(FEATS)
    assert tfrecord._py_decode(nat) == tfrecord._py_decode(py)
    assert dict(tfrecord.decode_example(py)) == tfrecord._py_decode(py)


def test_example_matches_protobuf_wire_format():
  pb = pytest.importorskip("google.protobuf.descriptor_pb2")
  from google.protobuf import descriptor_pool, message_factory
  fdp = pb.FileDescriptorProto(name="ex.proto", package="t", syntax="proto3")

  def msg(name, fields):
    m = fdp.message_type.add(name=name)
    for fname, num, typ, label, tname in fields:
      f = m.field.add(name=fname, number=num, type=typ, label=label)
      if tname:
        f.type_name = tname
    return m

  T, L = pb.FieldDescriptorProto, pb.FieldDescriptorProto
  msg("BytesList", [("value", 1, T.TYPE_BYTES, L.LABEL_REPEATED, None)])
  msg("FloatList", [("value", 1, T.TYPE_FLOAT, L.LABEL_REPEATED, None)])
  msg("Int64List", [("value", 1, T.TYPE_INT64, L.LABEL_REPEATED, None)])
  msg("Feature", [("bytes_list", 1, T.TYPE_MESSAGE, L.LABEL_OPTIONAL, ".t.BytesList"),
                  ("float_list", 2, T.TYPE_MESSAGE, L.LABEL_OPTIONAL, ".t.FloatList"),
                  ("int64_list", 3, T.TYPE_MESSAGE, L.LABEL_OPTIONAL, ".t.Int64List")])
  ent = msg("Entry", [("key", 1, T.TYPE_STRING, L.LABEL_OPTIONAL, None),
                      ("value", 2, T.TYPE_MESSAGE, L.LABEL_OPTIONAL, ".t.Feature")])
  msg("Features", [("feature", 1, T.TYPE_MESSAGE, L.LABEL_REPEATED, ".t.Entry")])
  msg("Example", [("features", 1, T.TYPE_MESSAGE, L.LABEL_OPTIONAL, ".t.Features")])
  del ent
  pool = descriptor_pool.DescriptorPool()
  pool.Add(fdp)
  Example = message_factory.GetMessageClass(pool.FindMessageTypeByName("t.Example"))
  ex = Example()
  ex.ParseFromString(tfrecord.encode_example(FEATS))
  got = {e.key: e.value for e in ex.features.feature}
  assert list(got["i"].int64_list.value) == [1, -2, 1 << 40]
  assert list(got["f"].float_list.value) == [1.5, -2.25]
  assert list(got["b"].bytes_list.value) == [b"xy", b""]
  # and the other direction: protobuf-serialised bytes decode with our codec
  assert dict(tfrecord.decode_example(ex.SerializeToString()))["i"] == ("int64", [1, -2, 1 << 40])


def test_parse_schema():
  s = dfutil.parse_schema(
      "struct<a:binary,b:boolean,c:int,d:long,e:bigint,f:float,g:double,h:string,i:array<float>>")
  assert s.simpleString() == ("struct<a:binary,b:boolean,c:int,d:bigint,e:bigint,f:float,g:double,"
                              "h:string,i:array<float>>")
  with pytest.raises(ValueError):
    dfutil.parse_schema("struct<a:decimal>")


def test_dataframe_roundtrip(sc, spark, tmp_path):
  rows = [("r%d" % i, i, [i, i + 1], i / 3.0, [float(i), 2.5], bytearray(b"\x00\x01" + bytes([i])))
          for i in range(10)]
  df = spark.createDataFrame(rows, ["a", "b", "c", "d", "e", "f"])
  out = str(tmp_path / "tfr")
  dfutil.saveAsTFRecords(df, out)
  df2 = dfutil.loadTFRecords(sc, out, binary_features=["f"])
  assert dfutil.isLoadedDF(df2) and not dfutil.isLoadedDF(df)
  assert not dfutil.isLoadedDF(df2.select("a"))
  assert df2.dtypes == df.dtypes
  got = sorted(df2.collect(), key=lambda r: r.b)
  for r, (a, b, c, d, e, f) in zip(got, rows):
    assert (r.a, r.b, r.c, bytes(r.f)) == (a, b, c, bytes(f))
    assert abs(r.d - d) < 1e-6 and [round(x, 6) for x in r.e] == e


def test_schema_hint_overrides_inference(sc, spark, tmp_path):
  df = spark.createDataFrame([(1, [7]), (2, [8])], ["k", "single"])
  out = str(tmp_path / "tfr2")
  dfutil.saveAsTFRecords(df, out)
  assert dict(dfutil.loadTFRecords(sc, out).dtypes)["single"] == "bigint"  # one value looks scalar
  hinted = dfutil.loadTFRecords(sc, out, schema_hint="struct<single:array<int>>")
  assert dict(hinted.dtypes)["single"] == "array<int>"
  assert sorted(r.single for r in hinted.collect()) == [[7], [8]]


def test_tfrecord_pipeline_interleave_shuffle_shard_batch(tmp_path):
  """utils/data.TFRecordPipeline (the tf.data chain of reference mnist_tf_ds.py:41-50): every
  record is seen once per epoch, shards are disjoint and complete (by file and, with fewer files
  than shards, by record), batches have the requested shape."""
  import numpy as np
  from tensorflowonspark_b200 import tfrecord
  from tensorflowonspark_b200.utils import data
  ids = list(range(103))
  for f in range(4):
    recs = [tfrecord.encode_example({"id": ("int64", [i]), "x": ("float", [i * 0.5, 1.0])})
            for i in ids[f::4]]
    tfrecord.write_records(str(tmp_path / "part-r-{:05d}".format(f)), recs)

  def parse(rec):
    ex = tfrecord.decode_example(rec)
    return np.int64(ex["id"][1][0]), np.asarray(ex["x"][1], dtype=np.float32)

  pat = str(tmp_path / "part-*")
  assert len(data.list_files(pat)) == 4 and data.list_files(str(tmp_path)) == data.list_files(pat)
  seen = [int(i) for i, _ in data.TFRecordPipeline(pat, epochs=2, shuffle_buffer=16, seed=1).map(parse)]
  assert sorted(seen) == sorted(ids * 2) and seen[:20] != sorted(seen)[:20]
  for world in (2, 8):     # 8 > number of files: sharding falls back to records
    parts = [sorted(int(i) for i, _ in data.TFRecordPipeline(pat, seed=0).shard(world, r).map(parse))
             for r in range(world)]
    assert sorted(sum(parts, [])) == ids and all(parts)
  batches = list(data.TFRecordPipeline(pat, shuffle_buffer=8).map(parse).batch(10))
  assert len(batches) == 10 and batches[0][0].shape == (10,) and batches[0][1].shape == (10, 2)
  tail = list(data.TFRecordPipeline(pat).map(parse).batch(10, drop_remainder=False))
  assert len(tail) == 11 and tail[-1][0].shape == (3,)


def _batch_examples(n=300, seed=0):
  import numpy as np
  rng = np.random.RandomState(seed)
  img = rng.randint(0, 255, (n, 784))
  big = rng.randint(-2 ** 62, 2 ** 62, (n, 3))            # ten-byte varints, negative values
  big[0] = [-1, -2 ** 63, 2 ** 63 - 1]
  big[1] = [127, 128, 16384]                               # the one / two / three byte boundaries
  flt = rng.randn(n, 5).astype(np.float32)
  raw = rng.randint(0, 256, (n, 12)).astype(np.uint8)
  recs = [tfrecord.encode_example({
      "image": ("int64", img[i].tolist()), "id": ("int64", big[i].tolist()),
      "f": ("float", flt[i].tolist()), "raw": ("bytes", [raw[i].tobytes()]),
      "unused": ("bytes", [b"skip me"])}) for i in range(n)]
  return recs, img, big, flt, raw


def test_decode_batch_writes_dense_arrays_and_agrees_with_the_python_twin():
  import numpy as np
  recs, img, big, flt, raw = _batch_examples()
  spec = {"image": ("int64", 784, np.uint8), "id": ("int64", 3), "f": ("float", 5), "raw": ("bytes", 12)}
  for threads in (1, 3):
    out = tfrecord.decode_batch(recs, spec, threads=threads)
    assert [out[k].dtype.name for k in spec] == ["uint8", "int64", "float32", "uint8"]
    assert np.array_equal(out["image"], img.astype(np.uint8)) and np.array_equal(out["id"], big)
    assert np.array_equal(out["f"], flt) and np.array_equal(out["raw"], raw)
  as32 = tfrecord.decode_batch(recs[:7], {"image": ("int64", 784, "int32")})["image"]
  assert as32.dtype == np.int32 and np.array_equal(as32, img[:7])
  native, tfrecord._native = tfrecord._native, lambda: None      # the pure-python twin
  try:
    twin = tfrecord.decode_batch(recs[:20], spec)
  finally:
    tfrecord._native = native
  full = tfrecord.decode_batch(recs[:20], spec)
  assert all(np.array_equal(twin[k], full[k]) for k in spec)
  assert tfrecord.decode_batch([], spec)["image"].shape == (0, 784)


def test_decode_batch_rejects_records_that_do_not_match_the_spec():
  recs, _, _, _, _ = _batch_examples(4)
  for bad, msg in (({"image": ("int64", 100)}, "more values"), ({"image": ("int64", 785)}, "784 values"),
                   ({"nope": ("int64", 1)}, "no values"), ({"f": ("int64", 5)}, "another type"),
                   ({"raw": ("bytes", 11)}, "requested length")):
    with pytest.raises(RuntimeError, match=msg):
      tfrecord.decode_batch(recs, bad)
  with pytest.raises(RuntimeError):
    tfrecord.decode_batch([recs[0][:40]], {"image": ("int64", 784)})     # truncated record


def test_pipeline_decode_stage_yields_the_same_batches_as_map(tmp_path):
  import numpy as np
  from tensorflowonspark_b200.utils import data
  recs, img, _, _, _ = _batch_examples(100)
  labelled = [tfrecord.encode_example({"image": ("int64", img[i].tolist()), "label": ("int64", [i % 10])})
              for i in range(100)]
  for k in range(2):
    tfrecord.write_records(str(tmp_path / "part-{:05d}".format(k)), labelled[k * 50:(k + 1) * 50])

  def parse(rec):
    ex = tfrecord.decode_example(rec)
    return np.asarray(ex["image"][1], dtype=np.uint8), np.asarray(ex["label"][1], dtype=np.int64)

  def make():
    return data.TFRecordPipeline(str(tmp_path), epochs=2, shuffle_buffer=16, seed=3)
  slow = list(make().map(parse).batch(32, drop_remainder=False))
  fast = list(make().decode({"image": ("int64", 784, np.uint8), "label": ("int64", 1)}).batch(32, drop_remainder=False))
  assert len(slow) == len(fast) == 7 and fast[-1][0].shape == (8, 784)
  for (a, b), (c, d) in zip(slow, fast):
    assert np.array_equal(a, c) and np.array_equal(b, d)
  with pytest.raises(ValueError):
    list(make().decode({"label": ("int64", 1)}))                        # decode needs batch


def test_mutated_records_decode_or_raise_never_crash():
  """Untrusted bytes reach native parsers (csrc/tfrecord.cc): bit flips, truncation and splices
  must end in a decoded value or a RuntimeError (tools/fuzz_tfrecord.sh runs the same loop under
  AddressSanitizer + UBSan)."""
  import random
  rng = random.Random(7)
  base = [tfrecord.encode_example({"image": ("int64", [rng.randrange(0, 70000) for _ in range(50)]),
                                   "f": ("float", [1.0, 2.0, 3.0]), "raw": ("bytes", [bytes(range(16))]),
                                   "label": ("int64", [3])}) for _ in range(4)]
  spec = {"image": ("int64", 50, "int32"), "f": ("float", 3), "raw": ("bytes", 16), "label": ("int64", 1)}
  decoded = rejected = 0
  for it in range(3000):
    recs = list(base)
    i = rng.randrange(len(recs))
    b = bytearray(recs[i])
    m = rng.random()
    if m < 0.5:
      for _ in range(rng.randrange(1, 4)):
        b[rng.randrange(len(b))] = rng.randrange(256)
    elif m < 0.8:
      del b[rng.randrange(len(b)):]
    else:
      p = rng.randrange(len(b))
      b[p:p] = bytes(rng.randrange(256) for _ in range(rng.randrange(1, 9)))
    recs[i] = bytes(b)
    try:
      out = tfrecord.decode_batch(recs, spec, threads=1)
      assert out["image"].shape == (4, 50)
      decoded += 1
    except RuntimeError:
      rejected += 1
    try:
      tfrecord.decode_example(recs[i])
    except (RuntimeError, ValueError, UnicodeDecodeError):
      pass
  assert decoded > 0 and rejected > 0


def test_tfrecord_files_and_input_pipeline_on_a_remote_filesystem():
  """InputMode.TENSORFLOW workers read their TFRecord shards from wherever the data lives - the
  reference gets ``hdfs://`` from TensorFlow's filesystem layer (examples/mnist/keras/
  mnist_tf_ds.py:41-50 with ``ctx.absolute_path``).  pyarrow's in-memory filesystem stands in."""
  pytest.importorskip("pyarrow")
  import numpy as np
  from tensorflowonspark_b200.utils import data
  base = "mock:///datasets/mnist/train"
  recs = [tfrecord.encode_example({"image": ("int64", [i % 256] * 16), "label": ("int64", [i % 10])})
          for i in range(60)]
  for k in range(3):
    tfrecord.write_records("{}/part-{:05d}".format(base, k), recs[k * 20:(k + 1) * 20])
  tfrecord.write_records(base + "/_SUCCESS", [])
  assert tfrecord.read_records(base + "/part-00001") == recs[20:40]
  assert data.list_files(base) == [base + "/part-0000{}".format(k) for k in range(3)]
  assert data.list_files(base + "/part-*1") == [base + "/part-00001"]
  with pytest.raises(ValueError):
    tfrecord.write_records(base + "/part-00000", recs[:1], append=True)
  spec = {"image": ("int64", 16, np.uint8), "label": ("int64", 1)}
  batches = list(data.TFRecordPipeline(base, epochs=1).shard(3, 1).decode(spec).batch(10))
  assert len(batches) == 2                                   # one of the three files, 20 records
  labels = np.concatenate([b[1].reshape(-1) for b in batches])
  assert sorted(labels.tolist()) == sorted(i % 10 for i in range(20, 40))
  # a flipped payload byte is caught by the CRC check on the remote path too
  from tensorflowonspark_b200.utils import fs
  with fs.open_read(base + "/part-00000") as f:
    blob = bytearray(f.read())
  blob[20] ^= 0xff
  fs.write_atomic(base + "/corrupt", lambda f: f.write(bytes(blob)))
  with pytest.raises(IOError, match="CRC"):
    tfrecord.read_records(base + "/corrupt")
