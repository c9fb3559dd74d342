"""TensorBoard event files written without TensorFlow (utils/summary.py) - what the TensorBoard
subprocess of TFSparkNode (reference TFSparkNode.py:293-329) reads.  The wire format is checked
against the real ``Event`` / ``Summary`` / ``HistogramProto`` schema, declared here through the
protobuf runtime, so the files are what TensorBoard's own loader parses."""
import os

import numpy as np
import pytest

from tensorflowonspark_b200 import tfrecord
from tensorflowonspark_b200.utils import summary


def _event_class():
  pb = pytest.importorskip("google.protobuf")
  from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
  T = descriptor_pb2.FieldDescriptorProto
  fd = descriptor_pb2.FileDescriptorProto(name="tfos_event_test.proto", package="tfos_test", syntax="proto3")

  def msg(name, fields):
    m = fd.message_type.add(name=name)
    for fname, num, ftype, label, tname in fields:
      f = m.field.add(name=fname, number=num, type=ftype, label=label)
      if tname:
        f.type_name = ".tfos_test." + tname
  OPT, REP = T.LABEL_OPTIONAL, T.LABEL_REPEATED
  msg("HistogramProto", [("min", 1, T.TYPE_DOUBLE, OPT, None), ("max", 2, T.TYPE_DOUBLE, OPT, None),
                         ("num", 3, T.TYPE_DOUBLE, OPT, None), ("sum", 4, T.TYPE_DOUBLE, OPT, None),
                         ("sum_squares", 5, T.TYPE_DOUBLE, OPT, None),
                         ("bucket_limit", 6, T.TYPE_DOUBLE, REP, None), ("bucket", 7, T.TYPE_DOUBLE, REP, None)])
  msg("Value", [("tag", 1, T.TYPE_STRING, OPT, None), ("simple_value", 2, T.TYPE_FLOAT, OPT, None),
                ("histo", 5, T.TYPE_MESSAGE, OPT, "HistogramProto")])
  msg("Summary", [("value", 1, T.TYPE_MESSAGE, REP, "Value")])
  msg("Event", [("wall_time", 1, T.TYPE_DOUBLE, OPT, None), ("step", 2, T.TYPE_INT64, OPT, None),
                ("file_version", 3, T.TYPE_STRING, OPT, None), ("summary", 5, T.TYPE_MESSAGE, OPT, "Summary")])
  pool = descriptor_pool.DescriptorPool()
  pool.Add(fd)
  del pb
  return message_factory.GetMessageClass(pool.FindMessageTypeByName("tfos_test.Event"))


def test_event_file_parses_with_the_real_schema(tmp_path):
  Event = _event_class()
  w = summary.SummaryWriter(str(tmp_path), flush_secs=1e9, max_queue=1000)
  for step in range(5):
    w.add_scalar("loss", 2.0 / (step + 1), step)
  w.add_scalars({"images_per_s": 13428.0, "lr": 0.1}, 5, wall_time=123.5)
  vals = np.random.default_rng(0).normal(0, 0.02, 4096)
  w.add_histogram("fc/weights", vals, 5)
  w.close()
  files = summary.event_files(str(tmp_path))
  assert len(files) == 1 and os.path.basename(files[0]).startswith("events.out.tfevents.")
  events = []
  for rec in tfrecord.read_records(files[0]):           # CRCs verified by the reader
    e = Event()
    e.ParseFromString(rec)
    events.append(e)
  assert events[0].file_version == "brain.Event:2" and events[0].wall_time > 1e9
  assert [e.step for e in events[1:6]] == list(range(5))
  assert events[3].summary.value[0].tag == "loss"
  assert abs(events[3].summary.value[0].simple_value - 2.0 / 3) < 1e-6
  multi = events[6]
  assert multi.step == 5 and multi.wall_time == 123.5
  assert {v.tag: v.simple_value for v in multi.summary.value} == {"images_per_s": 13428.0,
                                                                   "lr": pytest.approx(0.1, rel=1e-6)}
  h = events[7].summary.value[0]
  assert h.tag == "fc/weights" and h.histo.num == 4096
  assert h.histo.min == vals.min() and h.histo.max == vals.max()
  assert abs(h.histo.sum - vals.sum()) < 1e-9 and abs(h.histo.sum_squares - (vals * vals).sum()) < 1e-9
  assert len(h.histo.bucket) == len(h.histo.bucket_limit) and sum(h.histo.bucket) == 4096
  limits = list(h.histo.bucket_limit)
  assert limits == sorted(limits) and limits[-1] >= vals.max() and limits[0] < vals.min()
  # every bucket count matches a direct count over (previous limit, limit]
  prev = -np.inf
  for lim, cnt in zip(limits, h.histo.bucket):
    assert cnt == np.count_nonzero((vals > prev) & (vals <= lim))
    prev = lim


def test_reader_round_trip_buffering_and_append(tmp_path):
  w = summary.SummaryWriter(str(tmp_path), max_queue=4, flush_secs=1e9)
  for step in range(3):
    w.add_scalar("a", step, step)
  assert len(summary.read_events(w.path)) == 1          # header only: three events still buffered
  w.add_scalar("a", 3, 3)                               # fourth event reaches max_queue
  assert len(summary.read_events(w.path)) == 5
  w.add_histogram("h", [0.0, 0.0, 1.0, -1.0, float("nan")], 4)
  w.close()
  with pytest.raises(ValueError):
    w.add_scalar("a", 1, 1)
  ev = summary.read_events(w.path)
  assert ev[0]["file_version"] == "brain.Event:2"
  assert [e["scalars"].get("a") for e in ev[1:5]] == [0.0, 1.0, 2.0, 3.0]
  h = ev[5]["histograms"]["h"]
  assert h["num"] == 4 and h["min"] == -1.0 and h["max"] == 1.0 and sum(h["bucket"]) == 4
  # two writers in one directory never share a file
  w2 = summary.SummaryWriter(str(tmp_path))
  w2.close()
  assert len(summary.event_files(str(tmp_path))) == 2


def test_torch_tensor_histogram_and_context_manager(tmp_path):
  import torch
  with summary.SummaryWriter(str(tmp_path / "sub" / "dir")) as w:
    w.add_histogram("w", torch.linspace(-1, 1, 101, dtype=torch.bfloat16), 0)
  (path,) = summary.event_files(str(tmp_path / "sub" / "dir"))
  assert summary.read_events(path)[1]["histograms"]["w"]["num"] == 101
