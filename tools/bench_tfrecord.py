"""Host-side throughput of the native TFRecord / Example codec (csrc/tfrecord.cc) on MNIST-shaped
records (784 int64 pixels + 1 label, ~1.2 KB each) - the per-worker input path of
InputMode.TENSORFLOW programs (utils/data.TFRecordPipeline).  CPU only; prints one JSON line."""
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402

from tensorflowonspark_b200 import tfrecord  # noqa: E402


def best(fn, reps=5):
  out = []
  for _ in range(reps):
    t = time.perf_counter()
    fn()
    out.append(time.perf_counter() - t)
  return min(out)


def main():
  n = int(os.environ.get("N", "4000"))
  rng = np.random.RandomState(0)
  img = rng.randint(0, 255, (n, 784))
  recs = [tfrecord.encode_example({"image": ("int64", img[i].tolist()), "label": ("int64", [i % 10])})
          for i in range(n)]
  path = os.path.join(tempfile.mkdtemp(), "part-00000")
  nbytes = sum(len(r) for r in recs)
  res = {"records": n, "bytes_per_record": nbytes // n}
  res["write_MB_s"] = round(nbytes / best(lambda: tfrecord.write_records(path, recs)) / 1e6)
  res["read_verify_crc_MB_s"] = round(nbytes / best(lambda: tfrecord.read_records(path)) / 1e6)
  slow = best(lambda: [tfrecord.decode_example(r) for r in recs], 3)
  res["decode_per_record_python_objects_rec_s"] = round(n / slow)
  spec = {"image": ("int64", 784, np.uint8), "label": ("int64", 1)}
  res["decode_batch_native_rec_s"] = round(n / best(lambda: tfrecord.decode_batch(recs, spec, threads=1)))
  res["decode_batch_speedup"] = round(res["decode_batch_native_rec_s"] * slow / n, 1)
  out = tfrecord.decode_batch(recs, spec)
  assert np.array_equal(out["image"], img.astype(np.uint8))
  print(json.dumps(res))


if __name__ == "__main__":
  main()
