"""Segmentation U-Net fed through InputMode.SPARK: rows travel RDD -> feeder task -> shared-memory
ring -> DataFeed -> pinned staging -> cudaMemcpyAsync (copy stream) -> native U-Net step
(BASELINE.json config "segmentation U-Net InputMode.SPARK DataFeed queue").

  python bench/unet_datafeed.py --gpus 1 --batch 64 --examples 16384
"""
import json
import os
import re
import subprocess
import sys

if __name__ == "__main__":
  import argparse
  p = argparse.ArgumentParser()
  p.add_argument("--gpus", type=int, default=1)
  p.add_argument("--batch", type=int, default=64)
  p.add_argument("--examples", type=int, default=16384)
  a = p.parse_args()
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  cmd = [sys.executable, os.path.join(root, "examples", "segmentation", "segmentation_spark.py"),
         "--cluster_size", str(a.gpus), "--batch_size", str(a.batch), "--input_mode", "spark",
         "--num_examples", str(a.examples), "--epochs", "1"]
  out = subprocess.run(cmd, capture_output=True, text=True, timeout=1200)
  rates = [float(m.group(1)) for m in re.finditer(r"(\d+) images/s", out.stdout)]
  row_bytes = 128 * 128 * 3 + 128 * 128
  if not rates:
    print(out.stdout[-2000:], out.stderr[-3000:])
    sys.exit(1)
  print(json.dumps({"metric": "U-Net training images/s through InputMode.SPARK DataFeed",
                    "value": rates[-1], "unit": "images/s", "n_gpus": a.gpus,
                    "feed_MB_per_s": rates[-1] * row_bytes / 1e6, "batch": a.batch}))
