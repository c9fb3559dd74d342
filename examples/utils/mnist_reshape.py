"""Turn one CSV row (label,pix0..pix783) into the JSON a serving tool expects
(reference: examples/utils/mnist_reshape.py).   usage: mnist_reshape.py '<csv row>'"""
import json
import sys

if __name__ == "__main__":
  v = [int(x) for x in sys.argv[1].split(",")]
  print(json.dumps({"label": v[0], "image": [v[1 + 28 * r:1 + 28 * (r + 1)] for r in range(28)]}))
