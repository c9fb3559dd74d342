"""Summarise an `ncu --csv --metrics gpu__time_duration.sum` log: time per kernel family and the
slowest individual launches (with their grids)."""
import csv
import re
import sys


def main(path, top=40):
  rows = []
  with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
  for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
      continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = name.replace("tfos::(anonymous namespace)::", "").replace("tfos::", "")
    rows.append((int(r["ID"]), name, r["Grid Size"], r["Block Size"], us))
  total = sum(r[4] for r in rows)
  print("# {} launches, sum {:.2f} ms".format(len(rows), total / 1e3))
  fam = {}
  for _, n, _, _, us in rows:
    a = fam.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += us
  for n, (c, us) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
    print("{:52s} n={:4d} {:9.1f} us {:5.1f}%".format(n[:52], c, us, 100 * us / total))
  print("# slowest launches")
  for i, n, g, b, us in sorted(rows, key=lambda r: -r[4])[:top]:
    print("id={:5d} {:44s} grid={:18s} {:8.1f} us".format(i, n[:44], g, us))


if __name__ == "__main__":
  main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
