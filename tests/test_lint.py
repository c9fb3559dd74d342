"""Static hygiene that needs no external linter (none is installable offline): every Python file
parses, has no tab indentation and respects the line limit of tox.ini; when pycodestyle /
pyflakes happen to be installed they are run with the repo's configuration too."""
import ast
import importlib.util
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SKIP = ("_ext", "baseline/_ref", ".git", "gpurun_out", "__pycache__", "docs")


def _sources():
  for base, dirs, files in os.walk(ROOT):
    rel = os.path.relpath(base, ROOT)
    if any(rel == s or rel.startswith(s + os.sep) for s in SKIP):
      dirs[:] = []
      continue
    for f in files:
      if f.endswith(".py"):
        yield os.path.join(base, f)


def test_every_source_parses_and_respects_the_line_limit():
  bad = []
  for path in _sources():
    with open(path, encoding="utf-8") as f:
      text = f.read()
    try:
      ast.parse(text, path)
    except SyntaxError as e:
      bad.append("{}: {}".format(path, e))
      continue
    for i, line in enumerate(text.splitlines(), 1):
      if "\t" in line[:len(line) - len(line.lstrip())]:
        bad.append("{}:{}: tab indentation".format(path, i))
      if len(line) > 120:
        bad.append("{}:{}: {} columns".format(path, i, len(line)))
  assert not bad, "\n".join(bad[:20])


@pytest.mark.skipif(importlib.util.find_spec("pycodestyle") is None, reason="pycodestyle not installed")
def test_pycodestyle_with_repo_config():
  p = subprocess.run([sys.executable, "-m", "pycodestyle", "--config", os.path.join(ROOT, "tox.ini"),
                      os.path.join(ROOT, "tensorflowonspark_b200")], capture_output=True, text=True)
  assert p.returncode == 0, p.stdout[-3000:]
