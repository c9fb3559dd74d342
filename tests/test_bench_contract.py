"""bench.py contract pieces that can be checked without a GPU: the reference arm's
"unavailable" line, the watchdog (a stalled section must not hang the caller and must not lose
the kernel-timed number), the supervisor's retry."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def test_reference_arm_reports_unavailable_and_exits_zero():
  p = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--gpus", "1", "--steps", "2",
                      "--warmup", "3"], capture_output=True, text=True, timeout=120)
  assert p.returncode == 0
  rec = json.loads(p.stdout.strip().splitlines()[-1])
  assert rec["impl"] == "reference" and isinstance(rec["unavailable"], str) and rec["unavailable"]


def _load_bench():
  import importlib.util
  spec = importlib.util.spec_from_file_location("bench_under_test", BENCH)
  mod = importlib.util.module_from_spec(spec)
  argv, sys.argv = sys.argv, ["bench.py"]
  try:
    spec.loader.exec_module(mod)
  finally:
    sys.argv = argv
  return mod


def test_watchdog_prints_partial_result_and_exits_cleanly():
  code = (
      "import sys, time, importlib.util\n"
      "sys.argv = ['bench.py']\n"
      "spec = importlib.util.spec_from_file_location('b', {!r})\n"
      "b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)\n"
      "w = b.Watchdog(0.3, 0)\n"
      "w.partial = {{'metric': 'm', 'value': 1.5}}\n"
      "time.sleep(20)\n"
      "print('NOT REACHED')\n").format(BENCH)
  p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=60)
  assert p.returncode == 0 and "NOT REACHED" not in p.stdout
  rec = json.loads(p.stdout.strip().splitlines()[-1])
  assert rec["value"] == 1.5 and "watchdog" in rec
  assert "thread stacks follow" in p.stderr


def test_watchdog_without_result_exits_nonzero():
  code = (
      "import sys, time, importlib.util\n"
      "sys.argv = ['bench.py']\n"
      "spec = importlib.util.spec_from_file_location('b', {!r})\n"
      "b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)\n"
      "w = b.Watchdog(0.3, 0)\n"
      "time.sleep(20)\n").format(BENCH)
  p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=60)
  assert p.returncode == 3 and p.stdout.strip() == ""


def test_supervisor_retries_once_then_gives_up_without_a_gpu():
  import torch
  if torch.cuda.is_available():
    import pytest
    pytest.skip("meaningful only where the benchmark itself cannot run")
  # (earlier tests of this session export RANK / WORLD_SIZE into os.environ: start clean)
  env = {k: v for k, v in os.environ.items()
         if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "TFOS_BENCH_CHILD", "TFOS_BENCH_SUPERVISE")}
  p = subprocess.run([sys.executable, BENCH, "--steps", "2", "--warmup", "3"], capture_output=True,
                     text=True, timeout=300, env=env)
  assert p.returncode == 3
  assert "attempt 1 failed" in p.stderr and "attempt 2 failed" in p.stderr
  assert not [l for l in p.stdout.splitlines() if l.startswith("{")]


def _supervise(child_code, watchdog_s="1"):
  env = {k: v for k, v in os.environ.items()
         if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "TFOS_BENCH_CHILD", "TFOS_BENCH_SUPERVISE")}
  env["TFOS_BENCH_CHILD_CMD"] = json.dumps([sys.executable, "-c", child_code])
  env["TFOS_BENCH_WATCHDOG_S"] = watchdog_s
  return subprocess.run([sys.executable, BENCH, "--steps", "2", "--warmup", "3"],
                        capture_output=True, text=True, timeout=400, env=env)


def test_supervisor_passes_the_childs_result_line_through():
  p = _supervise("print('NCCL version noise'); print('{\"metric\": \"m\", \"value\": 2.5, \"n_gpus\": 1}')")
  assert p.returncode == 0
  lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
  assert len(lines) == 1 and json.loads(lines[0]) == {"metric": "m", "value": 2.5, "n_gpus": 1}


def test_supervisor_retries_a_stalled_child(tmp_path):
  marker = str(tmp_path / "first_attempt_done")
  code = ("import os, sys, time\n"
          "m = {!r}\n"
          "if not os.path.exists(m):\n"
          "  open(m, 'w').close(); time.sleep(600)\n"
          "print('{{\"metric\": \"m\", \"value\": 7}}')\n").format(marker)
  p = _supervise(code, watchdog_s="-55")   # limit = watchdog + 60 = 5 s
  assert p.returncode == 0, p.stderr[-500:]
  rec = json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][-1])
  assert rec["value"] == 7 and rec["attempt"] == 2
  assert "attempt 1 failed (no result within" in p.stderr
