#!/bin/sh
# Counterpart of the reference's tests/run_tests.sh (which starts a Spark Standalone master + 2
# workers first): the bundled sparklite engine forks its 2 executor processes per test module, so
# there is nothing to start.  CPU tier always; GPU tier when a CUDA device is visible.
set -e
cd "$(dirname "$0")/.."
python -m pytest tests -x -q -m "not gpu" "$@"
if python -c "import torch,sys; sys.exit(0 if torch.cuda.is_available() else 1)"; then
  python -m pytest tests -x -q -m gpu "$@"
  python tools/gpu_check.py
fi
