#!/bin/sh
# Compile-check (ptxas validates every inline-PTX form for sm_100a) the drafts under
# csrc/experimental/ - they are not part of the shipped extension and are never launched.
set -e
cd "$(dirname "$0")/.."
TORCH_INC=$(python -c "import torch, os; print(os.path.join(os.path.dirname(torch.__file__), 'include'))")
for f in csrc/experimental/*.cu; do
  echo "== $f"
  nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 --expt-relaxed-constexpr -O3 -lineinfo \
       -Icsrc -I"$TORCH_INC" -Xptxas -v -c "$f" -o /tmp/$(basename "$f").o 2>&1 | grep -v "^$" | tail -6
done
