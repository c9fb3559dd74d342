#!/bin/bash
# AddressSanitizer + UBSan fuzz of the native TFRecord / Example decoders (csrc/tfrecord.cc):
# builds the codec as a stand-alone pybind module with -fsanitize=address,undefined and feeds it
# mutated (bit-flipped / truncated / spliced) serialized Examples through example_decode and
# example_decode_batch.  Every input must decode or raise - never read out of bounds.
#   tools/fuzz_tfrecord.sh [iterations]          (host only, ~1 min per 60k iterations)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
W=$(mktemp -d)
PYINC=$(python -c "import sysconfig;print(sysconfig.get_paths()['include'])")
PB=$(python -c "import torch,os;print(os.path.join(os.path.dirname(torch.__file__),'include'))")
NPI=$(python -c "import numpy;print(numpy.get_include())")
SUF=$(python -c "import sysconfig;print(sysconfig.get_config_var('EXT_SUFFIX'))")
cat > $W/mod.cc <<'EOC'
#include "tfrecord.h"
PYBIND11_MODULE(tfr_asan, m) { tfos::bind_tfrecord(m); }
EOC
g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -shared -fPIC \
    -I$PYINC -I$PB -I$NPI -I$ROOT/csrc $W/mod.cc $ROOT/csrc/tfrecord.cc -o $W/tfr_asan$SUF
cat > $W/fuzz.py <<'EOP'
import random, sys
sys.path.insert(0, sys.argv[1])
import tfr_asan as n
iters = int(sys.argv[2])
rng = random.Random(1)
base = [n.example_encode({"image": ("int64", [rng.randrange(0, 70000) for _ in range(50)]),
                          "f": ("float", [1.0, 2.0, 3.0]), "raw": ("bytes", [bytes(range(16))]),
                          "label": ("int64", [3])}) for _ in range(8)]
specs = [[("image", "int64", 50, "int32"), ("f", "float", 3, "float32"), ("raw", "bytes", 16, "uint8"),
          ("label", "int64", 1, "int64")], [("image", "int64", 50, "uint8"), ("label", "int64", 1, "int64")]]
ok = err = 0
for it in range(iters):
  recs = list(base)
  i = rng.randrange(len(recs))
  b = bytearray(recs[i])
  m = rng.random()
  if m < 0.4:
    for _ in range(rng.randrange(1, 4)):
      b[rng.randrange(len(b))] = rng.randrange(256)
  elif m < 0.7:
    del b[rng.randrange(len(b)):]
  elif m < 0.85:
    p = rng.randrange(len(b))
    b[p:p] = bytes(rng.randrange(256) for _ in range(rng.randrange(1, 9)))
  else:
    p = rng.randrange(len(b))
    del b[p:p + rng.randrange(1, 6)]
  recs[i] = bytes(b)
  for s in specs:
    try:
      n.example_decode_batch(recs, s, 1 if it % 7 else 3)
      ok += 1
    except RuntimeError:
      err += 1
  try:
    n.example_decode(recs[i])
  except (RuntimeError, ValueError, UnicodeDecodeError):
    pass
print("asan+ubsan fuzz: {} iterations, {} batches decoded, {} rejected, no sanitizer report".format(iters, ok, err))
EOP
ASAN_OPTIONS=detect_leaks=0 LD_PRELOAD=$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so) \
    python $W/fuzz.py $W ${1:-60000}
rm -rf $W
