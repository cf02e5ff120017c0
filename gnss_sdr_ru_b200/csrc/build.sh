#!/usr/bin/env bash
# Builds libgnssb200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libgnssb200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS="--expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O2,-Wall -Xptxas -v -cudart static ${EXTRA_NVCC_FLAGS:-}"
mkdir -p "$HERE/build"
for f in track api acq synth softtrack navbits ingest gpssdr_acq; do
  if [ ! -f "$HERE/build/$f.o" ] || [ "$HERE/$f.cu" -nt "$HERE/build/$f.o" ] || [ -n "$(find "$HERE" "$HERE/../../include" -maxdepth 1 \( -name '*.cuh' -o -name '*.h' \) -newer "$HERE/build/$f.o" 2>/dev/null)" ]; then
    $NVCC $FLAGS -c "$HERE/$f.cu" -o "$HERE/build/$f.o" 2> "$HERE/build/$f.ptxas.log" || { cat "$HERE/build/$f.ptxas.log"; exit 1; }
  fi
done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o "$OUT" "$HERE/build/track.o" "$HERE/build/api.o" "$HERE/build/acq.o" "$HERE/build/synth.o" "$HERE/build/softtrack.o" "$HERE/build/navbits.o" "$HERE/build/ingest.o" "$HERE/build/gpssdr_acq.o"
echo "built $OUT"
