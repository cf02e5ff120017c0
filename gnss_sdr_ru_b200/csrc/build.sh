#!/usr/bin/env bash
# Builds libgnssb200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
# GNSSB200_OUT / GNSSB200_BUILD_DIR: an A/B build beside the product library (with EXTRA_NVCC_FLAGS), loaded through GNSSB200_LIB
OUT="${GNSSB200_OUT:-$HERE/../libgnssb200.so}"
BLD="${GNSSB200_BUILD_DIR:-$HERE/build}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS="--expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O2,-Wall -Xptxas -v -cudart static ${EXTRA_NVCC_FLAGS:-}"
mkdir -p "$BLD"
for f in track api acq synth softtrack navbits ingest gpssdr_acq; do
  if [ ! -f "$BLD/$f.o" ] || [ "$HERE/$f.cu" -nt "$BLD/$f.o" ] || [ -n "$(find "$HERE" "$HERE/../../include" -maxdepth 1 \( -name '*.cuh' -o -name '*.h' \) -newer "$BLD/$f.o" 2>/dev/null)" ]; then
    $NVCC $FLAGS -c "$HERE/$f.cu" -o "$BLD/$f.o" 2> "$BLD/$f.ptxas.log" || { cat "$BLD/$f.ptxas.log"; exit 1; }
  fi
done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o "$OUT" "$BLD/track.o" "$BLD/api.o" "$BLD/acq.o" "$BLD/synth.o" "$BLD/softtrack.o" "$BLD/navbits.o" "$BLD/ingest.o" "$BLD/gpssdr_acq.o"
echo "built $OUT"
