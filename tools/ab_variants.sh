#!/usr/bin/env bash
# A/B timing of library builds on the GPU box: every library named on the command line (default: the product
# library and everything under gnss_sdr_ru_b200/variants/) runs the same tracking shapes through tools/form_sweep.py
# (GNSSB200_LIB selects the build; the digest test keeps a faster build honest).
#   build a variant here (no GPU needed):
#     GNSSB200_OUT=$PWD/gnss_sdr_ru_b200/variants/libgnssb200_x.so GNSSB200_BUILD_DIR=/tmp/build_x \
#       EXTRA_NVCC_FLAGS="-DSEG_U2=0" bash gnss_sdr_ru_b200/csrc/build.sh
#   run there:  bash tools/ab_variants.sh [streams] [lib ...]      e.g. streams = 64,8
set -uo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
STREAMS="${1:-64,8}"
shift || true
LIBS=("$@")
if [ ${#LIBS[@]} -eq 0 ]; then
  LIBS=("$ROOT/gnss_sdr_ru_b200/libgnssb200.so" "$ROOT"/gnss_sdr_ru_b200/variants/*.so)
fi
for rep in 1 2; do
  for lib in "${LIBS[@]}"; do
    [ -f "$lib" ] || continue
    echo "== $(basename "$lib") (pass $rep)"
    GNSSB200_LIB="$lib" python "$ROOT/tools/form_sweep.py" "$STREAMS" "0:0:0"
  done
done
