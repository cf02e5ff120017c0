#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds the reference's own GPS-SDR fixed-point primitives (FFT class with its
# portable NO_SIMD butterflies, x86_* helpers, sine / wipe-off generators) from the sources where they lie
# under /root/reference into oracle/_ref/libgpssdr_ref.so.  Nothing is copied.  The Acquisition class
# itself needs the USRP headers and cannot be built.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF_ROOT="${REF_ROOT:-/root/reference}"
RT="$REF_ROOT/trunk/GNSS_SOFTWARE_RECEIVERS/REALTIME_RECEIVERS/GPS/GPS_SDR_REAL_TIME_GPS_RECEIVER"
OUT="$HERE/_ref"
if [ ! -d "$RT" ]; then
  echo "build_ref_gpssdr.sh: $RT not present (GPU box?) -- keeping prebuilt oracle/_ref" >&2
  exit 0
fi
mkdir -p "$OUT/gpssdr"
INC="-I$RT/includes -I$RT/objects -I$RT/simd -I$RT/accessories -I$RT/main -I$RT/usrp"
CF="-O2 -w -fpermissive -fPIC"
cd "$OUT/gpssdr"
g++ $CF -DNO_SIMD $INC -c "$RT/objects/fft.cpp" -o fft.o
g++ $CF $INC -c "$RT/simd/x86.cpp" -o x86.o
g++ $CF $INC -c "$RT/accessories/misc.cpp" -o misc.o
g++ $CF $INC -c "$HERE/gpssdr_ref_driver.cpp" -o driver.o
g++ -shared -Wl,-Bsymbolic -o "$OUT/libgpssdr_ref.so" fft.o x86.o misc.o driver.o -lm
echo "built $OUT/libgpssdr_ref.so"
