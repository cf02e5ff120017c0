#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds the *unmodified* reference C receiver (OSGPS-derived
# osgnss_next_step) from the sources where they lie under /root/reference into oracle/_ref/.
# Nothing is copied into the repo: the only generated files are include-redirect shims
# (the reference writes its includes with Windows back-slashes, e.g. #include ".\..\include\globals.h";
# on Linux that is a literal file name, so we create files of exactly that name under
# oracle/_ref/shim/ that #include the real header by absolute path) and a display stub
# (display/display.c needs <windows.h>, kbhit(); it is UI only: SURVEY.md §2 row 5).
#
# Outputs (all under oracle/_ref/, git-ignored, NOT gpurun-ignored so they travel to the GPU box):
#   libosgnss_ref.so    pristine arithmetic: correlator.c + gp2021.c + osgpsisr.c + osgnss_next_step.c
#                       (main renamed osgnss_main so the globals of '#define MAIN' exist)
#   libosgnss_ref34.so  same, but the three PRN tables are declared [34][2046] instead of [33][2046]
#                       (sed on the fly, never written to disk) so that the reference's own
#                       out-of-row read for PRN 32 (correlator.c:172-174 comment) is defined
#                       (row 33 = 0).  Arithmetic-neutral for PRN <= 31.
#   osgnss_ref          the stock command line receiver (./osgnss_ref -f record.bin), writes
#                       'e:\corr_out.csv' (literal file name) in the cwd like the reference does.
#   host_objs/*.o       the reference's host side WITHOUT correlator.c (main, gpsisr, gp2021
#                       accessors) -- linked against libgnssb200.so to form osgnss_gpu, the drop-in
#                       demonstration: reference host code unchanged, GPU correlator behind it.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF_ROOT="${REF_ROOT:-/root/reference}"
SRC="$REF_ROOT/trunk/GNSS_SOFTWARE_RECEIVERS/POSTPROCESSING_RECEIVERS/osgnss_next_step/src"
OUT="$HERE/_ref"
if [ ! -d "$SRC" ]; then
  echo "build_ref.sh: $SRC not present (GPU box?) -- keeping prebuilt oracle/_ref" >&2
  exit 0
fi
mkdir -p "$OUT/shim" "$OUT/host_objs"
mk_shim() { # $1 = literal include string used by the reference, $2 = real header
  printf '#include "%s"\n' "$2" > "$OUT/shim/$1"
}
mk_shim '.\..\include\globals.h'       "$SRC/include/globals.h"
mk_shim '.\include\globals.h'          "$SRC/include/globals.h"
mk_shim '.\..\correlator\correlator.h' "$SRC/correlator/correlator.h"
mk_shim '.\correlator\correlator.h'    "$SRC/correlator/correlator.h"
mk_shim '.\..\gp2021\gp2021.h'         "$SRC/gp2021/gp2021.h"
mk_shim '.\gp2021\gp2021.h'            "$SRC/gp2021/gp2021.h"
mk_shim '.\isr\osgpsisr.h'             "$SRC/isr/osgpsisr.h"
mk_shim '.\display\display.h'          "$SRC/display/display.h"
cat > "$OUT/display_stub.c" <<'EOS'
/* stub for the reference's win32 console UI (display/display.c) -- not arithmetic */
void clear_screen(void) {}
int display(void) { return 0; }
EOS
CF="-O2 -fcommon -w -fPIC -include errno.h -I$OUT/shim -I$SRC/include -I$SRC/correlator -I$SRC/isr -I$SRC/gp2021"
cd "$OUT"
gcc $CF -c "$SRC/gp2021/gp2021.c"        -o host_objs/gp2021.o
gcc $CF -c "$SRC/isr/osgpsisr.c"         -o host_objs/osgpsisr.o
gcc $CF -c "$SRC/osgnss_next_step.c"     -o host_objs/main.o
gcc $CF -Dmain=osgnss_main -c "$SRC/osgnss_next_step.c" -o host_objs/main_lib.o
gcc $CF -c display_stub.c                -o host_objs/display_stub.o
gcc $CF -I"$HERE/.." -c "$HERE/ref_driver.c" -o host_objs/ref_driver.o
gcc $CF -c "$SRC/correlator/correlator.c" -o correlator_ref.o
sed 's/\[33\]\[2046\]/[34][2046]/g' "$SRC/correlator/correlator.c" | gcc $CF -x c -c - -o correlator_ref34.o
gcc -shared -Wl,-Bsymbolic -o libosgnss_ref.so   correlator_ref.o   host_objs/gp2021.o host_objs/osgpsisr.o host_objs/main_lib.o host_objs/display_stub.o host_objs/ref_driver.o -lm
gcc -shared -Wl,-Bsymbolic -o libosgnss_ref34.so correlator_ref34.o host_objs/gp2021.o host_objs/osgpsisr.o host_objs/main_lib.o host_objs/display_stub.o host_objs/ref_driver.o -lm
gcc -o osgnss_ref   correlator_ref.o   host_objs/gp2021.o host_objs/osgpsisr.o host_objs/main.o host_objs/display_stub.o -lm
gcc -o osgnss_ref34 correlator_ref34.o host_objs/gp2021.o host_objs/osgpsisr.o host_objs/main.o host_objs/display_stub.o -lm
# drop-in demonstration: the reference's host side (main, gpsisr, gp2021 accessors) linked against
# libgnssb200.so INSTEAD of correlator.c.  -rdynamic exports the host's globals (REG_*, gps_*_ref, ...)
# so that the library's references bind to the executable's copies, as in the single-binary reference.
GPULIB="$HERE/../gnss_sdr_ru_b200"
if [ -f "$GPULIB/libgnssb200.so" ]; then
  gcc -o osgnss_gpu host_objs/gp2021.o host_objs/osgpsisr.o host_objs/main.o host_objs/display_stub.o \
      -L"$GPULIB" -lgnssb200 -Wl,-rpath,'$ORIGIN/../../gnss_sdr_ru_b200' -rdynamic -lm
fi
echo "built oracle/_ref: $(ls "$OUT" | tr '\n' ' ')"
