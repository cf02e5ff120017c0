#!/usr/bin/env bash
# Runs on the GPU box (via gpurun).  compute-sanitizer is closed on this pool (it answers "compute-sanitizer is closed on
# this pool and stays closed ... find a bad access with bounds checks and asserts of your own, small cases, and a
# comparison with the CPU reference"), so the evidence is a -DTRACK_CHECK build of the library: every shared-memory
# address the correlator warps form, the ownership rules of the segment partition and the work-queue hand-over are
# checked on the device and counted; the same build runs the bit-exact oracle comparisons.
TAG=${1:-r02}
mkdir -p gpurun_out
OUT=gpurun_out/${TAG}_sanitizer.txt
{
  echo "# compute-sanitizer on this pool:"
  /usr/local/cuda/bin/compute-sanitizer --tool memcheck python -c "print(1)" 2>&1 | tail -1
  echo
  echo "# substitute: library rebuilt with -DTRACK_CHECK (device-side address / invariant checks, csrc/track_common.cuh TCHECK)"
} > $OUT
# self-test of the instrumentation: with the mixer-table bound halved on purpose the counter must be non-zero
(cd gnss_sdr_ru_b200/csrc && touch track.cu && EXTRA_NVCC_FLAGS="-DTRACK_CHECK -DTRACK_CHECK_SELFTEST" bash build.sh > /dev/null 2>&1)
python - >> $OUT 2>&1 <<'PY'
import ctypes as C, runpy, sys
sys.path.insert(0, ".")
sys.argv = ["tools/track_digest.py", "2", "40", "0"]
runpy.run_path(sys.argv[0], run_name="__main__")
from gnss_sdr_ru_b200.lib import lib
out = (C.c_uint * 8)()
n = lib().gnssb200_track_check_failures(out)
print(f"self-test (mixer-table bound halved on purpose): violations = {n}  per check {list(out)}  -> the checks are live: {n > 0 and out[3] == n}")
PY
(cd gnss_sdr_ru_b200/csrc && touch track.cu && EXTRA_NVCC_FLAGS=-DTRACK_CHECK bash build.sh > /dev/null 2>&1)
cat > gpurun_out/check_run.py <<'PY'
import ctypes as C, subprocess, sys, os
sys.path.insert(0, ".")
from gnss_sdr_ru_b200.lib import lib
L = lib()
names = ["run starts on a code-NCO wrap", "code-table entries inside the table window", "sample window inside the tile",
         "mixer-table entry inside the table", "owned segments inside the block", "closed-form table index inside the window",
         "queue item behind the ticket exists", "slice hand-over: previous slice stored its state first"]
def report(label):
    out = (C.c_uint * 8)()
    n = L.gnssb200_track_check_failures(out)
    print(f"{label}: violations = {n}  per check {list(out)}", flush=True)
    return n
import runpy
for argv, label in ((["tools/track_digest.py", "64", "520", "8"], "64 streams x 12 channels x 520 blocks, work-queue slices of 8 blocks (packed, segment kernel)"),
                    (["tools/track_digest.py", "64", "300", "37"], "64 streams x 300 blocks, slices of 37 blocks"),
                    (["tools/track_digest.py", "3", "1200", "0"], "3 streams x 1200 blocks, one item per channel (384-thread segment kernel)")):
    sys.argv = argv
    runpy.run_path(argv[0], run_name="__main__")
    report(label)
print("checks:", "; ".join(f"[{i}] {n}" for i, n in enumerate(names)))
PY
python gpurun_out/check_run.py >> $OUT 2>&1
echo >> $OUT
echo "# bit-exact oracle comparisons with the checked build (pytest tests/test_track_gpu.py -k 'kernel_forms or segment_form or work_queue or config5')" >> $OUT
python - >> $OUT 2>&1 <<'PY'
import ctypes as C, sys, pytest
sys.path.insert(0, ".")
rc = pytest.main(["tests/test_track_gpu.py", "-m", "gpu", "-q", "-k", "kernel_forms or segment_form or work_queue", "-p", "no:cacheprovider"])
from gnss_sdr_ru_b200.lib import lib
out = (C.c_uint * 8)()
n = lib().gnssb200_track_check_failures(out)
print(f"pytest exit code {int(rc)}; device-side violations during those tests = {n}  per check {list(out)}")
PY
(cd gnss_sdr_ru_b200/csrc && touch track.cu && bash build.sh > /dev/null 2>&1)
rm -f gpurun_out/check_run.py
cat $OUT
