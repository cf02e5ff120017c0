#!/usr/bin/env bash
# Runs on the GPU box (via gpurun): a short bench run, then -- only after it exited 0 -- the ncu launch list of the same
# command and one full capture of the tracking kernel from it (packed input; TAG_int8 the same with --fmt int8).
set -x
mkdir -p gpurun_out
TAG=${1:-r02}
SHORT="python bench.py --steps 1 --warmup 3 --seconds 2 --no-cpu --no-also --no-acq"
$SHORT > gpurun_out/short_$TAG.json 2> gpurun_out/short_$TAG.err || { tail -5 gpurun_out/short_$TAG.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_bench_launches.csv $SHORT > gpurun_out/ncu_list_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:track_ws_kernel -s 3 -c 1 -f -o gpurun_out/${TAG}_track_ws_bench $SHORT > gpurun_out/ncu_full_$TAG.log 2>&1
$SHORT --fmt int8 > gpurun_out/short_${TAG}_int8.json 2>/dev/null && ncu --set full --clock-control none --import-source on -k regex:track_ws_kernel -s 3 -c 1 -f -o gpurun_out/${TAG}_track_ws_int8 $SHORT --fmt int8 > gpurun_out/ncu_full_${TAG}_int8.log 2>&1
if [ -n "${WITH_ACQ:-}" ]; then
  python tools/acq_probe.py 10 20 32 > gpurun_out/acq_probe_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:acq_rows_kernel -c 1 -f -o gpurun_out/${TAG}_acq_rows python tools/acq_probe.py 10 20 32 > gpurun_out/ncu_acq_$TAG.log 2>&1
fi
tail -c 600 gpurun_out/short_$TAG.json
