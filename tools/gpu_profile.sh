#!/usr/bin/env bash
# Runs on the GPU box (via gpurun): the default bench, then -- only after it exited 0 -- the ncu launch
# list of a short bench command and one full capture of the tracking kernel from that same command.
set -x
mkdir -p gpurun_out
TAG=${1:-r01c}
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err || { tail -5 gpurun_out/bench_$TAG.err; exit 1; }
SHORT="python bench.py --steps 1 --warmup 3 --seconds 2 --no-cpu --no-also --no-acq"
$SHORT > gpurun_out/short_$TAG.json 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $SHORT > gpurun_out/ncu_list_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:track_ws_kernel -s 3 -c 1 -f -o gpurun_out/${TAG}_track_ws $SHORT > gpurun_out/ncu_full_$TAG.log 2>&1
[ -n "${SKIP_GSA:-}" ] || python tools/gpssdr_probe.py > gpurun_out/gpssdr_probe_$TAG.log 2>&1 && [ -z "${SKIP_GSA:-}" ] && ncu --set full --clock-control none --import-source on -k regex:gsa_weak_kernel -c 1 -f -o gpurun_out/${TAG}_gsa_weak python tools/gpssdr_probe.py > gpurun_out/ncu_gsa_$TAG.log 2>&1
tail -c 1500 gpurun_out/bench_$TAG.json
