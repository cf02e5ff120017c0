set -x
TAG=r02h
SHORT="python bench.py --steps 1 --warmup 3 --seconds 2 --no-cpu --no-also --no-acq"
$SHORT > gpurun_out/short_$TAG.json 2> gpurun_out/short_$TAG.err || { tail -5 gpurun_out/short_$TAG.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_bench_launches.csv $SHORT > gpurun_out/ncu_list_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:track_ws_kernel -s 3 -c 1 -f -o gpurun_out/${TAG}_track_ws_bench $SHORT > gpurun_out/ncu_full_$TAG.log 2>&1
tail -c 400 gpurun_out/short_$TAG.json
