#!/usr/bin/env bash
# speed (no profiler) and DRAM / L2-write traffic of the tracking kernel on the bench shape
mkdir -p gpurun_out
SHORT="python bench.py --steps 2 --warmup 3 --seconds 2 --no-cpu --no-also --no-acq"
for i in 1 2; do
  $SHORT 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'kernel_ms', d['roofline']['kernel_ms'])"
done
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,l1tex__m_l1tex2xbar_write_sectors_mem_lg_op_st.sum,smsp__inst_executed.sum --clock-control none -k regex:track_ws_kernel -s 3 -c 1 --csv --log-file gpurun_out/traffic.csv $SHORT > /dev/null 2>&1
grep -v "^==" gpurun_out/traffic.csv | python -c "
import csv,sys
for r in csv.DictReader(sys.stdin): print('  ', r['Metric Name'], r['Metric Unit'], r['Metric Value'])"
