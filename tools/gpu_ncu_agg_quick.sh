#!/bin/bash
# quick memory-traffic counters of the aggregation kernel (few replays), for several experiment modes
mkdir -p gpurun_out
for e in ${EXPS:-0 1}; do
SEGVLAD_AGG_EXP=$e CALLS=2 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sector_op_read_hit_rate.pct,lts__t_sector_op_write_hit_rate.pct,lts__t_bytes.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed \
  --clock-control none -k regex:'aggregate_tc_kernel' --launch-skip 1 --launch-count 1 --csv --log-file gpurun_out/ncu_aggq_$e.csv python tools/agg_run.py > /dev/null 2>&1
echo "== exp $e"; python - <<PY
import csv
rows = [r for r in csv.reader(open('gpurun_out/ncu_aggq_$e.csv')) if len(r) > 10]
h = rows[0]
for r in rows[1:]:
    d = dict(zip(h, r)); print(d['Metric Name'], d['Metric Unit'], d['Metric Value'])
PY
done
