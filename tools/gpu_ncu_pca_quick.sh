#!/bin/bash
mkdir -p gpurun_out
CALLS=2 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_bytes.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,l1tex__t_bytes_pipe_lsu_mem_local_op_st.sum,l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum \
  --clock-control none -k regex:'pca_tc_kernel' --launch-skip 1 --launch-count 1 --csv --log-file gpurun_out/ncu_pcaq.csv python tools/pca_run.py > /dev/null 2>&1
python - <<PY
import csv
rows = [r for r in csv.reader(open('gpurun_out/ncu_pcaq.csv')) if len(r) > 10]
h = rows[0]
for r in rows[1:]:
    d = dict(zip(h, r)); print(d['Metric Name'], d['Metric Unit'], d['Metric Value'])
PY
