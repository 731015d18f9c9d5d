#!/bin/bash
mkdir -p gpurun_out
for ch in 2 4 8 16; do
SEGVLAD_RT_CH=$ch CALLS=2 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'rt_from_tokens' --csv --log-file gpurun_out/rt_ch.csv python tools/agg_run.py > /dev/null 2>&1
echo "CH $ch: $(grep rt_from_tokens gpurun_out/rt_ch.csv | tail -1 | awk -F'","' '{print $NF}')"
done
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_mid.json 2> gpurun_out/bench_mid.err; echo "bench rc=$?"; cat gpurun_out/bench_mid.json
