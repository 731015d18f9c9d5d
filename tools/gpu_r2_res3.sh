#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_r2_res.sh
( echo "== la timeline"; NPASS=30 timeout 300 python tools/agg_tc_timeline.py 2>&1 | tail -32 ) > gpurun_out/r2_agg_single_sweep_probe.txt 2>&1
tail -3 gpurun_out/r2_agg_single_sweep_probe.txt
