#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29511 --nproc-per-node"
timeout 600 $TR 8 bench.py --gpus 8 --steps 20 --warmup 3 --no-e2e --trace gpurun_out/r2_timeline_n8_b.txt > gpurun_out/r2_bench_n8_b.json 2> gpurun_out/r2_bench_n8_b.err; echo "cfg2 n8 rc=$?"
grep "^{" gpurun_out/r2_bench_n8_b.json | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d.get('identity_check',{}).get('idx_identical'))"
grep -n "nccl\|merge_packed" gpurun_out/r2_timeline_n8_b.txt | tail -4
