#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py > gpurun_out/mgpu_check_n$N.log 2>&1; echo "mgpu rc=$?"; tail -3 gpurun_out/mgpu_check_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"; grep '^{' gpurun_out/bench_n$N.json | cut -c1-700; tail -2 gpurun_out/bench_n$N.err
