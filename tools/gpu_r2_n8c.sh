#!/bin/bash
# final multi-GPU lines of the round (final build): config 2 at N = 8 with timeline, config 4 at N = 8, config 3 at N = 4
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29511 --nproc-per-node"
timeout 600 $TR 8 bench.py --gpus 8 --steps 20 --warmup 3 --trace gpurun_out/r2_timeline_n8.txt > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "cfg2 n8 rc=$?"
timeout 900 $TR 8 bench.py --gpus 8 --config 4 --steps 5 --warmup 3 > gpurun_out/r2_bench_cfg4.json 2> gpurun_out/r2_bench_cfg4.err; echo "cfg4 n8 rc=$?"
timeout 900 $TR 4 bench.py --gpus 4 --config 3 --steps 5 --warmup 3 > gpurun_out/r2_bench_cfg3.json 2> gpurun_out/r2_bench_cfg3.err; echo "cfg3 n4 rc=$?"
python - <<'PY'
import json
def load(p):
    for line in open(p):
        if line.startswith("{"): return json.loads(line)
for f in ("r2_bench_n8","r2_bench_cfg4","r2_bench_cfg3"):
    d=load(f"gpurun_out/{f}.json")
    print(f, {k:d[k] for k in ("value","ms_per_step","n_gpus")}, "e2e", d.get("e2e",{}).get("ms_per_step"), "ident", d.get("identity_check",{}).get("idx_identical"), "frac", round(d["roofline"]["frac"],3))
PY
