#!/bin/bash
# r2: the multi-GPU evidence of the round on one 8-GPU box
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29511 --nproc-per-node"
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 $TR 8 tests/mgpu_check.py 2>&1 | tail -2
timeout 900 $TR 8 bench.py --gpus 8 --steps 20 --warmup 3 --trace gpurun_out/r2_timeline_n8.txt > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "cfg2 n8 rc=$?"; tail -2 gpurun_out/r2_bench_n8.err
timeout 900 $TR 8 bench.py --gpus 8 --config 4 --steps 5 --warmup 3 > gpurun_out/r2_bench_cfg4.json 2> gpurun_out/r2_bench_cfg4.err; echo "cfg4 n8 rc=$?"; tail -2 gpurun_out/r2_bench_cfg4.err
timeout 900 $TR 4 bench.py --gpus 4 --config 3 --steps 5 --warmup 3 > gpurun_out/r2_bench_cfg3.json 2> gpurun_out/r2_bench_cfg3.err; echo "cfg3 n4 rc=$?"; tail -2 gpurun_out/r2_bench_cfg3.err
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --legs netvlad --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_netvlad_n1.json 2>/dev/null; echo "nv n1 rc=$?"
for n in 2 4 8; do
  timeout 600 $TR $n bench.py --gpus $n --steps 3 --warmup 3 --legs netvlad --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_netvlad_n$n.json 2> gpurun_out/r2_bench_netvlad_n$n.err; echo "nv n$n rc=$?"
done
timeout 600 $TR 4 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r2_bench_n4.json 2>/dev/null; echo "cfg2 n4 rc=$?"
python - <<'PY'
import json, glob
def load(p):
    for line in open(p):
        if line.startswith("{"): return json.loads(line)
for f in sorted(glob.glob("gpurun_out/r2_bench_n[48].json")+glob.glob("gpurun_out/r2_bench_cfg[34].json")):
    d=load(f)
    if d: print(f, {k:d[k] for k in ("value","ms_per_step","n_gpus")}, "e2e", d.get("e2e",{}).get("ms_per_step"), "ident", d.get("identity_check",{}).get("idx_identical"), "frac", round(d["roofline"]["frac"],3))
for f in sorted(glob.glob("gpurun_out/r2_bench_netvlad_n*.json")):
    d=load(f)
    if d: print(f, d.get("netvlad",{}).get("images_per_s"), d.get("netvlad",{}).get("ms_per_batch"))
PY
