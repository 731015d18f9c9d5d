#!/bin/bash
# 2-GPU check of the lean multi-GPU step: identity to single GPU, bench line, timeline
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tests/mgpu_check.py 2>&1 | tail -3
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 --trace gpurun_out/r2_timeline_n2.txt > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/r2_bench_n2.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench_n2.json"))
print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, d["e2e"], d.get("identity_check"))
PY
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --legs none --no-cpu-baseline > gpurun_out/r2_bench_n1_b.json 2>/dev/null
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench_n1_b.json"))
print("n1", {k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["ms_per_step"])
PY
timeout 900 $TR bench.py --gpus 2 --config 4 --steps 3 --warmup 3 --no-e2e > gpurun_out/r2_bench_cfg4_n2.json 2> gpurun_out/r2_bench_cfg4_n2.err; echo "cfg4 n2 rc=$?"; tail -3 gpurun_out/r2_bench_cfg4_n2.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench_cfg4_n2.json"))
print("cfg4", {k:d[k] for k in ("value","ms_per_step","n_gpus")}, d["roofline"]["frac"], d["roofline"]["kernel_ms_per_step"], d["roofline"]["rescore_ms_per_step"], d.get("identity_check"))
PY
tail -5 gpurun_out/r2_timeline_n2.txt
