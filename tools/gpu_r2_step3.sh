#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python tools/nv_run.py
timeout 600 python bench.py --config 4 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --legs none > gpurun_out/r2_bench_cfg4_n1.json 2>/dev/null
python - <<'PY'
import json
for line in open("gpurun_out/r2_bench_cfg4_n1.json"):
    if line.startswith("{"):
        d=json.loads(line); print("cfg4 n1", d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_ms_per_step"])
PY
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench_n1.json"))
print("ms_per_step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "roofline", {k:round(d["roofline"][k],4) for k in ("frac","step_frac","kernel_ms_per_step","rescore_ms_per_step")})
c=d["config1"]; print("config1", {k:c[k] for k in ("images_per_s","superseg_per_s","aggregate_pca_s","host_synth_s","match_vote_s","recall_at_1_5")}, c["parity"])
print("fused", d["aggregate_pca_fused"]["fused"], "netvlad", d["netvlad"]["images_per_s"])
PY
