#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_knn.py tests/test_gpu_scale.py tests/test_gpu_vote.py tests/test_gpu_e2e.py -m gpu -x -q 2>&1 | tail -8
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_n1_a.json 2> gpurun_out/r2_bench_n1_a.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_n1_a.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench_n1_a.json"))
print("ms_per_step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "roofline", {k:d["roofline"][k] for k in ("frac","step_frac","kernel_ms_per_step","rescore_ms_per_step")})
for leg in ("aggregation","pca","netvlad","config1","cpu_baseline"):
    print(leg, json.dumps(d.get(leg))[:1500])
PY
