#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pca.py tests/test_gpu_e2e.py tests/test_gpu_aggregate.py tests/test_gpu_anyloc.py -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py --steps 5 --warmup 3 --legs aggregation,pca --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_f1.json 2> gpurun_out/r2_bench_f1.err; echo "bench rc=$?"; tail -5 gpurun_out/r2_bench_f1.err
python - <<'PY'
import json
for line in open("gpurun_out/r2_bench_f1.json"):
    if line.startswith("{"):
        d=json.loads(line)
        print(json.dumps(d.get("aggregate_pca_fused"), indent=1))
        print("agg", d["aggregation"]["kernel_ms"], d["aggregation"]["ms_per_batch"], "pca", d["pca"]["kernel_ms"])
PY
