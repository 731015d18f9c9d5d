#!/bin/bash
# r2: kNN tests, bench A/B of the re-score / refine variants, launch list of one bench run
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_knn.py tests/test_gpu_scale.py tests/test_gpu_vote.py tests/test_gpu_e2e.py -m gpu -x -q 2>&1 | tail -8
for cfg in "1 1" "0 1" "1 0"; do
  set -- $cfg
  echo "== RESCORE_REF=$1 LOOSE=$2"
  SEGVLAD_KNN_RESCORE_REF=$1 SEGVLAD_KNN_LOOSE=$2 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-aggregation > gpurun_out/r2_knn_ab_$1$2.json 2> gpurun_out/r2_knn_ab_$1$2.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_knn_ab_$1$2.json"))
print("ms_per_step", round(d["ms_per_step"],3), "filter", round(d["roofline"]["kernel_ms_per_step"],3), "rescore", round(d["roofline"]["rescore_ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3))
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_knn.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-aggregation > gpurun_out/ncu_list.log 2>&1; echo "list rc=$?"
