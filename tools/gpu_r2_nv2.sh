#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_netvlad.csv python tools/nv_run.py > gpurun_out/nv_list.log 2>&1; echo "list rc=$?"
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2_launches_netvlad.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=r; start=i; break
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
seq=[(r[ki][:60], float(r[vi].replace(',',''))/1000) for r in rows[start+2:] if len(r)>vi]
for k,v in seq[-8:]: print(f"{v:9.1f} us  {k}")
PY
{
echo "compute-sanitizer --tool racecheck (B200, r2 sources) over small cases of the shared-memory kernels"
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_knn.py -q -k "small_vs_oracle or fewer_refs or merge_packed" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|hazard|error" | head -8
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_vote.py tests/test_gpu_netvlad.py -q -k "golden or bitexact or 64-5-16 or 96-9-32" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|hazard|error" | head -8
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_aggregate.py -q -k "golden or mask_centroids or 64-32-0-5" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|hazard|error" | head -8
} > gpurun_out/r2_sanitizer_racecheck.txt 2>&1
cat gpurun_out/r2_sanitizer_racecheck.txt
