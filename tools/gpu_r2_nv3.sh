#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_netvlad.py tests/test_gpu_knn.py tests/test_gpu_scale.py -m gpu -x -q 2>&1 | tail -6
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_netvlad.csv python tools/nv_run.py > gpurun_out/nv_list.log 2>&1; echo "list rc=$?"
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2_launches_netvlad.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=r; start=i; break
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
seq=[(r[ki][:60], float(r[vi].replace(',',''))/1000) for r in rows[start+2:] if len(r)>vi]
for k,v in seq[-5:]: print(f"{v:9.1f} us  {k}")
PY
timeout 300 python tools/nv_run.py
