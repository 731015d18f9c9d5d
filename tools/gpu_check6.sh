#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_knn.py tests/test_gpu_e2e.py tests/test_gpu_scale.py tests/test_gpu_anyloc.py -q -x > gpurun_out/pytest_knn.log 2>&1; echo "knn rc=$?"
tail -5 gpurun_out/pytest_knn.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-aggregation > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); r=d['roofline']; print('ms/step', round(d['ms_per_step'],3), 'tc_ms', round(r['kernel_ms_per_step'],3), 'rescore', round(r['rescore_ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'value', d['value']/1e9, 'e2e', d['e2e']['value']/1e9)"
tail -3 gpurun_out/bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-aggregation > gpurun_out/ncu_list.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/launches.csv')) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
tot=collections.Counter(); cnt=collections.Counter()
for r in rows[1:]:
    name=r[ik].split('(')[0][-50:]; tot[name]+=float(r[iv].replace(',','')); cnt[name]+=1
for n,v in tot.most_common(8): print(f'{n:52s} n={cnt[n]:4d} total_us={v/1e3:10.1f} avg_us={v/1e3/cnt[n]:8.1f}')
PY
