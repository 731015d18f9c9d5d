#!/bin/bash
mkdir -p gpurun_out
for f in 2.5 4 6 10; do
  SEGVLAD_KNN_FILL=$f timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-aggregation 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('fill=$f', 'ms/step', round(d['ms_per_step'],3), 'tc_ms', round(r['kernel_ms_per_step'],3), 'launches', r['launches_per_step'], 'rescore', round(r['rescore_ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3))"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 0 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-aggregation > gpurun_out/ncu_list.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/launches.csv')) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value'); iid=hdr.index('ID')
agg=collections.OrderedDict()
# the bench runs 3 warm-up + 2 timed resident steps, then e2e steps; take launches of resident steps 4-5 by position is complex: print totals per kernel name over all
tot=collections.Counter(); cnt=collections.Counter()
for r in rows[1:]:
    name=r[ik].split('(')[0][-60:]; tot[name]+=float(r[iv].replace(',','')); cnt[name]+=1
for n,v in tot.most_common(): print(f'{n:62s} n={cnt[n]:4d} total_us={v/1e3:10.1f} avg_us={v/1e3/cnt[n]:8.1f}')
PY
