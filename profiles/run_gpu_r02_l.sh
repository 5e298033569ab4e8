#!/bin/bash
# round 2: one bench line at <loci> + the per-launch device times of one device-resident step at 300k loci
tag=$1; loci=${2:-937000}
python bench.py --loci $loci --steps 3 --warmup 3 --no-cpu-baseline --no-census > gpurun_out/t_$tag.json 2>gpurun_out/t_$tag.err || tail -3 gpurun_out/t_$tag.err
python -c "
import json
d=json.loads(open('gpurun_out/t_$tag.json').read().strip().splitlines()[-1])
print('$tag', round(d['value']), d['ms_per_step'], d['breakdown_ms_per_step'], d['tier_pairs'], d['thread_tier_pairs'])"
CMD="python bench.py --loci 300000 --steps 1 --warmup 3 --no-cpu-baseline --no-census"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_launch_$tag.log 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/launches_$tag.csv')) if len(r)>5]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); ui=h.index('Metric Unit')
seq=[(r[ki].split('(')[0][:60], float(r[vi].replace(',',''))*(1e-3 if r[ui]=='us' else (1e-6 if r[ui]=='ns' else 1))) for r in rows[1:]]
# last device-resident step = the last occurrence block: sum per kernel over the final quarter of launches
tail=seq[-len(seq)//6:]
agg=collections.OrderedDict()
for k,v in tail: agg[k]=agg.get(k,0)+v
for k,v in agg.items(): print(f'{v:8.3f} ms  {k}')
PY
