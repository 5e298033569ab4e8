#!/bin/bash
# tuning sweep: same bench, different builds of the library (build/sweep/lib_*.so)
for lib in build/sweep/lib_*.so; do
TDB_LIB=$PWD/$lib timeout 600 python bench.py --loci 100000 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
b=d['breakdown_ms_per_step']
print('$lib','value',round(d['value']),'thread %.2f quad %.2f warp %.2f glob %.2f'%(b['align_thread'],b['align_quad'],b['align_warp'],b['align_global']),'thread',d['thread_tier_pairs'],'tiers',d['tier_pairs'])
"
done
