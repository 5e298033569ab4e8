#!/bin/bash
# warp tier variants: one pair per warp (default), lockstep 2 pairs x 16 lanes, lockstep 4 pairs x 8 lanes (W = 64)
run() {
TDB_LIB=$PWD/$1 timeout 600 python bench.py --loci 100000 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
b=d['breakdown_ms_per_step']
print('$1 duo=$TDB_ENABLE_DUO','value',round(d['value']),'thread %.2f quad %.2f warp %.2f glob %.2f'%(b['align_thread'],b['align_quad'],b['align_warp'],b['align_global']),'tiers',d['tier_pairs'])
"
}
run build/sweep/lib_duoG16.so
TDB_ENABLE_DUO=1 run build/sweep/lib_duoG16.so
TDB_ENABLE_DUO=1 run build/sweep/lib_duoG8.so
