#!/bin/bash
# tier 1 variants: two pairs per warp in lockstep (default, build with -DTDB_DUO_G=16), one pair per warp
# (TDB_DISABLE_DUO=1), four pairs x 8 lanes (a build with -DTDB_DUO_G=8 at build/sweep/lib_duoG8.so)
run() {
TDB_LIB=$1 timeout 600 python bench.py --loci 100000 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
b=d['breakdown_ms_per_step']
print('$2','value',round(d['value']),'thread %.2f quad %.2f tier1 %.2f glob %.2f'%(b['align_thread'],b['align_quad'],b['align_warp'],b['align_global']),'tiers',d['tier_pairs'])
"
}
run "" "two pairs x 16 lanes (default)"
TDB_DISABLE_DUO=1 run "" "one pair per warp"
[ -f build/sweep/lib_duoG8.so ] && run $PWD/build/sweep/lib_duoG8.so "four pairs x 8 lanes"
