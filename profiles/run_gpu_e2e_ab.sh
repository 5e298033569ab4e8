#!/bin/bash
# A/B of the e2e legs under environment toggles (300k loci, no CPU baseline / census)
run() { echo "== $*"; env "$@" python bench.py --loci 300000 --steps 3 --warmup 2 --no-cpu-baseline --no-census 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value ms',round(d['ms_per_step'],2),'e2e ms',round(d['e2e']['ms_per_step'],2),'e2e_packed ms',round(d['e2e_packed']['ms_per_step'],2), 'queued', round(d['e2e_packed_host']['ms_last_chunk_queued'],1), 'call', round(d['e2e_packed_host']['ms_align_call'],1), d['breakdown_ms_per_step']['align_thread'])
"; }
run A=1
run TDB_TEST_D2H_S=1 TDB_REDUCE_SLICES=1
run TDB_TEST_D2H_S=1 TDB_REDUCE_SLICES=1 TDB_TEST_EARLY_LOCI=1
