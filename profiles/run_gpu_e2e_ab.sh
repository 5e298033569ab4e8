#!/bin/bash
# A/B of the e2e legs under environment settings (full size, no CPU baseline / census): run_gpu_e2e_ab.sh "VAR=val ..." ...
run() { echo "== $*"; env $* python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-census 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value ms',round(d['ms_per_step'],2),'e2e ms',round(d['e2e']['ms_per_step'],2),'e2e_packed ms',round(d['e2e_packed']['ms_per_step'],2), 'queued', round(d['e2e_packed_host']['ms_last_chunk_queued'],1), 'call', round(d['e2e_packed_host']['ms_align_call'],1), 'chunks', d['e2e_packed_host']['chunks'])
"; }
for v in "$@"; do run $v; done
