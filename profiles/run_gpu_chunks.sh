#!/bin/bash
for cp in 2097152 1048576 524288; do
TDB_CHUNK_PAIRS=$cp timeout 600 python bench.py --loci 100000 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('chunk_pairs',$cp,'value',round(d['value']),'e2e',round(d['e2e']['value']), 'chunks', d['e2e_chunks'], d['e2e_host'])
"
done
