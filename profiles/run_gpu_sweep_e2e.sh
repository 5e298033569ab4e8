#!/bin/bash
for lib in build/sweep/lib_*.so; do for cp in 2097152 1048576; do
TDB_CHUNK_PAIRS=$cp TDB_LIB=$PWD/$lib timeout 600 python bench.py --loci ${1:-200000} --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$lib', $cp, 'value',round(d['value']),'e2e',round(d['e2e']['value']),d['e2e_chunks'],d['e2e_host'])
"
done; done
