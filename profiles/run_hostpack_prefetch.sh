for nt in 0 1; do echo "nt $nt"; TDB_HOSTPACK_NT=$nt python profiles/hostpack_bw.py 2>&1 | grep "pinned=1" | grep -E "threads= 1 |threads= 8|threads=16" ; done
timeout 600 python bench.py --loci 100000 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',round(d['value']),'e2e',round(d['e2e']['value']), d['e2e_host'])"
