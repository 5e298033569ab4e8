#!/bin/bash
# quick GPU iteration: parity tests + a 100k-locus bench (not the judged configuration)
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -15
timeout 600 python bench.py --loci ${1:-100000} --steps 3 --warmup 3 --cpu-sample-loci 50000 2>gpurun_out/quick.err | tee gpurun_out/quick.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'cpu',round(d['cpu_baseline']['value']) if d['cpu_baseline'] else None, 'chunks', d['e2e_chunks'])
print(d['breakdown_ms_per_step'])
print(d['e2e_host']); print('thread',d['thread_tier_pairs'],'tiers',d['tier_pairs'],'ident',d['pairs_identical'],'dup',d['pairs_duplicate'],'closed',d['pairs_closed_form'], 'frac', d['roofline']['int_issue']['frac'], d['roofline']['kernel'])
"
tail -3 gpurun_out/quick.err
