#!/bin/bash
# A/B of environment knobs on the device-resident step: run_gpu_ab.sh <loci> "<env A>" "<env B>" ...
loci=$1; shift
for v in "$@"; do
  env $v python bench.py --loci $loci --steps 3 --warmup 3 --no-cpu-baseline --no-census > gpurun_out/ab.json 2>gpurun_out/ab.err || tail -3 gpurun_out/ab.err
  python -c "
import json
d=json.loads(open('gpurun_out/ab.json').read().strip().splitlines()[-1])
print('[$v]', round(d['value']), round(d['ms_per_step'],2), {k: round(x,2) for k,x in d['breakdown_ms_per_step'].items()}, d['tier_pairs'], d['thread_tier_pairs'], d.get('long_tier_pairs'))"
done
