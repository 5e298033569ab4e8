#!/bin/bash
# round 2: parity tests, then device-resident A/B of the lockstep tiers (k_align_frame vs k_align_lock) at 300k loci
tag=$1
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -8
for v in new old; do
  if [ $v = old ]; then export TDB_OLD_LOCK=1; else unset TDB_OLD_LOCK; fi
  python bench.py --loci 300000 --steps 3 --warmup 3 --no-cpu-baseline --no-census > gpurun_out/ab_${v}_$tag.json 2>gpurun_out/ab_${v}_$tag.err || tail -3 gpurun_out/ab_${v}_$tag.err
  python -c "
import json
d=json.loads(open('gpurun_out/ab_${v}_$tag.json').read().strip().splitlines()[-1])
print('$v', round(d['value']), d['ms_per_step'], d['breakdown_ms_per_step'], d['tier_pairs'])"
done
