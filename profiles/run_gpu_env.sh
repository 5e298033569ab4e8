#!/bin/bash
# one device-resident bench line under an environment setting:  run_gpu_env.sh <tag> <loci> VAR=VALUE ...
tag=$1; loci=$2; shift 2
env "$@" python bench.py --loci $loci --steps 3 --warmup 3 --no-cpu-baseline --no-census > gpurun_out/t_$tag.json 2>gpurun_out/t_$tag.err || tail -3 gpurun_out/t_$tag.err
python -c "
import json
d=json.loads(open('gpurun_out/t_$tag.json').read().strip().splitlines()[-1])
print('$tag', '$*', round(d['value']), d['ms_per_step'], d['breakdown_ms_per_step'], d['tier_pairs'], d['thread_tier_pairs'])"
