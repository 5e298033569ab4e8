#!/bin/bash
# round 2, third session: GPU tests, one device-resident bench line (breakdown) at <loci>, config 4 stage times
# usage: run_gpu_r03.sh <tag> [loci] [pytest-selector|none] [cfg4: 1|0]
tag=$1; loci=${2:-937000}; sel=${3:-tests}; cfg4=${4:-1}
mkdir -p gpurun_out
if [ "$sel" != "none" ]; then timeout 1500 python -m pytest $sel -q -m gpu -x 2>&1 | tail -8; fi
python bench.py --loci $loci --steps 3 --warmup 3 --no-cpu-baseline --no-census > gpurun_out/t_$tag.json 2>gpurun_out/t_$tag.err || tail -3 gpurun_out/t_$tag.err
python -c "
import json
d=json.loads(open('gpurun_out/t_$tag.json').read().strip().splitlines()[-1])
print('$tag', round(d['value']), d['ms_per_step'], d['breakdown_ms_per_step'], d['tier_pairs'], d['thread_tier_pairs'], 'e2e', round(d['e2e']['value']), 'e2e_packed', round(d['e2e_packed']['value']))"
if [ "$cfg4" = "1" ]; then python profiles/config4_stages.py 2>&1 | grep "call ms" | tail -1 | cut -c1-260 | tee gpurun_out/cfg4_$tag.txt; fi
