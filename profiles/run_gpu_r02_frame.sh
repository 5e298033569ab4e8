#!/bin/bash
# round 2: parity tests of the frame kernels + ncu --set full captures on a 100k-locus device-resident step
# usage: run_gpu_r02_frame.sh <tag> [pytest -k expression]
tag=$1; sel=${2:-}
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -q -m gpu -x ${sel:+-k "$sel"} 2>&1 | tail -15
CMD="python bench.py --loci 100000 --steps 1 --warmup 3 --no-cpu-baseline --no-census"
$CMD > gpurun_out/plain100k_$tag.log 2>&1 || { tail -5 gpurun_out/plain100k_$tag.log; exit 1; }
python -c "
import json
d=json.loads(open('gpurun_out/plain100k_$tag.log').read().strip().splitlines()[-1])
print(d['breakdown_ms_per_step'], d['tier_pairs'])"
for k in "k_align_frame.*Li16ELi64" "k_align_frame.*Li8ELi32"; do
  name=$(echo $k | tr -c 'A-Za-z0-9_\n' '_')
  ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:$k -s 3 -c 1 -f \
      -o gpurun_out/prof_${name}_$tag $CMD > gpurun_out/ncu_${name}_$tag.log 2>&1
  echo "$k rc=$?"
done
