#!/bin/bash
# ncu --set full captures of selected kernels on a 100k-locus device-resident step (round 2)
# usage: run_gpu_ncu_r02.sh <tag> <kernel-regex> [<kernel-regex> ...]
tag=$1; shift
CMD="python bench.py --loci 100000 --steps 1 --warmup 3 --no-cpu-baseline --no-census"
$CMD > gpurun_out/plain100k_$tag.log 2>&1 || { tail -5 gpurun_out/plain100k_$tag.log; exit 1; }
for k in "$@"; do
  name=$(echo $k | tr -c 'A-Za-z0-9_\n' '_')
  ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:$k -s 3 -c 1 -f \
      -o gpurun_out/prof_${name}_$tag $CMD > gpurun_out/ncu_${name}_$tag.log 2>&1
  echo "$k rc=$?"
done
