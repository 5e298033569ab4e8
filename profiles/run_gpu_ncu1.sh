#!/bin/bash
# one ncu --set full capture of a kernel (regex) on a 100k-locus device-resident step
tag=$1; k=$2
CMD="python bench.py --loci 100000 --steps 1 --warmup 3 --no-cpu-baseline --no-census"
name=$(echo $k | tr -c 'A-Za-z0-9_\n' '_')
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_${name}_$tag $CMD > gpurun_out/ncu_${name}_$tag.log 2>&1
echo "$k rc=$?"
