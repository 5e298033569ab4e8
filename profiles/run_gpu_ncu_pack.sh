#!/bin/bash
# ncu --set full capture of k_pack / k_hash on a 100k-locus device-resident step
tag=${1:-x}
CMD="python bench.py --loci 100000 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain100k_$tag.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_pack -s 3 -c 1 -f -o gpurun_out/prof_pack_$tag $CMD > gpurun_out/ncu_pack_$tag.log 2>&1
echo rc=$?
