#!/bin/bash
# ncu --set full capture of the global-scratch tier on the long-expansion stress set (config 4, 300 loci)
tag=${1:-x}
CMD="python profiles/bench_configs.py --long-loci 300 --skip 135"
$CMD > gpurun_out/plain_cfg4_$tag.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_align_global -s 1 -c 1 -f -o gpurun_out/prof_global_$tag $CMD > gpurun_out/ncu_global_$tag.log 2>&1
echo rc=$?
