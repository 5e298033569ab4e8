#!/bin/bash
# ncu --set full capture of the long tier (k_align_lock<LongCfg>) on the long-expansion stress set (config 4, 300 loci)
tag=${1:-x}
CMD="python profiles/bench_configs.py --long-loci 300 --skip 135"
$CMD > gpurun_out/plain_cfg4_$tag.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_align_lock.*Li32ELi256 -s 1 -c 1 -f -o gpurun_out/prof_long_$tag $CMD > gpurun_out/ncu_long_$tag.log 2>&1
echo rc=$?
ncu -i gpurun_out/prof_long_$tag.ncu-rep --page raw --csv | python profiles/ncu_key_metrics.py
