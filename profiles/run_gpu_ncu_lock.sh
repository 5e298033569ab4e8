#!/bin/bash
# ncu --set full captures of the two lockstep tiers on a 30k-locus step (what run_gpu_final.sh does for them)
tag=${1:-x}
CMD="python bench.py --loci 30000 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain30k_$tag.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:LockCfgILi8E -s 3 -c 1 -f -o gpurun_out/prof_quad_$tag $CMD > gpurun_out/ncu_quad_$tag.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:LockCfgILi16E -s 3 -c 1 -f -o gpurun_out/prof_duo_$tag $CMD > gpurun_out/ncu_duo_$tag.log 2>&1
echo rc=$?
