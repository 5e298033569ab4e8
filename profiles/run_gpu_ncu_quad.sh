#!/bin/bash
# one --set full capture of the quad and warp tier kernels (30k-locus bench), after a plain run
CMD="python bench.py --loci 30000 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_ncu.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_align_lock -s 3 -c 1 -f -o gpurun_out/prof_quad_r01e $CMD > gpurun_out/ncu_quad.log 2>&1
tail -n 2 gpurun_out/ncu_quad.log gpurun_out/ncu_warp.log
