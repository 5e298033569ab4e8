#!/bin/bash
# ncu --set full capture of the thread-per-pair tier on a 30k-locus step (single chunk)
tag=${1:-x}
CMD="python bench.py --loci 30000 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain30k_$tag.log 2>&1 || { echo plain failed; tail -5 gpurun_out/plain30k_$tag.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:k_align_thread -s 3 -c 1 -f -o gpurun_out/prof_thread_$tag $CMD > gpurun_out/ncu_thread_$tag.log 2>&1
echo rc=$?
