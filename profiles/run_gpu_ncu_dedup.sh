#!/bin/bash
# ncu --set full capture of the duplicate-elimination kernels on a 100k-locus device-resident step
tag=${1:-x}
CMD="python bench.py --loci 100000 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain100k_$tag.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k 'regex:k_pair_classify|k_pair_insert|k_seq_canon|k_seq_insert|k_hash|k_pack|k_scatter|k_reduce' -s 32 -c 8 -f -o gpurun_out/prof_dedup_$tag $CMD > gpurun_out/ncu_dedup_$tag.log 2>&1
echo rc=$?
