#!/bin/bash
# launch list (per-kernel device time) of one device-resident step at 100k loci
CMD="python bench.py --loci 100000 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_launch.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/launches_cur.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo rc=$?
