#!/bin/bash
# round-1 baseline capture on one B200: tests, host facts, full-size bench, ncu launch list
mkdir -p gpurun_out
lscpu | head -25 > gpurun_out/host.txt; free -g >> gpurun_out/host.txt; nvidia-smi >> gpurun_out/host.txt
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_full_r01a.json 2> gpurun_out/bench_full_r01a.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_full_r01a.err; cat gpurun_out/bench_full_r01a.json | head -c 3000
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r01a.json 2>/dev/null; cat gpurun_out/bench_ref_r01a.json | head -c 1500
timeout 600 python bench.py --loci 30000 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_r01b.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01b.csv \
  python bench.py --loci 30000 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r01b.log 2>&1
echo done
