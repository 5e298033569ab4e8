#!/bin/bash
# End-of-round capture on one B200: parity tests, smoke, both bench arms at the judged configuration,
# per-launch device times, and one --set full capture of each of the three alignment kernels.
tag=${1:-r01g}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$tag.json 2>/dev/null
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_full_$tag.json 2> gpurun_out/bench_full_$tag.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_full_$tag.json")); r=json.load(open("gpurun_out/bench_ref_$tag.json"))
print("value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"ref arm",round(r["value"]),"cpu",round(d["cpu_baseline"]["value"]),"ms/step",d["ms_per_step"])
print(d["breakdown_ms_per_step"]); print(d["e2e_host"], d["e2e_chunks"]); print(d["roofline"]["int_issue"]["frac"], d["roofline"]["kernel"], d["clocks"])
PY
CMD="python bench.py --loci 100000 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_$tag.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_launch_$tag.log 2>&1
echo launches rc=$?
# the judged command itself (937k loci): first 150 launches = the counting step, the warm-ups and the timed steps
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/launches_full_$tag.csv $CMD > gpurun_out/ncu_launch_full_$tag.log 2>&1
echo launches_full rc=$?
CMD="python bench.py --loci 30000 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain30k_$tag.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:LockCfgILi8E -s 3 -c 1 -f -o gpurun_out/prof_quad_$tag $CMD > gpurun_out/ncu_quad_$tag.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:LockCfgILi16E -s 3 -c 1 -f -o gpurun_out/prof_duo_$tag $CMD > gpurun_out/ncu_duo_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_align_thread -s 3 -c 1 -f -o gpurun_out/prof_thread_$tag $CMD > gpurun_out/ncu_thread_$tag.log 2>&1
echo full rc=$?
