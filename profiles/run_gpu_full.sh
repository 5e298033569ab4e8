#!/bin/bash
# judged configuration: full 937k-locus trio, both arms, + launch list and one full capture per top kernel
tag=${1:-r01d}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$tag.json 2>/dev/null; head -c 600 gpurun_out/bench_ref_$tag.json; echo
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_full_$tag.json 2> gpurun_out/bench_full_$tag.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_full_$tag.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_full_$tag.json"))
print("value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"cpu",round(d["cpu_baseline"]["value"]),"ms/step",d["ms_per_step"])
print(d["breakdown_ms_per_step"]); print(d["e2e_host"], d["e2e_chunks"]); print(d["roofline"]["int_issue"]["frac"], d["roofline"]["kernel"], d["clocks"])
PY
CMD="python bench.py --loci 100000 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_$tag.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_launch_$tag.log 2>&1
echo launches rc=$?
