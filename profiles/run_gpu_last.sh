#!/bin/bash
# last check of the committed code: all GPU tests, smoke, both bench arms at the judged configuration
tag=${1:-r04y}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$tag.json 2>/dev/null
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full_$tag.json 2> gpurun_out/bench_full_$tag.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_full_$tag.json").read().strip().splitlines()[-1]); r=json.loads(open("gpurun_out/bench_ref_$tag.json").read().strip().splitlines()[-1])
print("value",round(d["value"]),"ms/step",d["ms_per_step"],"e2e",round(d["e2e"]["value"]),"e2e_packed",round(d["e2e_packed"]["value"]),"ref arm",round(r["value"]),"cpu",round(d["cpu_baseline"]["value"]))
print(d["breakdown_ms_per_step"]); print("roofline",d["roofline"]["kernel"],d["roofline"]["frac"],d["roofline"]["all_alignment_tiers"]); print(d["clocks"], d["self_check_vs_oracle"], d["heuristic_census"])
PY
