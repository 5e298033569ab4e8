#!/bin/bash
# A/B on one box: e2e with the dense layout (seq_off = NULL) against explicit offsets, full configuration
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
for rep in 1 2; do
for d in 1 0; do
TDB_BENCH_DENSE=$d timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/ab_$d.err > gpurun_out/ab_${d}_$rep.json
python - <<PY
import json
d=json.load(open("gpurun_out/ab_${d}_$rep.json"))
print("dense=$d rep=$rep value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "h2d", d["e2e"]["h2d_bytes_per_step"], d["e2e_host"]["ms_align_call"], d["e2e_host"]["ms_last_block_packed"])
PY
done
done
