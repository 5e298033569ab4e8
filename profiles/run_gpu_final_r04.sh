#!/bin/bash
# End-of-round capture (third session of round 2) on one B200: parity tests, smoke, both bench arms at the judged
# configuration (937k loci), per-launch device times of the judged command, BASELINE configs 1/3/4/5 with the heuristic
# census, and --set full captures of the top kernels (incl. the long tier on config 4).
tag=${1:-r04z}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader | head -1; nproc
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$tag.json 2>/dev/null
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full_$tag.json 2> gpurun_out/bench_full_$tag.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_full_$tag.json").read().strip().splitlines()[-1]); r=json.loads(open("gpurun_out/bench_ref_$tag.json").read().strip().splitlines()[-1])
print("value",round(d["value"]),"ms/step",d["ms_per_step"],"e2e",round(d["e2e"]["value"]),"e2e_packed",round(d["e2e_packed"]["value"]),"ref arm",round(r["value"]),"cpu",round(d["cpu_baseline"]["value"]))
print(d["breakdown_ms_per_step"]); print("roofline",d["roofline"]["kernel"],d["roofline"]["frac"],d["roofline"]["all_alignment_tiers"]); print(d["clocks"], d["self_check_vs_oracle"])
PY
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-census"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_full_$tag.csv $CMD > gpurun_out/ncu_launch_full_$tag.log 2>&1
echo launches_full rc=$?
timeout 1500 python profiles/bench_configs.py > gpurun_out/configs_$tag.jsonl 2> gpurun_out/configs_$tag.err; echo "configs rc=$?"; tail -2 gpurun_out/configs_$tag.err
CMD="python bench.py --loci 100000 --steps 1 --warmup 3 --no-cpu-baseline --no-census"
for k in "k_align_lock.*Li8ELi64" "k_align_lock.*Li8ELi32" k_align_thread k_dedup_window; do
  name=$(echo $k | tr -c 'A-Za-z0-9_\n' '_')
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:$k -s 3 -c 1 -f \
      -o gpurun_out/prof_${name}_$tag $CMD > gpurun_out/ncu_${name}_$tag.log 2>&1
  echo "$k rc=$?"
done
CMD="python profiles/bench_configs.py --long-loci 300 --skip 135 --census-long-loci 0"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_align_lock.*Li32ELi192 -s 1 -c 1 -f \
    -o gpurun_out/prof_k_align_lock__Li32ELi192_$tag $CMD > gpurun_out/ncu_long_$tag.log 2>&1
echo "long rc=$?"
