#!/bin/bash
# round 2 quick GPU iteration: parity tests + the full-size bench line
# usage: run_gpu_r02_quick.sh <tag> [loci] [pytest-selector]
TAG=${1:-x}; LOCI=${2:-937000}; SEL=${3:-tests}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv,noheader | head -2
python - <<'PY'
import torch, time
a = torch.empty(1 << 30, dtype=torch.uint8).pin_memory(); b = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
for _ in range(2):
    torch.cuda.synchronize(); t = time.time(); b.copy_(a, non_blocking=True); torch.cuda.synchronize(); dt = time.time() - t
print(f"H2D pinned 1 GiB: {1.0737 / dt:.1f} GB/s")
PY
nproc; free -g | head -2
timeout 1500 python -m pytest $SEL -q -m gpu -x 2>&1 | tail -15
timeout 900 python bench.py --loci $LOCI --steps 5 --warmup 3 2>gpurun_out/bench_$TAG.err | tee gpurun_out/bench_$TAG.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value',round(d['value']),'ms',d['ms_per_step'],'e2e',round(d['e2e']['value']),d['e2e']['ms_per_step'],'e2e_packed',round(d['e2e_packed']['value']),d['e2e_packed']['ms_per_step'],'cpu',round(d['cpu_baseline']['value']) if d['cpu_baseline'] else None)
print(d['breakdown_ms_per_step'])
print('e2e_host',d['e2e_host']); print('e2e_packed_host',d['e2e_packed_host'])
print('thread',d['thread_tier_pairs'],'tiers',d['tier_pairs'],'ident',d['pairs_identical'],'dup',d['pairs_duplicate'],'closed',d['pairs_closed_form'],'failed',d['pairs_failed'])
print('roofline',d['roofline']['kernel'],d['roofline']['frac'],'all tiers',d['roofline']['all_alignment_tiers'])
print('census',d['heuristic_census']); print('check',d['self_check_vs_oracle'])
"
tail -5 gpurun_out/bench_$TAG.err
