#!/bin/bash
# strong-scaling runs of bench.py on the GPUs of one box: run_gpu_scale.sh <tag> <N> [<N> ...]
TAG=$1; shift
nvidia-smi --query-gpu=index,name,pcie.link.gen.current,pcie.link.width.current --format=csv,noheader
nproc; free -g | head -2; lscpu | grep -i "numa\|socket\|model name" | head -6
for N in "$@"; do
  if [ "$N" = "1" ]; then python profiles/h2d_probe.py 2>/dev/null | tail -1;
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 profiles/h2d_probe.py 2>/dev/null | tail -1; fi
  if [ "$N" = "1" ]; then CMD="python bench.py --gpus 1 --steps 5 --warmup 3";
  else CMD="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3"; fi
  timeout 1200 $CMD > gpurun_out/scale_${TAG}_$N.json 2> gpurun_out/scale_${TAG}_$N.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/scale_${TAG}_$N.json") if l.startswith("{")][-1])
    print("N=$N value %.2f M loci/s (%.2f ms)  e2e %.2f M (%.1f ms)  e2e_packed %.2f M (%.1f ms)" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"], d["e2e_packed"]["value"]/1e6, d["e2e_packed"]["ms_per_step"]))
    for r in d["per_rank"]: print("   ", r)
except Exception as e:
    print("N=$N failed", e); print(open("gpurun_out/scale_${TAG}_$N.err").read()[-1500:])
PY
done
