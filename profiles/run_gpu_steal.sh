#!/bin/bash
# e2e (ASCII hand-over) with and without the GPU taking raw blocks over from the host packer (one GPU, 15 packer threads)
for v in "" "TDB_STEAL_AUTO=1" "TDB_RAW_EVERY=4"; do
  env $v python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-census > gpurun_out/steal.json 2>gpurun_out/steal.err || tail -3 gpurun_out/steal.err
  python -c "
import json
d=json.loads(open('gpurun_out/steal.json').read().strip().splitlines()[-1])
print('[$v]', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e']['ms_per_step'], 'h2d', d['e2e']['h2d_bytes_per_step'], 'raw', d['e2e_host'].get('ascii_bytes_taken_over'), 'pack_wait', d['e2e_host']['ms_pack_wait'], 'last_packed', d['e2e_host']['ms_last_block_packed'], 'e2e_packed', round(d['e2e_packed']['value']))"
done
