timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -4; timeout 600 python bench.py --loci 100000 --steps 3 --warmup 3 --cpu-sample-loci 50000 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'cpu',round(d['cpu_baseline']['value']))
print(d['breakdown_ms_per_step'], d['tier_pairs'], d['roofline']['int_issue']['frac'])
"
