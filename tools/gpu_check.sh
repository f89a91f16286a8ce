#!/bin/bash
# dev tool: GPU parity tests + both bench workloads (short), prints value / e2e / kernel ms / roofline fraction
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in mhpc hkd; do
python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('$w', round(d['value'],1), round(d['e2e']['value'],1), d['roofline']['frac'], d['roofline'].get('avg_launch_ms'), d['roofline'].get('kernel_ms'))"
done
