#!/bin/bash
for d in 4 2; do
timeout 200 python bench.py --steps 40 --warmup 6 --no-cpu-baseline --depth $d 2>/dev/null | tail -1 | python -c "
import sys,json
r=json.loads(sys.stdin.read()); print('depth=$d', 'ms', r['ms_per_step'], 'fps', r['value'], 'e2e', r['e2e']['value'], 'seq', r['sequential']['ms_per_step'], r['clocks'])"
done
