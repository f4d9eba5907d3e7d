#!/bin/bash
# quick A/B numbers: tiger 4096 stage times + 8192 solid/linear fill
python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('tiger fps %.1f'%d['value'], d['stages_ms'], 'e2e %.0f/%.0f'%(d['e2e']['value'], d['e2e']['serial_value']))"
FILL_SIZE=8192 python tools/fill_bench.py 2>&1 | grep -E "solid   source_over|linear  source_over|image   source_over"
