for n in 2 3 4; do python bench.py --steps 60 --warmup 5 --no-cpu-baseline --e2e-lanes $n 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('lanes', d['e2e']['in_flight'], 'e2e %.0f serial %.0f value %.0f'%(d['e2e']['value'], d['e2e']['serial_value'], d['value']))"; done
