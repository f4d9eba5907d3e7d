#!/bin/bash
# bench.py at N GPUs the way the driver launches it: weak (independent frames), bands (+ all_gather), reference arm
N=${1:-2}; O=gpurun_out/round; mkdir -p $O
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@"; }
run --steps 40 --warmup 5 --no-cpu-baseline 2> $O/n${N}_weak.err | tail -1 > $O/n${N}_weak.json
run --steps 40 --warmup 5 --no-cpu-baseline --mode bands 2> $O/n${N}_bands.err | tail -1 > $O/n${N}_bands.json
run --impl reference --steps 1 --warmup 1 2> $O/n${N}_ref.err | tail -1 > $O/n${N}_ref.json
for f in weak bands ref; do python - $O/n${N}_$f.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], 'value %.1f'%d['value'], d.get('scaling'), d['config'].get('parallelism'), 'e2e', d.get('e2e',{}).get('value'))
except Exception as e: print(sys.argv[1], 'unreadable', e)
PY
done; for f in $O/n${N}_*.err; do tail -n 3 $f; done
