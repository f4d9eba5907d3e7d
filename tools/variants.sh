for v in eager lazy; do
  if [ $v = lazy ]; then export CB200_LAZY_LOAD=1; else unset CB200_LAZY_LOAD; fi
  echo "== $v"; python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['stages_ms'], d['roofline']['kernel_ms'], d['e2e'])"
done
