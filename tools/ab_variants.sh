#!/bin/bash
# A/B of prebuilt library variants (variants/lib_*.so, built here with different flags or sources):
# each is copied over the in-tree library on the GPU box and measured with the same two commands.
for v in "$@"; do
  cp variants/lib_$v.so canvas_ity_b200/libcanvas_b200.so
  echo "== variant $v"
  python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('tiger fps %.1f single %.1f comp_kernel_ms %.4f frac %.3f' % (d['value'], d['single_canvas']['value'], d['roofline']['kernel_ms'], d['roofline']['frac']), d['stages_ms'], 'cfg3', {k: round(v['ms'],3) for k,v in d['passes']['config3_tiger_alpha0.9_shadow_blur16'].items() if isinstance(v,dict)}, round(d['passes']['config3_tiger_alpha0.9_shadow_blur16']['frames_per_s'],1))"
  FILL_SIZE=8192 python tools/fill_bench.py 2>&1 | grep -E "solid   source_over|linear  source_over"
done
