#!/bin/bash
# Everything profiles/ is built from, in one GPU round trip (outputs under gpurun_out/round/).
O=gpurun_out/round; mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > $O/pytest_gpu.txt; cat $O/pytest_gpu.txt
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; tail -c 600 $O/bench.json; tail -3 $O/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; tail -c 700 $O/bench_reference.json
FILL_SIZE=8192 timeout 600 python tools/fill_bench.py > $O/fill_8192.txt 2>&1; tail -12 $O/fill_8192.txt
timeout 600 python tools/batch_bench.py 2048 > $O/batch.txt 2>&1; tail -4 $O/batch.txt
timeout 300 python tools/text_bench.py > $O/text_bench.json 2>&1; tail -30 $O/text_bench.json
timeout 300 python tools/png_bench.py > $O/png_bench.json 2>&1; cat $O/png_bench.json
timeout 300 python tools/hit_bench.py > $O/hit_bench.json 2>&1; cat $O/hit_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_tiger4096.csv python bench.py --steps 2 --warmup 3 --lanes 1 --no-cpu-baseline --skip-configs > $O/ncu_launches.log 2>&1
for k in k_composite k_readback k_rows k_sort_scatter; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 2 -f -o $O/prof_$k python bench.py --steps 2 --warmup 3 --lanes 1 --no-cpu-baseline --skip-configs > $O/prof_$k.log 2>&1
done
for k in k_png_rows k_hit_test k_glyph_instances; do      # launched a handful of times only (bench passes / text bench)
  cmd="bench.py --steps 2 --warmup 3 --lanes 1 --no-cpu-baseline --skip-configs"; [ $k = k_glyph_instances ] && cmd="tools/text_bench.py 2000"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o $O/prof_$k python $cmd > $O/prof_$k.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_blur_x|k_blur_y" -s 6 -c 2 -f -o $O/prof_shadow python tools/shadow_bench.py > $O/prof_shadow.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_composite" -s 3 -c 1 -f -o $O/prof_composite_config3 python tools/shadow_bench.py > $O/prof_composite_config3.log 2>&1
FILL_KINDS=solid FILL_OPS=exclusive_or timeout 600 ncu --set full --clock-control none -k regex:k_composite -s 5 -c 1 -f -o $O/prof_fill python tools/fill_bench.py > $O/prof_fill.log 2>&1
for k in linear image; do
  FILL_KINDS=$k FILL_OPS=exclusive_or timeout 600 ncu --set full --clock-control none -k regex:k_composite -s 5 -c 1 -f -o $O/prof_fill_$k python tools/fill_bench.py > $O/prof_fill_$k.log 2>&1
done
ls -la $O
