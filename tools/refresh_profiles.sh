#!/bin/bash
# profiles/rNN_* (argument; default r01) from the outputs of tools/gpu_round.sh (gpurun_out/round/), read here on the CPU box.
R=gpurun_out/round; P=profiles; N=${1:-r01}
cp $R/bench.json $P/${N}_bench_tiger4096.json; cp $R/bench_reference.json $P/${N}_bench_reference_arm.json
cp $R/fill_8192.txt $P/${N}_fill_bench_8192.txt; [ -f gpurun_out/fill_bench.json ] && cp gpurun_out/fill_bench.json $P/${N}_fill_bench_8192.json
tail -4 $R/batch.txt > $P/${N}_batch_2048x256.txt
cp $R/text_bench.json $P/${N}_text_instancing.json; cp $R/png_bench.json $P/${N}_png_encode.json; cp $R/hit_bench.json $P/${N}_hit_test.json
cp $R/launches_tiger4096.csv $P/${N}_launches_tiger4096.csv
python tools/launch_summary.py $R/launches_tiger4096.csv > $P/${N}_launches_tiger4096_summary.txt
python tools/ncu_summary.py $R/prof_k_composite.ncu-rep > $P/${N}_ncu_composite.txt
python tools/ncu_summary.py $R/prof_k_readback.ncu-rep > $P/${N}_ncu_readback.txt
python tools/ncu_summary.py $R/prof_shadow.ncu-rep > $P/${N}_ncu_shadow_blur.txt
python tools/ncu_summary.py $R/prof_k_png_rows.ncu-rep > $P/${N}_ncu_png.txt
(for k in k_rows k_sort_scatter k_hit_test k_glyph_instances; do python tools/ncu_summary.py $R/prof_$k.ncu-rep; done) > $P/${N}_ncu_small_kernels.txt
rm -f $P/${N}_ncu_traffic.json
python tools/ncu_traffic.py $P/${N}_ncu_traffic.json k_composite=$R/prof_k_composite.ncu-rep k_readback=$R/prof_k_readback.ncu-rep \
  k_blur_x=$R/prof_shadow.ncu-rep k_blur_y=$R/prof_shadow.ncu-rep \
  k_composite_fill=$R/prof_fill.ncu-rep k_png_rows=$R/prof_k_png_rows.ncu-rep k_hit_test=$R/prof_k_hit_test.ncu-rep > /dev/null
ls $P
