#!/bin/bash
# A/B of the bulk-copy (TMA) build of the lean compositor against the plain LDG/STG build: parity tests with the
# switch on, then the 8192^2 solid fills and the tiger bench with the switch off and on.  Output: gpurun_out/tma_ab/
O=gpurun_out/tma_ab; mkdir -p $O
CB200_TMA=1 timeout 300 python -m pytest tests -q -m gpu -x -k "reference_suite or tiger_on_gpu or full_size_properties_4096 or tiger_4096_against" 2>&1 | tail -3 | tee $O/pytest_tma.txt
for v in 0 1; do
  CB200_TMA=$v FILL_SIZE=8192 FILL_KINDS=solid python tools/fill_bench.py 2>&1 | tail -4 | sed "s/^/TMA=$v  /" | tee -a $O/fills.txt
  CB200_TMA=$v python bench.py --steps 20 --warmup 5 --no-cpu-baseline --skip-configs > $O/bench_tma$v.json 2> $O/bench_tma$v.err
  python - <<P | tee -a $O/tiger.txt
import json
d=json.load(open("$O/bench_tma$v.json"))
print("TMA=$v tiger 4096: %.1f frames/s (8 lanes), single canvas %.1f, k_composite %.4f ms, stages %s" % (d["value"], d["single_canvas"]["value"], d["roofline"]["kernel_ms"], d["stages_ms"]))
P
done
for v in 0 1; do
  CB200_TMA=$v FILL_SIZE=8192 FILL_KINDS=solid FILL_OPS=exclusive_or timeout 200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed_op_global_ld.sum,smsp__inst_executed_op_global_st.sum,launch__registers_per_thread --clock-control none -k regex:k_composite -s 5 -c 1 python tools/fill_bench.py 2>&1 | grep -E "k_composite|gpu__time|inst_executed|dram__|issue_active|warps_active|registers" | sed "s/^/TMA=$v /" | tee -a $O/ncu_fill.txt
done
