#!/bin/bash
# ncu --set full capture of one kernel (regex $1) from a short bench run; report to gpurun_out/prof_$2
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$1 -s ${SKIP:-4} -c ${COUNT:-2} -f -o gpurun_out/prof_$2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof_$2.log 2>&1
tail -2 gpurun_out/prof_$2.log | cut -c1-300
ls -la gpurun_out/prof_$2.ncu-rep
