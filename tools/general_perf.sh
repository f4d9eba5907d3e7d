#!/bin/bash
# general-compositor A/B: 8192^2 gradient / image fills + config 3
FILL_SIZE=8192 FILL_OPS=source_over python tools/fill_bench.py 2>&1 | grep -E "composite"
python tools/shadow_bench.py | tail -2
