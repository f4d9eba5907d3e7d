#!/usr/bin/env python
"""Config 5 with several batches in flight on one GPU (bench.py's config5_in_flight) for 1, 2, 4 and 8 batches."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from tests import harness as H
from canvas_ity_b200 import _native

lib = _native.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
scripts = [H.config5_script(i) for i in range(n)]
for parts in [int(a) for a in (sys.argv[2:] or ["1", "2", "4", "8"])]:
    r = bench.config5_in_flight(lib, scripts, n, 1, 0, torch.cuda.synchronize, torch, None, parts=parts)
    print(json.dumps(r), flush=True)
