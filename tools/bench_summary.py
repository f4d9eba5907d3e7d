#!/usr/bin/env python
"""Print the figures of a bench.py JSON line that the kernel work watches."""
import json, sys
b = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print({k: b[k] for k in ("value", "span_ms", "single_canvas", "stages_ms")})
print("k_composite ms", b["roofline"]["kernel_ms"], "frac", b["roofline"]["frac"], "e2e", b["e2e"]["value"], b["e2e"]["serial_value"])
p = b.get("passes", {})
for k in ("readback", "coverage", "sort"):
    if k in p: print(k, p[k])
c3 = p.get("config3_tiger_alpha0.9_shadow_blur16")
if c3: print("config3 frame", c3["frame_ms"], "blur + raster", c3["blur_and_shadow_raster"]["ms"], "composite", c3["composite"]["ms"], "frac of ceiling", c3["frac_of_per_pass_hbm_ceiling"])
c4 = p.get("config4_full_canvas_fills")
if c4 and "fills" in c4: print("config4", {k: round(v["composite_ms"], 3) for k, v in c4["fills"].items()})
if "config5_batch_of_256x256_canvases" in p: print("config5", p["config5_batch_of_256x256_canvases"])
if "config3_one_frame_in_scanline_bands" in p: print("bands", p["config3_one_frame_in_scanline_bands"])
