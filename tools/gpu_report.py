#!/usr/bin/env python
"""Developer report (GPU box): every golden script through the CUDA back end vs the oracle.

Prints one line per test: float mismatches vs the oracle (same lowered frames), 8-bit
difference vs the committed reference output, hash distance vs the reference's expected hash.
"""
import sys, os, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import harness as H


def main():
    names = sys.argv[1:] or [t["name"] for t in H.manifest()["tests"]] + ["tiger_512"]
    lib = H.product_library()
    print("devices:", lib.cb200_device_count())
    fails = 0
    for name in names:
        if name.startswith("tiger_"):
            size = int(name.split("_")[1])
            script, w, h, expect = H.tiger_script(size, size), size, size, None
        else:
            t = [t for t in H.manifest()["tests"] if t["name"] == name][0]
            script, w, h, expect = H.golden_script(name), t["width"], t["height"], int(t["hash"], 16)
        try:
            t0 = time.time()
            got = H.render_script(lib, script, w, h)
            dt = time.time() - t0
            want = H.render_oracle(script, w, h)
            nbad, maxabs = H.float_mismatch(got["f32"], want["f32"])
            gold = H.golden_rgba8(name) if name in np.load(os.path.join(H.GOLD, "reference_rgba8.npz")).files else want["rgba8"]
            da, dc, n8 = H.rgba8_mismatch(got["rgba8"], gold)
            hd = H.hamming(H.hash_image(got["rgba8"]), expect) if expect is not None else -1
            qbad = sum(1 for q in got["queries"] if q[1] != q[2] and q[0] != H.OP["GET_IMAGE_DATA"])
            ok = nbad == 0 and n8 == 0 and hd <= 5
            fails += 0 if ok else 1
            print("%-28s %s float>tol=%-7d max|d|=%.3e  d8: a=%d c=%.2f n>1=%-6d hash_d=%-3d qbad=%d  %.0f ms" %
                  (name, "ok  " if ok else "FAIL", nbad, maxabs, da, dc, n8, hd, qbad, dt * 1e3), flush=True)
        except Exception as e:
            fails += 1
            print("%-28s ERROR %s" % (name, e), flush=True)
            traceback.print_exc()
    print("failures:", fails)


if __name__ == "__main__":
    main()
