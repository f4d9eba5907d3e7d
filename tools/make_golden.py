#!/usr/bin/env python
"""Regenerate tests/golden/ from the reference (needs /root/reference; run here, not on the GPU box).

1. builds oracle/_ref/capture_tests and capture_tiger (reference drivers compiled against
   oracle/capture_shim.hpp, sources read from /root/reference where they lie),
2. runs them: per reference test (test/test.cpp:2185-2262) the API call stream (.cvs) and the
   reference's own RGBA8 output; for demos/tiger/tiger.cpp the call stream of one frame,
3. renders the tiger at 512x512 with the reference (config 1 of BASELINE.json) for a golden,
4. writes tests/golden/{scripts/*.cvs, reference_rgba8.npz, manifest.json}.
"""
import ctypes as C, json, os, subprocess, sys, tempfile
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref", "capture"])
    out = tempfile.mkdtemp(prefix="cb200_golden_")
    subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "capture_tests"), out])
    subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "capture_tiger"), os.path.join(out, "tiger.cvs")],
                          stdout=subprocess.DEVNULL)
    os.makedirs(os.path.join(GOLD, "scripts"), exist_ok=True)
    manifest = {"tests": [], "source": "a-e-k/canvas_ity test/test.cpp + demos/tiger/tiger.cpp, captured by tools/make_golden.py"}
    images = {}
    for line in open(os.path.join(out, "manifest.txt")):
        name, expected, got, w, h = line.split()
        assert expected == got, (name, expected, got)
        manifest["tests"].append({"name": name, "hash": expected, "width": int(w), "height": int(h)})
        images[name] = np.fromfile(os.path.join(out, name + ".rgba8"), np.uint8).reshape(int(h), int(w), 4)
        with open(os.path.join(out, name + ".cvs"), "rb") as f, open(os.path.join(GOLD, "scripts", name + ".cvs"), "wb") as g:
            g.write(f.read())
    with open(os.path.join(out, "tiger.cvs"), "rb") as f, open(os.path.join(GOLD, "scripts", "tiger.cvs"), "wb") as g:
        g.write(f.read())
    # tiger goldens rendered by the reference build itself
    sys.path.insert(0, ROOT)
    from tests import harness
    ref = harness.reference_library()
    for size in (512,):
        script = harness.tiger_script(size, size)
        images["tiger_%d" % size] = harness.render_script(ref, script, size, size)["rgba8"]
    manifest["tiger"] = {"native_size": [733, 757], "draws": 305, "goldens": ["tiger_512"]}
    np.savez_compressed(os.path.join(GOLD, "reference_rgba8.npz"), **images)
    json.dump(manifest, open(os.path.join(GOLD, "manifest.json"), "w"), indent=1)
    print("wrote", GOLD)


if __name__ == "__main__":
    main()
