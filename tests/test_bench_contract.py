"""The JSON line bench.py prints is a contract with the driver: check its shape on a small canvas.
CPU: the reference arm (and its oracle-port fallback).  GPU: our arm."""
import json
import os
import subprocess
import sys

import pytest

from tests import harness as H

BENCH = os.path.join(H.ROOT, "bench.py")
COMMON = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
          "vs_baseline", "dtype", "data", "config", "e2e"}


def run_bench(*flags, env=None):
    out = subprocess.run([sys.executable, BENCH, *flags], capture_output=True, text=True, timeout=900,
                         env=dict(os.environ, **(env or {})))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, "bench.py must print exactly one JSON line"
    return json.loads(lines[0])


def test_reference_arm_line():
    line = run_bench("--impl", "reference", "--size", "200", "--steps", "3", "--warmup", "2")
    assert COMMON <= set(line) and line["impl"] == "reference"
    assert line["steps"] == 3 and line["warmup"] == 2                       # the arm honours --steps / --warmup
    sys.path.insert(0, H.ROOT)
    import bench
    assert line["config"]["workload"] == bench.workload_name(200)           # the same string our arm prints
    assert line["value"] > 0 and line["unit"] == "frames/s" and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


def test_reference_arm_falls_back_to_the_oracle_port(monkeypatch):
    """Where oracle/_ref was never built the arm times the oracle port instead of reporting nothing."""
    sys.path.insert(0, H.ROOT)
    import bench
    monkeypatch.setattr(H, "reference_library", lambda fast=False: None)
    kind, frame = bench.cpu_renderer(96)
    assert kind == "port"
    frame()


@pytest.mark.gpu
def test_our_arm_line():
    line = run_bench("--size", "1024", "--steps", "6", "--warmup", "3", "--lanes", "2", "--no-cpu-baseline", "--skip-configs")
    assert COMMON <= set(line) and "impl" not in line
    assert line["value"] > 0 and line["n_gpus"] == 1 and line["steps"] == 6 and line["scaling"] == "weak"
    assert line["config"]["frames_per_step"] == 2 and line["gpu_launches"] >= 6 * 2 * 30
    assert line["timed"]["regions"] >= 3 and (line["timed"]["seconds_timed"] >= 0.5 or line["timed"]["regions"] == 64)
    assert line["span_ms"]["min"] <= line["span_ms"]["median"] <= line["span_ms"]["max"]
    assert abs(line["value"] - 6 * 2 / (line["span_ms"]["median"] * 1e-3)) < 1e-6 * line["value"]
    assert abs(line["ms_per_step"] - line["span_ms"]["median"] / 6) < 1e-9
    roof = line["roofline"]
    assert roof["bound"] == "hbm" and roof["unit"] == "GB/s" and roof["peak"] > 1000
    assert abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-9 and roof["achieved"] > 0
    e2e = line["e2e"]
    assert e2e["value"] > 0 and e2e["d2h_bytes_per_frame"] == 1024 * 1024 * 4 and e2e["h2d_bytes_per_frame"] > 10000
    assert e2e["d2h_bytes_per_step"] == e2e["frames_per_step"] * e2e["d2h_bytes_per_frame"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(line["clocks"])
    assert line["config"]["workload"].startswith("tiger_1024")
    assert line["passes"]["readback"]["ms"] > 0 and line["single_canvas"]["value"] > 0
    assert line["passes"]["png_encode"]["ms"] > 0 and line["passes"]["hit_test"]["rule_evaluations_per_s"] > 0
