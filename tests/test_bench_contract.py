"""bench.py's output contract (driver-facing): the reference arm runs on CPU, the product arm needs a B200."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e"}


def _run(*args, timeout=600):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                         timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    return json.loads(lines[0])


def test_reference_arm_prints_one_json_line():
    """`--impl reference`: the CPU restatement of the reference on a bounded sample, same metric/unit/config."""
    line = _run("--impl", "reference", "--steps", "1", "--warmup", "1")
    assert BASE_KEYS <= set(line)
    assert line["impl"] == "reference" and line["metric"] == "images_per_sec" and line["unit"] == "images/s"
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["value"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "sample" in cb
    assert "workload" in line["config"]


@pytest.mark.gpu
def test_product_arm_line_has_roofline_and_launch_count():
    line = _run("--steps", "3", "--warmup", "3", "--no-cpu-baseline")
    assert BASE_KEYS <= set(line) and "impl" not in line
    assert line["gpu_launches"] > 0 and line["n_gpus"] == 1 and line["dtype"] == "fp16"
    e2e = line["e2e"]
    assert e2e["h2d_bytes_per_step"] == 256 * 224 * 224 * 3 and e2e["d2h_bytes_per_step"] > 0 and e2e["value"] > 0
    roof = line["roofline"]
    assert roof["bound"] in ("tensor", "hbm") and 0 < roof["frac"] < 1 and roof["unit"] == "TFLOP/s"
    assert abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-9
    assert set(line["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}


def test_reference_arm_does_not_map_the_product_library():
    """the reference arm must not load libroomnet.so (it times the CPU restatement only)."""
    code = (
        "import sys, runpy\n"
        "sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '1']\n"
        "try:\n"
        "    runpy.run_path(%r, run_name='__main__')\n"
        "except SystemExit:\n"
        "    pass\n"
        "print('MAPPED', 'libroomnet' in open('/proc/self/maps').read())\n" % os.path.join(ROOT, "bench.py"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert "MAPPED False" in out.stdout, out.stdout[-300:] + out.stderr[-300:]
