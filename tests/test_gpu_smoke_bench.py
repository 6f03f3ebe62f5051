"""GPU checks of the driver entry points: __graft_entry__.smoke() and a tiny bench.py run (JSON contract)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_smoke():
    import __graft_entry__ as ge
    ge.smoke()


def test_bench_json_contract():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3", "--scenarios", "64",
                          "--e2e-scenarios", "16", "--e2e-steps", "3", "--cpu-seconds", "1", "--no-extra"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in line, k
    assert line["value"] > 0 and line["gpu_launches"] > 0 and line["dtype"] == "f64" and line["vs_baseline"] is None
    assert line["roofline"]["bound"] == "hbm" and 0 < line["roofline"]["frac"] < 1.5
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["scaling"] == "strong" and set(line["callbacks"]) == {"g", "g+J"} and line["config"]["scenarios_total"] == 64
    # the reference arm describes the same workload with the same keys (the driver compares the two config dicts)
    ref = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--scenarios", "64"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert ref.returncode == 0, ref.stderr[-2000:]
    assert json.loads(ref.stdout.strip().splitlines()[-1])["config"] == line["config"]


def test_reference_arm_contract():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["cpu_baseline"]["kind"] == "port"
