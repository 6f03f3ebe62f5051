"""CPU check of the warp-tile tables the plan builder hands to bus_warp_kernel (tests/native/plan_check.cpp):
every contribution has exactly one destination, the multi-slot lists reproduce the left-to-right slot sums of the
slot -> contribution maps, the tiles cover the whole bus block, a bus with more than 32 half-edges disables the path."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_warp_tile_tables(tmp_path):
    exe = str(tmp_path / "plan_check")
    csrc = os.path.join(ROOT, "powersense.jl_b200", "csrc")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", csrc, "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "native", "plan_check.cpp"), os.path.join(csrc, "ps_plan.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.strip().startswith("ok")
