"""The work decomposition of the exhaustive launch (pipsort_b200/csrc/exh_plan.h) on the CPU: tests/cpp/exh_plan_check.cpp
replays every chunk of a plan with the walk the kernel performs and checks that the chunks tile the warp-step space
exactly once and that every union subset is visited by exactly one lane -- for chunk targets from 1 step to 'everything',
U around the tile boundaries, whole and partial first-SNP ranges."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_chunks_tile_the_step_space():
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, "exh_plan_check")
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "cpp", "exh_plan_check.cpp")])
        p = subprocess.run([exe], capture_output=True, text=True)
        assert p.returncode == 0, p.stdout + p.stderr
        assert p.stdout.startswith("ok")
