"""The host side in the reference's own language: include/loop_closing_system_b200.hpp (loop_closing::b200::LoopClosingSystem,
the interface of the reference's include/loop_closing.hpp:29-66) compiled with g++ against the C ABI and the stand-in OpenCV
types of tests/stubs, and driven by tests/cpp/system_check.cpp, which checks every frame's results against the C oracle.
Without a GPU the binary must report lc_create's "no CPU fallback" and exit 0; on a B200 it must PASS."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    import slam_loop_closing_b200 as m
    from oracle import c_oracle
    if not shutil.which("g++"):
        pytest.skip("g++ not found")
    lib, ora = m.library_path(), c_oracle.build()
    exe = str(tmp_path / "system_check")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "tests", "stubs"), "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "cpp", "system_check.cpp"), lib, ora,
                           "-Wl,-rpath," + os.path.dirname(lib), "-Wl,-rpath," + os.path.dirname(ora), "-o", exe])
    return exe


def _has_gpu():
    import torch
    return torch.cuda.is_available()


@pytest.mark.skipif(_has_gpu(), reason="the GPU tier runs the real thing")
def test_cpp_system_builds_and_fails_loudly_without_a_gpu(tmp_path):
    r = subprocess.run([_build(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "no CPU fallback" in r.stdout, r.stdout


@pytest.mark.gpu
def test_cpp_system_equals_the_oracle_on_the_gpu(tmp_path):
    out = tmp_path / "results"
    out.mkdir()
    r = subprocess.run([_build(tmp_path)], capture_output=True, text=True, timeout=600, env=dict(os.environ, LC_TEST_OUT=str(out)))
    assert r.returncode == 0 and r.stdout.startswith("PASS"), r.stdout + r.stderr
    txt = (out / "loop_closures.txt").read_text()  # saveResults (include/loop_closing.hpp:66, README.md:150-165)
    assert txt.startswith("Total frames processed: 64\nLoop closures detected: ") and "Frame " in txt and "  Similarity: " in txt
