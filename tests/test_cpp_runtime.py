"""Compiles and runs the C++ scheduler tests (tests/cpp/test_runtime.cpp), which
re-express aby3_tests/Sh3RuntimeTests.cpp against the facade's Sh3Runtime.  Host only."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_runtime_schedule_orders(tmp_path):
    exe = str(tmp_path / "test_runtime")
    pkg = os.path.join(ROOT, "aby3_b200")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-pthread", "-I", ROOT, "-I", os.path.join(ROOT, "include"),
                           "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_runtime.cpp"),
                           "-L", pkg, "-lsh3", "-laby3cu", "-Wl,-rpath," + pkg])
    out = subprocess.run([exe], stdout=subprocess.PIPE, text=True)
    assert out.returncode == 0, out.stdout
    assert "ALL OK" in out.stdout
