"""The C++17 host API (lidar-slam-from-scratch_b200/host/slam_viz/core/*.hpp: the reference's class and function names
on top of the C ABI) compiles with the in-tree toolchain, refuses to run without a GPU, and — on a B200 — reproduces
the oracle when driven like slam_node.cpp drives the reference (tests/cpp/host_api_test.cpp)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "host_api_test")
EXE_EIGEN_API = EXE + "_eigenapi"  # the headers' Eigen branch, compiled against oracle/eigen_standin


def build():
    import oracle_lib
    oracle_lib.Oracle(), oracle_lib.Synth()  # builds liboracle.so / libsynth.so if needed
    import sys
    sys.path.insert(0, os.path.join(ROOT, "lidar-slam-from-scratch_b200", "python"))
    import slam_b200
    slam_b200.load_library()
    subprocess.check_call(["sh", os.path.join(ROOT, "tests", "cpp", "build.sh")])


def test_cpp_host_api_compiles_and_has_no_cpu_fallback():
    build()
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    for exe in (EXE, EXE_EIGEN_API):
        p = subprocess.run([exe], capture_output=True, text=True)
        assert p.returncode != 0
        assert "no usable sm_100a device" in p.stderr


@pytest.mark.gpu
def test_cpp_host_api_matches_oracle():
    build()
    for exe in (EXE, EXE_EIGEN_API):
        p = subprocess.run([exe], capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, exe + "\n" + p.stdout + p.stderr
        assert "all checks passed" in p.stdout
