"""The C++ mirror of the reference's reader API (go-sdr_b200/host/hzsdr.hpp): it must build against
the C ABI on CPU, and its re-expression of the reference's reader tests must pass on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "go-sdr_b200", "host")
LIB = os.path.join(ROOT, "go-sdr_b200", "lib")
BIN = os.path.join(ROOT, "go-sdr_b200", "build", "test_host")


def build_test_binary():
    if not os.path.exists(os.path.join(LIB, "libhzsdrcuda.so")):
        subprocess.check_call(["bash", os.path.join(ROOT, "go-sdr_b200", "build.sh")])
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    srcs = [os.path.join(HOST, "test_host.cpp"), os.path.join(HOST, "hzsdr.hpp")]
    if not os.path.exists(BIN) or any(os.path.getmtime(s) > os.path.getmtime(BIN) for s in srcs):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-o", BIN, srcs[0], "-L" + LIB, "-lhzsdrcuda",
                               "-Wl,-rpath," + LIB])
    return BIN


def test_host_mirror_builds_against_the_c_abi():
    build_test_binary()
    # without a GPU the binary must refuse to run rather than compute anything on the CPU
    out = subprocess.run([BIN], capture_output=True, text=True)
    if out.returncode == 77:
        assert "no CPU fallback" in out.stdout


@pytest.mark.gpu
def test_host_mirror_reader_tests_on_gpu():
    out = subprocess.run([build_test_binary()], capture_output=True, text=True, timeout=600)
    print(out.stdout[-3000:], out.stderr[-2000:])
    assert out.returncode == 0, out.stdout[-3000:]
    assert " 0 failed" in out.stdout
