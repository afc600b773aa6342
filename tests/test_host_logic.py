"""Host-side logic of the library (NCO segment tables, launch planning, fixed-point phase, the window
that decides which launches may overlap), compiled from the same headers the CUDA sources include and run
on the CPU."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host_logic", "test_host_logic.cpp")
BIN = os.path.join(ROOT, "go-sdr_b200", "build", "test_host_logic")


def test_host_logic_on_cpu():
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    # the headers also carry device code, so nvcc does the compiling (it runs without a GPU); nothing is launched
    nvcc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    subprocess.check_call([nvcc, "-x", "cu", "-std=c++17", "-O2", "--expt-relaxed-constexpr", "-gencode", "arch=compute_100a,code=sm_100a",
                           "-o", BIN, SRC])
    out = subprocess.run([BIN], capture_output=True, text=True, timeout=300)
    print(out.stdout[-3000:], out.stderr[-2000:])
    assert out.returncode == 0, out.stdout[-3000:]
    assert " 0 failed" in out.stdout
