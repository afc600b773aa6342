"""CPU tests of the drop-in boundary: the C-ABI library loads, exports exactly what
include/hzsdr_cuda.h declares, fails loudly without a GPU, and its host-side logic (steering
vectors, format sizes) matches the reference's known answers.  No GPU compute here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import hzsdr as H
import kat_runner as K

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "hzsdr_cuda.h")


@pytest.fixture(scope="session", autouse=True)
def built_library():
    if not os.path.exists(os.path.abspath(H.LIB_PATH)):
        subprocess.check_call(["bash", os.path.join(ROOT, "go-sdr_b200", "build.sh")])
    return H.load()


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hzsdr_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built_library):
    decl = declared_symbols()
    assert len(decl) >= 45
    for name in decl:
        assert hasattr(built_library, name), f"{name} declared in include/hzsdr_cuda.h but not exported"
    # and the Python harness binds the same set -- nothing undeclared, nothing forgotten
    assert sorted(H.SYMBOLS) == decl


def test_exports_are_plain_c(built_library):
    out = subprocess.check_output(["nm", "-D", "--defined-only", os.path.abspath(H.LIB_PATH)], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    assert set(declared_symbols()) <= exported
    assert not any(s.startswith("_Z") for s in exported), "C++ symbols leak out of the C ABI"


def test_library_has_sm100a_code_only():
    out = subprocess.run(["cuobjdump", "--list-elf", os.path.abspath(H.LIB_PATH)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_format_size_and_version(built_library):
    # SampleFormat.Size(), iq.go:99-110
    assert [built_library.hzsdr_format_size(f) for f in (1, 2, 3, 4, 0, 9)] == [8, 2, 4, 2, 0, 0]
    assert b"sm_100a" in built_library.hzsdr_version()


def _no_gpu():
    n = C.c_int(-1)
    rc = H.load().hzsdr_device_count(C.byref(n))
    return rc != 0 or n.value == 0


@pytest.mark.skipif(not _no_gpu(), reason="a GPU is present")
def test_no_gpu_fails_loudly():
    """No CPU fallback: creating a context without a device is an error with a message, the same
    spirit as the reference's SIMD CPU-feature gate (internal/simd/enabled_amd64.go:35-50)."""
    with pytest.raises(H.HzsdrError) as ei:
        H.Context(0)
    assert ei.value.status == H.ERR_NO_DEVICE
    assert "no CPU fallback" in str(ei.value)


def test_null_context_is_rejected(built_library):
    n = C.c_size_t()
    rc = built_library.hzsdr_convert_to_c64(None, 2, None, 0, None, 0, C.byref(n))
    assert rc == H.ERR_INVALID and b"null context" in built_library.hzsdr_last_error()


def test_beamform_angles_kats_through_the_library():
    """stream/beamform_test.go:34-247 replayed against hzsdr_beamform_angles_2d (host math)."""
    K.run_beamform_angles(H)
    assert H.beamform_angles(900e6, 0, []) is None
    assert H.beamform_angles_2d(900e6, 0, (0, 10), []) is None


def test_beamform_angles_equal_oracle():
    import go_sdr_oracle as O
    ants = [(0.1 * i, 0.05 * (i % 3)) for i in range(64)]
    a = H.beamform_angles_2d(433.92e6, 27.5, (0.3, 0.0), ants)
    b = O.beamform_angles_2d(433.92e6, 27.5, (0.3, 0.0), ants)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
