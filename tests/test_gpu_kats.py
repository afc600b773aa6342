"""GPU: the reference's own known-answer tests (tests/golden/reference_kats.json) replayed
through the C ABI of libhzsdrcuda.so."""
import numpy as np
import pytest

import hzsdr as H
import kat_runner as K
from gpu_impl import GpuImpl

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    return GpuImpl()


def test_kats_convert(gpu):
    K.run_convert(gpu)


def test_kat_convert_over_under(gpu):
    """iq_u8_test.go:65-85: converting the sub-slice [32:69] must not touch its neighbours."""
    k = K.KATS["convert_u8_over_under"]
    ctx = gpu.ctx
    src = ctx.to_device(np.full(2 * k["n"], k["value"], dtype=np.uint8))
    dst = ctx.to_device(np.zeros(k["n"], dtype=np.complex64))
    lo, hi = k["lo"], k["hi"]
    got = ctx.convert_to_c64(H.FORMAT_U8, src.ptr + 2 * lo, hi - lo, dst.ptr + 8 * lo, hi - lo)
    assert got == hi - lo
    out = dst.download(np.complex64, k["n"])
    assert np.all(out[:lo] == 0) and np.all(out[hi:] == 0)
    K.in_epsilon(k["out"], out[lo:hi].real, k["tol"])
    K.in_epsilon(k["out"], out[lo:hi].imag, k["tol"])


def test_kats_scale_rotate_add(gpu):
    K.run_scale(gpu)
    K.run_rotate(gpu)
    K.run_add(gpu)


def test_kats_shift_roundtrip(gpu):
    K.run_shift_roundtrip(gpu)


def test_kats_decimate_downsample(gpu):
    K.run_decimate(gpu)
    K.run_downsample(gpu)
    k = K.KATS["decimate_short_dst"]
    with pytest.raises(H.HzsdrError) as ei:  # stream/decimate_test.go:87-96
        gpu.decimate_buffer(np.zeros(k["n"], dtype=np.complex64), k["factor"], to_len=k["dst_len"])
    assert ei.value.status == H.ERR_DST_TOO_SMALL


def test_kats_convert_matrix(gpu):
    K.run_convert_matrix(gpu)


def test_kats_fft_contract(gpu):
    K.run_fft_contract(gpu)
    with pytest.raises(H.HzsdrError) as ei:  # testutils/fft.go:127-138
        H.FftPlan(gpu.ctx, 1024, 128, H.FFT_FORWARD)
    assert ei.value.status == H.ERR_DST_TOO_SMALL
    with pytest.raises(H.HzsdrError) as ei:
        H.FftPlan(gpu.ctx, 128, 1024, H.FFT_BACKWARD)
    assert ei.value.status == H.ERR_DST_TOO_SMALL
