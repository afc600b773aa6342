"""CPU tests: pin the oracle (numpy restatement + its C twin) against the reference's own
known-answer vectors and against the reference's literal serial loops."""
import math

import numpy as np
import pytest

import cpu_ref as CR
import go_sdr_oracle as O
import kat_runner as K


# ---- the reference's KATs against the numpy oracle -------------------------------------------
def test_kats_convert():
    K.run_convert(O)


def test_kats_scale_rotate_add():
    K.run_scale(O)
    K.run_rotate(O)
    K.run_add(O)


def test_kats_shift_roundtrip():
    K.run_shift_roundtrip(O)


def test_kats_decimate_downsample():
    K.run_decimate(O)
    K.run_downsample(O)
    k = K.KATS["decimate_short_dst"]
    with pytest.raises(O.ErrDstTooSmall):
        O.decimate_buffer(np.zeros(k["n"], dtype=np.complex64), k["factor"], to_len=k["dst_len"])


def test_kats_convert_matrix():
    K.run_convert_matrix(O)


def test_lut_multiply_equals_convert_multiply_convert():
    """stream/multiply_test.go:71-112,189-230 (TestRotateU8 / TestRotateI8): the LUT readers must equal
    Convert -> Multiply -> Convert exactly, on the test's own counter pattern."""
    counter = np.arange(1024 * 60, dtype=np.uint32) & 0xffff
    i8 = np.stack([(counter & 0xff), ((counter >> 8) & 0xff).astype(np.int64) - 127], axis=1).astype(np.int8).reshape(-1)
    want = O.convert_from_c64(O.rotate(O.convert_i8_to_c64(i8), -1j), O.FORMAT_I8)
    assert np.array_equal(O.multiply_lut_i8(i8, -1j), want)
    # the u8 test pattern never reaches 255 in Q ((counter & 0xFF00) >> 8 <= 239 for 61440 samples), so
    # the x0*255+x1 collisions do not show in the reference's own test
    u8 = np.stack([(counter & 0xff), (counter >> 8) & 0xff], axis=1).astype(np.uint8).reshape(-1)
    want = O.convert_from_c64(O.rotate(O.convert_u8_to_c64(u8), -1j), O.FORMAT_U8)
    got = O.multiply_lut_u8(u8, -1j)
    ok = u8.reshape(-1, 2)[:, 1] != 255  # (x0, 255) collides with (x0 + 1, 0) in the reference's index
    assert np.array_equal(got[ok], want[ok])


def test_kats_beamform_angles():
    K.run_beamform_angles(O)
    assert O.beamform_angles(900e6, 0, []) is None  # stream/beamform_test.go:64-79
    assert O.beamform_angles_2d(900e6, 0, (0, 10), []) is None


def test_kats_fft_contract():
    K.run_fft_contract(O)


# ---- the same KATs against the C twin (what bench.py times as the CPU baseline) --------------
class _CImpl:
    convert_to_c64 = staticmethod(CR.convert_to_c64)
    scale = staticmethod(CR.scale)
    rotate = staticmethod(CR.rotate)
    shift_buffer = staticmethod(CR.shift_buffer)

    @staticmethod
    def add(*bufs):
        out = np.zeros_like(bufs[0])
        for b in bufs:
            out = CR.add2(out, b)
        return out


def test_kats_c_twin():
    K.run_convert(_CImpl)
    K.run_scale(_CImpl)
    K.run_rotate(_CImpl)
    K.run_add(_CImpl)
    K.run_shift_roundtrip(_CImpl)


# ---- numpy oracle == C twin, bit for bit where the arithmetic is exact ------------------------
def test_convert_exhaustive_numpy_vs_c():
    u8 = np.stack([np.arange(256), np.arange(256)[::-1]], axis=1).astype(np.uint8).reshape(-1)
    assert np.array_equal(O.convert_u8_to_c64(u8).view(np.uint32), CR.convert_to_c64(u8, 2).view(np.uint32))
    i8 = u8.view(np.int8)
    assert np.array_equal(O.convert_i8_to_c64(i8).view(np.uint32), CR.convert_to_c64(i8, 4).view(np.uint32))
    v = np.arange(-32768, 32768).astype(np.int16)
    i16 = np.stack([v, v[::-1]], axis=1).reshape(-1)
    assert np.array_equal(O.convert_i16_to_c64(i16).view(np.uint32), CR.convert_to_c64(i16, 3).view(np.uint32))
    # odd length exercises the SSE head remainder (iq_u8_amd64.go:34-37)
    assert np.array_equal(O.convert_u8_to_c64(u8[:2 * 37]).view(np.uint32), CR.convert_to_c64(u8[:2 * 37], 2).view(np.uint32))


def test_convert_spot_values():
    # 127 -> -0.003921569, 128 -> +0.003921569, i16 -32768 -> -1.0000305
    out = O.convert_u8_to_c64(np.array([127, 128], dtype=np.uint8))
    assert out.real[0] == np.float32(-0.5) / np.float32(127.5)
    assert out.imag[0] == np.float32(0.5) / np.float32(127.5)
    assert O.convert_i16_to_c64(np.array([-32768, 0], dtype=np.int16)).real[0] == np.float32(-32768.0) / np.float32(32767.0)
    assert O.convert_i8_to_c64(np.array([127, -128], dtype=np.int8))[0] == np.complex64(0.9921875 - 1j)


def test_shift_lsb_to_msb():
    x = np.array([2047, -2048, 1, -1], dtype=np.int16)
    assert np.array_equal(O.shift_lsb_to_msb_bits(x, 12), np.array([32752, -32768, 16, -16], dtype=np.int16))


def test_lookup_identity_and_index():
    ident = O.lookup_identity_u8()
    raw = np.array([1, 2, 255, 0, 0, 255], dtype=np.uint8)
    assert np.array_equal(O.lookup_index(raw), np.array([1 + 2 * 256, 255, 255 * 256]))
    assert np.array_equal(O.lookup(ident, raw).reshape(-1), raw)
    # a c64 table built the way stream/multiply.go:212-238 builds its tables
    tab = O.rotate(O.convert_u8_to_c64(ident.reshape(-1)), 0 - 1j)
    assert np.array_equal(O.lookup(tab, raw), O.rotate(O.convert_u8_to_c64(raw), 0 - 1j))


# ---- the shift accumulator ------------------------------------------------------------------
@pytest.mark.parametrize("fs,n,ts0", [(2_400_000, 5000, 0.0), (20_000_000, 4000, 0.0), (1_800_000, 3000, 6.2815),
                                      (61_440_000, 2000, 3.99999), (1000, 7000, 0.0)])
def test_shift_ts_vectorised_equals_serial(fs, n, ts0):
    a, ea = O.shift_ts_serial(fs, n, ts0)
    b, eb = O.shift_ts(fs, n, ts0)
    assert np.array_equal(a, b) and ea == eb
    c, ec = CR.shift_ts(fs, n, ts0)
    assert np.array_equal(a, c) and ea == ec


@pytest.mark.parametrize("fs,n,ts0", [(2_400_000, 1 << 20, 0.0), (20_000_000, 1 << 22, 0.0), (61_440_000, 1 << 22, 0.0),
                                      (20_000_000, 1 << 21, 6.2), (1_800_000, 61440, 0.0), (1000, 20000, 0.0),
                                      (3, 100, 0.0), (48_000, 1 << 18, 6.28)])
def test_shift_segments_equal_serial_loop(fs, n, ts0):
    """The closed-form segment table (the host logic libhzsdrcuda mirrors) reproduces the
    reference's serial fp64 accumulator bit-for-bit, including the 2*pi wrap."""
    want, want_end = CR.shift_ts(fs, n, ts0)
    segs, end = O.shift_segments(fs, n, ts0)
    got = O.expand_segments(segs, n)
    assert np.array_equal(want, got)
    assert end == want_end
    assert len(segs) < 400


def test_shift_buffer_numpy_vs_c():
    raw = O.synth_raw(O.FORMAT_U8, 1 << 16, 2_400_000, 300e3, seed=1)
    x = O.convert_u8_to_c64(raw)
    a, ta = O.shift_buffer(x, -300e3, 2_400_000, 0.0)
    b, tb = CR.shift_buffer(x, -300e3, 2_400_000, 0.0)
    assert ta == tb
    # glibc sincos vs numpy sin/cos may differ in the last fp64 ulp -> at most a last-bit fp32 difference
    assert O.rel_l2(a, b) < 1e-7
    # phase continuity: two half buffers == one buffer
    h1, t1 = O.shift_buffer(x[: 1 << 15], -300e3, 2_400_000, 0.0)
    h2, t2 = O.shift_buffer(x[1 << 15:], -300e3, 2_400_000, t1)
    assert t2 == ta and np.array_equal(np.concatenate([h1, h2]), a)


def test_shift_moves_carrier_to_dc():
    # what the reference's own test admits it never checks (stream/shifter_test.go:44-46)
    fs, f0, n = 2_400_000, 300e3, 1 << 14
    x = O.cw(n, f0, fs)
    y, _ = O.shift_buffer(x, -f0, fs)
    assert int(np.argmax(np.abs(np.fft.fft(y)))) == 0
    assert int(np.argmax(np.abs(np.fft.fft(x)))) == round(f0 / fs * n)


def test_go_complex_mul_is_widened():
    a = np.array([1.0000001 + 0.99999994j], dtype=np.complex64)
    b = np.complex64(0.99999994 + 1.0000001j)
    got = O.go_complex64_mul(a, b)
    ar, ai, br, bi = map(float, (a.real[0], a.imag[0], b.real, b.imag))
    assert got.real[0] == np.float32(ar * br - ai * bi) and got.imag[0] == np.float32(ar * bi + ai * br)
    assert np.array_equal(got, CR.rotate(a, b))


# ---- FFT / convolution (parity unpinned by the reference; pin numpy vs the C stand-in) -------
@pytest.mark.parametrize("n", [8, 64, 1024, 4096])
def test_fft_numpy_vs_c_standin(n):
    rng = np.random.default_rng(n)
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    p = CR.Plan(n)
    assert O.rel_l2(p.fft(x, -1), O.fft_forward(x)) < 2e-6
    assert O.rel_l2(p.fft(x, +1), O.fft_backward(x)) < 2e-6
    assert O.rel_l2(O.fft_backward(O.fft_forward(x)), x * n) < 1e-6  # unnormalised both ways


def test_convolution_reader_is_block_circular():
    n, taps = 256, 31
    h = O.lowpass_taps(taps, 0.1)
    H = O.filter_freq(h, n)
    rng = np.random.default_rng(7)
    x = (rng.standard_normal(3 * n + 17) + 1j * rng.standard_normal(3 * n + 17)).astype(np.complex64)
    z = O.convolution_reader(x, H)
    assert z.shape[0] == 3 * n  # trailing partial block dropped
    for b in range(3):
        blk = x[b * n:(b + 1) * n].astype(np.complex128)
        want = np.array([sum(h[k] * blk[(m - k) % n] for k in range(taps)) for m in range(n)])
        assert O.rel_l2(z[b * n:(b + 1) * n], want) < 1e-6


def test_chain_numpy_vs_c():
    fs, n, nfft, d = 20_000_000, 1 << 17, 1024, 10
    raw = O.synth_raw(O.FORMAT_I8, n, fs, 2.5e6, seed=2)
    H = O.filter_freq(O.lowpass_taps(255, 1 / 20), nfft)
    a, ta = O.chain(raw, O.FORMAT_I8, fs, -2.5e6, H, d)
    b, tb = CR.chain(raw, O.FORMAT_I8, fs, -2.5e6, H, d)
    assert a.shape == b.shape == ((n // 32768) * (32768 // d),)
    assert ta == tb
    assert O.rel_l2(b, a) < 5e-6  # radix-2 fp32 stand-in vs complex128 pocketfft
    # the carrier lands at DC: output power dominated by the mean
    assert abs(a.mean()) > 0.45


def test_beamform_numpy_vs_c():
    n, c = 4096, 8
    chans = [O.synth_raw(O.FORMAT_U8, n, 2_400_000, 100e3, seed=10 + i, phase=0.3 * i) for i in range(c)]
    w = O.beamform_angles(433e6, 20.0, [0.1 * i for i in range(c)])
    assert w[0] == 1  # exercised: Multiply skips m == 1 (stream/multiply.go:59-62)
    assert np.array_equal(O.beamform(chans, O.FORMAT_U8, w).view(np.uint32), CR.beamform_u8(chans, w).view(np.uint32))


def test_synth_shapes():
    for fmt, dt in ((O.FORMAT_U8, np.uint8), (O.FORMAT_I8, np.int8), (O.FORMAT_I16, np.int16)):
        r = O.synth_raw(fmt, 1000, 2_400_000, 300e3, seed=3)
        assert r.dtype == dt and r.shape == (2000,)


# ---- coherent-receiver helpers (rtl/kerberos/internal; no reference tests exist for them) -----
def test_fftshift_and_scale_restates_the_loop():
    # reader.go:57-64 written out literally on a small vector
    data = np.array([1 + 2j, 3 + 4j, 5 + 6j, 7 + 8j, 9 + 10j, 11 + 12j], dtype=np.complex64)
    scale = np.float32(3.0)
    want = data.copy()
    half = len(want) // 2
    for i in range(half):
        a, b = want[i], want[half + i]
        want[i] = np.complex64(complex(np.float32(b.real) / scale, np.float32(b.imag) / scale))
        want[half + i] = np.complex64(complex(np.float32(a.real) / scale, np.float32(a.imag) / scale))
    assert np.array_equal(O.fftshift_and_scale(data, 3.0).view(np.uint32), want.view(np.uint32))
    # odd length: the middle element is left alone, like the reference's loop
    odd = np.arange(5).astype(np.complex64)
    assert O.fftshift_and_scale(odd, 1.0).tolist() == [2, 3, 0, 1, 4]


def test_graft_properties():
    rng = np.random.default_rng(5)
    size = 256
    x = (rng.standard_normal((1, size)) + 1j * rng.standard_normal((1, size))).astype(np.complex64)
    sign = np.where(np.arange(size) % 2 == 0, 1.0, -1.0)
    assert O.rel_l2(O.graft(x), x[0] * sign) < 1e-6  # one reader: identity up to the fftshift's (-1)^n
    # two readers carrying tones at their own centre: the graft puts them at -fs/2 and +fs/2 of the
    # doubled-rate stream, i.e. reader 0's DC lands in bin N/4 and reader 1's in bin 3N/4
    two = np.ones((2, size), dtype=np.complex64)
    spec = np.abs(O.fft_forward(O.graft(two)[None, :])[0])
    assert sorted(np.argsort(spec)[-2:].tolist()) == [size // 2, 3 * size // 2]


def test_correlate_peak_and_phase_offsets():
    rng = np.random.default_rng(6)
    n = 4096
    a = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    for delay, want in ((0, 0), (5, 5), (n // 2, n // 2), (n // 2 + 1, -(n // 2) + 1), (n - 3, -3)):
        assert O.correlate_peak(O.cross_correlate(a, np.roll(a, -delay))) == want
    assert O.correlate_peak(np.zeros(16, dtype=np.complex64)) == -1
    bufs = np.stack([a, a * np.exp(-0.5j), a * np.exp(1.25j)]).astype(np.complex64)
    ph = O.phase_offsets(bufs)
    assert abs(np.angle(ph[0]) - 1.0 / n) < 1e-7          # the reference's quirk (align.go:265)
    assert np.allclose(np.angle(ph[1:]), [0.5, -1.25], atol=1e-4)
