"""GPU parity: libhzsdrcuda.so (through the C ABI) against the oracle on identical seeded inputs.

Bars (BASELINE.json north_star): integer -> complex64 conversion bit-exact; everything else
relative L2 <= 1e-5 per buffer (we assert tighter where the arithmetic allows, and bit-equality
where the library reproduces the Go arithmetic exactly: rotate, scale, add, decimate, lookup and
the carried NCO time `ts`)."""
import numpy as np
import pytest

import cpu_ref as CR
import go_sdr_oracle as O
import hzsdr as H
from gpu_impl import GpuImpl

pytestmark = pytest.mark.gpu

TOL = 1e-5  # north_star: relative L2 per buffer


@pytest.fixture(scope="module")
def gpu():
    return GpuImpl()


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


# ---------------------------------------------------------------------------------------------
# K1 convert: bit-exact, exhaustive
# ---------------------------------------------------------------------------------------------
def test_convert_u8_i8_all_codes(gpu):
    a = np.arange(256, dtype=np.uint8)
    u8 = np.stack(np.meshgrid(a, a, indexing="ij"), axis=-1).reshape(-1)  # all 65536 IQ pairs
    assert np.array_equal(bits(gpu.convert_to_c64(u8, H.FORMAT_U8)), bits(O.convert_u8_to_c64(u8)))
    i8 = u8.view(np.int8)
    assert np.array_equal(bits(gpu.convert_to_c64(i8, H.FORMAT_I8)), bits(O.convert_i8_to_c64(i8)))


def test_convert_i16_all_codes(gpu):
    v = np.arange(-32768, 32768).astype(np.int16)
    i16 = np.stack([v, v[::-1]], axis=1).reshape(-1)
    assert np.array_equal(bits(gpu.convert_to_c64(i16, H.FORMAT_I16)), bits(O.convert_i16_to_c64(i16)))


@pytest.mark.parametrize("fmt", [H.FORMAT_U8, H.FORMAT_I8, H.FORMAT_I16])
@pytest.mark.parametrize("n", [0, 1, 2, 3, 37, 255, 1000, 4097, (1 << 20) + 1])
def test_convert_ragged_lengths(gpu, fmt, n):
    rng = np.random.default_rng(n + fmt)
    dt = H.NP_DTYPE[fmt]
    raw = rng.integers(np.iinfo(dt).min, np.iinfo(dt).max, size=2 * n, endpoint=True).astype(dt)
    got = gpu.convert_to_c64(raw, fmt)
    assert got.shape == (n,)
    assert np.array_equal(bits(got), bits(O.convert_to_c64(raw, fmt)))


@pytest.mark.parametrize("fmt", [H.FORMAT_U8, H.FORMAT_I16])
@pytest.mark.parametrize("off", [1, 2, 3, 33])
def test_convert_offset_subslices(gpu, fmt, off):
    """Sub-slices at odd sample offsets (misaligned for the vector path) stay exact and in bounds."""
    ctx = gpu.ctx
    n, m = 4096, 1001
    dt = H.NP_DTYPE[fmt]
    raw = np.random.default_rng(off).integers(np.iinfo(dt).min, np.iinfo(dt).max, size=2 * n, endpoint=True).astype(dt)
    sb = raw.itemsize * 2
    src = ctx.to_device(raw)
    dst = ctx.to_device(np.zeros(n, dtype=np.complex64))
    ctx.convert_to_c64(fmt, src.ptr + sb * off, m, dst.ptr + 8 * off, m)
    out = dst.download(np.complex64, n)
    assert np.all(out[:off] == 0) and np.all(out[off + m:] == 0)
    assert np.array_equal(bits(out[off:off + m]), bits(O.convert_to_c64(raw[2 * off:2 * (off + m)], fmt)))
    # source and destination offsets of different parity -> scalar path
    dst2 = ctx.to_device(np.zeros(n, dtype=np.complex64))
    ctx.convert_to_c64(fmt, src.ptr + sb * off, m, dst2.ptr + 8 * (off + 1), m)
    out2 = dst2.download(np.complex64, n)
    assert np.array_equal(bits(out2[off + 1:off + 1 + m]), bits(out[off:off + m]))


def test_convert_errors_and_copy(gpu):
    ctx = gpu.ctx
    with pytest.raises(H.HzsdrError) as ei:  # conv.go:60-62
        gpu.convert_to_c64(np.zeros(200, dtype=np.uint8), H.FORMAT_U8, dst_len=50)
    assert ei.value.status == H.ERR_DST_TOO_SMALL
    with pytest.raises(H.HzsdrError) as ei:
        gpu.convert_to_c64(np.zeros(8, dtype=np.uint8), 9)
    assert ei.value.status == H.ERR_FORMAT_UNKNOWN
    x = (np.arange(100) + 1j * np.arange(100)).astype(np.complex64)  # same format = copy, conv.go:56-58
    assert np.array_equal(gpu.convert_to_c64(x, H.FORMAT_C64), x)


# ---------------------------------------------------------------------------------------------
# the rest of the conversion matrix, integer Add, LUT Multiply: bit-exact
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("fmt", [H.FORMAT_U8, H.FORMAT_I8, H.FORMAT_I16])
def test_convert_from_c64_bit_exact(gpu, fmt):
    rng = np.random.default_rng(fmt)
    n = 200_003
    x = (rng.uniform(-1.2, 1.2, n) + 1j * rng.uniform(-1.2, 1.2, n)).astype(np.complex64)  # includes out-of-range wraps
    x[:8] = [1, -1, 0, 1j, -1j, 0.999999 + 0.5j, -0.0, 1e-8]
    # every exact code boundary of the forward conversion round-trips
    codes = O.convert_to_c64(np.arange(-128, 128).astype(np.int8).repeat(2), O.FORMAT_I8)
    x[8:8 + codes.size] = codes
    assert np.array_equal(gpu.convert_from_c64(x, fmt), O.convert_from_c64(x, fmt))


@pytest.mark.parametrize("src,dst", [(H.FORMAT_U8, H.FORMAT_I8), (H.FORMAT_U8, H.FORMAT_I16), (H.FORMAT_I8, H.FORMAT_U8),
                                      (H.FORMAT_I8, H.FORMAT_I16), (H.FORMAT_I16, H.FORMAT_U8), (H.FORMAT_I16, H.FORMAT_I8)])
def test_convert_int_all_codes(gpu, src, dst):
    dt = H.NP_DTYPE[src]
    info = np.iinfo(dt)
    v = np.arange(info.min, info.max + 1).astype(dt)
    raw = np.stack([v, v[::-1]], axis=1).reshape(-1)
    assert np.array_equal(gpu.convert_int(raw, src, dst), O.convert_int(raw, src, dst))


def test_convert_matrix_errors(gpu):
    with pytest.raises(H.HzsdrError) as ei:
        gpu.convert(np.zeros(64, np.uint8), H.FORMAT_U8, H.FORMAT_I16, dst_len=4)
    assert ei.value.status == H.ERR_DST_TOO_SMALL
    with pytest.raises(H.HzsdrError) as ei:
        gpu.convert(np.zeros(64, np.uint8), H.FORMAT_U8, 7)
    assert ei.value.status == H.ERR_FORMAT_UNKNOWN
    same = np.arange(64, dtype=np.int16)
    assert np.array_equal(gpu.convert(same, H.FORMAT_I16, H.FORMAT_I16).reshape(-1), same)


def test_add_int_wraps(gpu):
    rng = np.random.default_rng(9)
    for dt in (np.int8, np.int16):
        bufs = [rng.integers(np.iinfo(dt).min, np.iinfo(dt).max, size=2 * 5001, endpoint=True).astype(dt) for _ in range(5)]
        assert np.array_equal(gpu.add_int(*bufs), O.add_int(*bufs))


def test_lut_multiply_u8_i8(gpu):
    """stream/multiply_test.go:71-112,189-230: LUT Multiply on raw streams == the reference's table,
    including the u8 index collisions."""
    rng = np.random.default_rng(3)
    m = np.complex64(0.6 - 0.8j)
    i8 = rng.integers(-128, 127, size=2 * 70_001, endpoint=True).astype(np.int8)
    assert np.array_equal(gpu.multiply_lut(i8, m, H.FORMAT_I8), O.multiply_lut_i8(i8, m))
    u8 = rng.integers(0, 255, size=2 * 70_001, endpoint=True).astype(np.uint8)
    assert np.array_equal(gpu.multiply_lut(u8, m, H.FORMAT_U8), O.multiply_lut_u8(u8, m))


def test_i16_shift_lsb_to_msb(gpu):
    ctx = gpu.ctx
    raw = np.random.default_rng(5).integers(-2048, 2047, size=2 * 5001, endpoint=True).astype(np.int16)
    d = ctx.to_device(raw)
    H._check(H.load().hzsdr_i16_shift_lsb_to_msb(ctx.h, d.ptr, 5001, 12))
    assert np.array_equal(d.download(np.int16, raw.size), O.shift_lsb_to_msb_bits(raw, 12))


def test_lookup_tables(gpu):
    rng = np.random.default_rng(11)
    raw = rng.integers(0, 255, size=2 * 70001, endpoint=True).astype(np.uint8)
    ident = O.lookup_identity_u8()
    # c64 table built the way stream/multiply.go:212-238 builds it: Convert -> Multiply
    tab = O.rotate(O.convert_u8_to_c64(ident.reshape(-1)), 0 - 1j)
    assert np.array_equal(bits(gpu.lookup(tab, raw)), bits(O.lookup(tab, raw)))
    assert np.array_equal(gpu.lookup(ident, raw, table_fmt=H.FORMAT_U8), O.lookup(ident, raw))
    tab16 = rng.integers(-32768, 32767, size=(65536, 2), endpoint=True).astype(np.int16)
    assert np.array_equal(gpu.lookup(tab16, raw.view(np.int8), src_fmt=H.FORMAT_I8, table_fmt=H.FORMAT_I16), O.lookup(tab16, raw))


# ---------------------------------------------------------------------------------------------
# K2 shift
# ---------------------------------------------------------------------------------------------
SHIFT_CASES = [
    # fs, n, shift, ts0
    (2_400_000, 1 << 20, -300e3, 0.0),           # C1: stream start
    (20_000_000, 1 << 21, -2.5e6, 0.0),          # C2 rate
    (20_000_000, 1 << 20, 5e6, 6.25),            # crosses the 2*pi-second wrap
    (61_440_000, 1 << 20, -7.68e6, 3.9999),      # crosses a binade edge (4.0)
    (1_800_000, 61440 + 3, 1000.0, 0.0),         # odd length
    (48_000, 1 << 17, 1234.5, 6.2),              # low rate, wraps
]


@pytest.mark.parametrize("fs,n,shift,ts0", SHIFT_CASES)
def test_shift_parity(gpu, fs, n, shift, ts0):
    x = O.convert_u8_to_c64(O.synth_raw(O.FORMAT_U8, n, fs, -shift if abs(shift) < fs / 2 else 0.0, seed=n % 97))
    want, ts_want = CR.shift_buffer(x, shift, fs, ts0)  # the literal serial loop, compiled
    got, ts_got = gpu.shift_buffer(x, shift, fs, ts0)
    assert ts_got == ts_want, "carried ts must be bit-equal to the reference accumulator"
    err = O.rel_l2(got, want)
    assert err <= TOL, err
    assert err <= 5e-7, err  # what the arithmetic actually achieves (rotation recurrence, depth <= 4)


def test_shift_continues_across_buffers(gpu):
    fs, shift = 20_000_000, -2.5e6
    x = O.convert_i8_to_c64(O.synth_raw(O.FORMAT_I8, 3 << 18, fs, 2.5e6, seed=3))
    want, ts_want = CR.shift_buffer(x, shift, fs, 0.0)
    ts = 0.0
    parts = []
    for part in np.split(x, 3):
        y, ts = gpu.shift_buffer(part, shift, fs, ts)
        parts.append(y)
    assert ts == ts_want
    assert O.rel_l2(np.concatenate(parts), want) <= 5e-7


def test_shift_unaligned_buffer(gpu):
    ctx = gpu.ctx
    fs, n = 2_400_000, 10001
    x = O.cw(n + 3, 1e3, fs)
    d = ctx.to_device(x)
    st = H.NcoState(fs, 0.0)
    ctx.shift(d.ptr + 8, n, 50e3, st)  # 8-byte (not 16-byte) aligned start
    out = d.download(np.complex64, n + 3)
    want, ts = CR.shift_buffer(x[1:1 + n], 50e3, fs, 0.0)
    assert st.ts == ts
    assert np.array_equal(out[0], x[0]) and np.array_equal(out[n + 1:], x[n + 1:])
    assert O.rel_l2(out[1:1 + n], want) <= 5e-7


@pytest.mark.parametrize("fmt", [H.FORMAT_U8, H.FORMAT_I8, H.FORMAT_I16])
def test_convert_shift_fused(gpu, fmt):
    fs, n, shift = 2_400_000, (1 << 19) + 5, -300e3
    raw = O.synth_raw(fmt, n, fs, 300e3, seed=fmt)
    want, ts_want = CR.shift_buffer(O.convert_to_c64(raw, fmt), shift, fs, 0.0)
    got, ts = gpu.convert_shift(raw, fmt, shift, fs, 0.0)
    assert ts == ts_want
    assert O.rel_l2(got, want) <= 5e-7
    # the carrier is now at DC
    assert abs(got[:4096].mean()) > 0.4


# ---------------------------------------------------------------------------------------------
# K3/K4/K5
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 31, 1024, 100003])
def test_rotate_scale_bit_exact(gpu, n):
    rng = np.random.default_rng(n)
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    m = np.complex64(0.70710678 - 0.31234j)
    assert np.array_equal(bits(gpu.rotate(x, m)), bits(O.rotate(x, m)))  # fp64-widened like gc
    assert np.array_equal(bits(gpu.scale(x, 0.3333)), bits(O.scale(x, 0.3333)))


@pytest.mark.parametrize("k,n", [(1, 1000), (2, 31), (3, 1000), (16, 8192), (70, 4096)])
def test_add_ordered_bit_exact(gpu, k, n):
    rng = np.random.default_rng(k * n)
    bufs = [(rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64) for _ in range(k)]
    assert np.array_equal(bits(gpu.add(*bufs)), bits(O.add(*bufs)))


# ---------------------------------------------------------------------------------------------
# K7
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("factor", [1, 3, 10, 16, 4097])
def test_decimate_reader_blocks(gpu, factor):
    n = 5 * 32768 + 1234  # trailing partial block is dropped
    x = (np.arange(n) + 1j * (np.arange(n) % 977)).astype(np.complex64)
    assert np.array_equal(gpu.decimate_reader(x, factor), O.decimate_reader(x, factor))
    assert np.array_equal(gpu.decimate_buffer(x, factor), O.decimate_buffer(x, factor))


def test_decimate_integer_formats(gpu):
    rng = np.random.default_rng(2)
    u8 = rng.integers(0, 255, size=(3 * 32768, 2), endpoint=True).astype(np.uint8)
    assert np.array_equal(gpu.decimate_reader(u8, 10, fmt=H.FORMAT_U8), O.decimate_reader(u8, 10))
    i16 = rng.integers(-32768, 32767, size=(2 * 32768, 2), endpoint=True).astype(np.int16)
    assert np.array_equal(gpu.decimate_reader(i16, 7, fmt=H.FORMAT_I16), O.decimate_reader(i16, 7))
    with pytest.raises(H.HzsdrError) as ei:  # stream/decimate.go:85-97 has no I8 case
        gpu.decimate_reader(u8.view(np.int8), 10, fmt=H.FORMAT_I8)
    assert ei.value.status == H.ERR_FORMAT_UNKNOWN


@pytest.mark.parametrize("factor", [2, 4, 10, 100])
def test_downsample_parity(gpu, factor):
    n = 3 * 32768
    rng = np.random.default_rng(factor)
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    assert np.array_equal(bits(gpu.downsample_reader(x, factor)), bits(O.downsample_reader(x, factor)))
    raw = O.synth_raw(O.FORMAT_U8, n, 2_400_000, 1e5, seed=factor)
    assert np.array_equal(bits(gpu.downsample_reader(raw, factor, fmt=H.FORMAT_U8)),
                          bits(O.downsample_reader(raw, factor, fmt=O.FORMAT_U8)))


# ---------------------------------------------------------------------------------------------
# K6 FFT / convolution
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384])
def test_fft_all_lengths(gpu, n):
    rng = np.random.default_rng(n)
    batch = 5 if n >= 4096 else 37
    x = (rng.standard_normal((batch, n)) + 1j * rng.standard_normal((batch, n))).astype(np.complex64)
    f = gpu.fft_forward(x)
    assert O.rel_l2(f, O.fft_forward(x)) <= 1e-6
    b = gpu.fft_backward(x)
    assert O.rel_l2(b, O.fft_backward(x)) <= 1e-6
    assert O.rel_l2(gpu.fft_backward(f), x * n) <= 1e-6  # unnormalised both ways


@pytest.mark.parametrize("lg", [15, 16, 17, 18, 19, 20])
def test_fft_big_lengths(gpu, lg):
    """2^15 .. 2^20 points (two kernels through context scratch, bigfft.cu): the reference's Kerberos
    helpers plan 65536 points and n_readers x 65536 (align.go:92-100, graft.go:73-80)."""
    n = 1 << lg
    rng = np.random.default_rng(lg)
    batch = 3 if lg <= 17 else 1
    x = (rng.standard_normal((batch, n)) + 1j * rng.standard_normal((batch, n))).astype(np.complex64)
    f = gpu.fft_forward(x)
    assert O.rel_l2(f, O.fft_forward(x)) <= 1e-6
    b = gpu.fft_backward(x)
    assert O.rel_l2(b, O.fft_backward(x)) <= 1e-6
    assert O.rel_l2(gpu.fft_backward(f), x * n) <= 2e-6  # unnormalised both ways
    # a complex exponential lands in exactly one bin (the planner contract, testutils/fft.go:60-92)
    k = (5 * n) // 16 + 3
    tone = np.exp(2j * np.pi * k * np.arange(n) / n).astype(np.complex64)[None, :]
    assert int(np.argmax(np.abs(gpu.fft_forward(tone)[0]))) == k


def test_fft_in_place_big(gpu):
    n = 1 << 16
    rng = np.random.default_rng(16)
    x = (rng.standard_normal((2, n)) + 1j * rng.standard_normal((2, n))).astype(np.complex64)
    ctx = gpu.ctx
    d = ctx.to_device(x)
    plan = H.FftPlan(ctx, n, n, H.FFT_FORWARD)
    plan.transform(d.ptr, d.ptr, 2)
    assert O.rel_l2(d.download(np.complex64, x.size).reshape(x.shape), O.fft_forward(x)) <= 1e-6


def test_fft_unsupported_lengths(gpu):
    for n in (3, 1000, 3 << 14, 1 << 21):
        with pytest.raises(H.HzsdrError) as ei:
            H.FftPlan(gpu.ctx, n, n, H.FFT_FORWARD)
        assert ei.value.status == H.ERR_UNSUPPORTED


@pytest.mark.parametrize("n,taps", [(8, 3), (64, 15), (256, 63), (1024, 255), (2048, 255), (4096, 1023), (8192, 1023), (16384, 4095)])
def test_convolution_reader_parity(gpu, n, taps):
    nblk = 9 if n >= 4096 else 67
    rng = np.random.default_rng(n + taps)
    x = (rng.standard_normal(nblk * n + 5) + 1j * rng.standard_normal(nblk * n + 5)).astype(np.complex64)
    Hf = O.filter_freq(O.lowpass_taps(taps, 1 / 20), n)
    got = gpu.convolution_reader(x, Hf)
    want = O.convolution_reader(x, Hf)
    assert got.shape == want.shape == (nblk * n,)
    assert O.rel_l2(got, want) <= TOL
    assert O.rel_l2(got, want) <= 2e-6


def test_dependent_launches_are_not_overlapped(gpu):
    """Kernels that let their successor start early (programmatic dependent launch: the N = 1024
    ConvolveFreq / chain kernels, convert, shift, beamform) must still be ordered when the successor
    depends on them.  ConvolutionReaders in series -- ping-pong and in place, no host
    synchronisation in between -- equal the same stages run one synchronised step at a time; a
    large launch followed by a one-block launch into the same destination leaves the small one's
    samples on top; convert -> shift in place -> shift in place on one buffer likewise."""
    ctx = gpu.ctx
    n, nblk = 1024, 8192  # 64 MiB per buffer: many waves, so an early successor would see stale data
    rng = np.random.default_rng(99)
    x = (rng.standard_normal(n * nblk) + 1j * rng.standard_normal(n * nblk)).astype(np.complex64)
    dH1 = ctx.to_device(O.filter_freq(O.lowpass_taps(255, 1 / 20), n))
    dH2 = ctx.to_device(O.filter_freq(O.lowpass_taps(127, 1 / 8), n))
    a, b = ctx.alloc(x.nbytes), ctx.alloc(x.nbytes)

    def series(sync):
        a.upload(x)
        for src, dst, hf in [(a, b, dH1), (b, a, dH2), (a, a, dH1), (a, b, dH2)]:
            ctx.convolve_freq(src.ptr, dst.ptr, hf.ptr, n, nblk)
            if sync:
                ctx.sync()
        return b.download(np.complex64, n * nblk)

    got, want = series(sync=False), series(sync=True)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))

    # write-after-write: a full-size launch, then one block into the tail of the same destination
    small = ctx.to_device(x[:n] * np.complex64(3))
    a.upload(x)
    ctx.convolve_freq(a.ptr, b.ptr, dH1.ptr, n, nblk)
    ctx.convolve_freq(small.ptr, b.ptr + 8 * n * (nblk - 1), dH2.ptr, n, 1)
    tail = b.download(np.complex64, n, 8 * n * (nblk - 1))
    ctx.convolve_freq(small.ptr, b.ptr, dH2.ptr, n, 1)
    ctx.sync()
    assert np.array_equal(tail.view(np.uint32), b.download(np.complex64, n).view(np.uint32))

    # convert -> shift in place -> shift in place, unsynchronised, against the staged oracle
    fs, m = 2_400_000, 1 << 23
    raw = O.synth_raw(H.FORMAT_U8, m, fs, 3e5, seed=1)
    src, y = ctx.to_device(raw), ctx.alloc(m * 8)
    st = H.NcoState(fs, 0.0)
    st2 = H.NcoState(fs, 0.0)
    ctx.convert_to_c64(H.FORMAT_U8, src.ptr, m, y.ptr, m)
    ctx.shift(y.ptr, m, -3e5, st)
    ctx.shift(y.ptr, m, 1e5, st2)
    c = O.convert_u8_to_c64(raw)
    c, _ = O.shift_buffer(c, -3e5, fs, 0.0)
    c, _ = O.shift_buffer(c, 1e5, fs, 0.0)
    assert O.rel_l2(y.download(np.complex64, m), c) <= TOL


@pytest.mark.parametrize("n", [64, 1024, 4096, 65536])
@pytest.mark.parametrize("xc", [False, True])
def test_fft_convolve_and_cross_correlate(gpu, n, xc):
    """fft.Convolve / fft.CrossCorrelate (fft/convolution.go:97-139); the correlation of a buffer
    with a circularly delayed copy peaks at the delay (what rtl/kerberos/internal/align.go uses it for)."""
    rng = np.random.default_rng(n + xc)
    batch, delay = 5, 37 % n
    a = (rng.standard_normal((batch, n)) + 1j * rng.standard_normal((batch, n))).astype(np.complex64)
    b = np.roll(a, -delay, axis=1) if xc else (rng.standard_normal((batch, n)) + 1j * rng.standard_normal((batch, n))).astype(np.complex64)
    ctx = gpu.ctx
    da, db = ctx.to_device(a), ctx.to_device(b)
    out, scratch = ctx.alloc(a.nbytes), ctx.alloc(a.nbytes)
    H._check(H.load().hzsdr_fft_convolve(ctx.h, out.ptr, da.ptr, db.ptr, n, batch, int(xc), scratch.ptr))
    got = out.download(np.complex64, a.size).reshape(a.shape)
    assert O.rel_l2(got, O.fft_convolve(a, b, xc)) <= TOL
    if xc:
        assert np.all(np.argmax(np.abs(got), axis=1) == delay)
        # checkAlignment's search on the device (align.go:125-146)
        wrapped = delay - n if delay > n // 2 else delay
        assert np.array_equal(ctx.correlate_peak(out.ptr, n, batch), np.full(batch, wrapped, dtype=np.int32))


# ---------------------------------------------------------------------------------------------
# coherent-receiver helpers: rtl/kerberos/internal
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,scale", [(8, 8.0), (65536, 65536.0), (1024, 1000.0), (10, 3.0)])
def test_fftshift_scale_bit_exact(gpu, n, scale):
    rng = np.random.default_rng(n)
    x = (rng.standard_normal((3, n)) + 1j * rng.standard_normal((3, n))).astype(np.complex64)
    d = gpu.ctx.to_device(x)
    gpu.ctx.fftshift_scale(d.ptr, n, 3, scale)
    got = d.download(np.complex64, x.size).reshape(x.shape)
    assert np.array_equal(bits(got), bits(O.fftshift_and_scale(x, scale)))


@pytest.mark.parametrize("delay", [0, 1, 37, 32768, 32769, 65535])
def test_correlate_peak_alignment(gpu, delay):
    """CrossCorrelater + checkAlignment at the reference's 65536 points: reader 1 lags reader 0 by
    `delay` samples (noise added); offsets beyond n/2 come back negative (align.go:142-145)."""
    n = 1 << 16
    rng = np.random.default_rng(delay)
    a = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    b = (np.roll(a, -delay) + 0.1 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64)
    ctx = gpu.ctx
    da, db = ctx.to_device(a), ctx.to_device(b)
    out, scratch = ctx.alloc(a.nbytes), ctx.alloc(a.nbytes)
    ctx.cross_correlate(out.ptr, da.ptr, db.ptr, n, 1, scratch.ptr)
    cc = out.download(np.complex64, n)
    want_cc = O.cross_correlate(a, b)
    assert O.rel_l2(cc, want_cc) <= TOL
    want = O.correlate_peak(want_cc)
    assert want == (delay - n if delay > n // 2 else delay)
    assert int(ctx.correlate_peak(out.ptr, n, 1)[0]) == want


def test_correlate_peak_ties_and_zeros(gpu):
    n = 4096
    cc = np.zeros((3, n), dtype=np.complex64)
    cc[0, [100, 900, 3000]] = [2 + 0j, 2j, -2 + 0j]  # equal powers: the first wins
    cc[1, 4000] = 1e-20 + 0j                          # past n/2: negative offset
    # cc[2] all zero: the reference leaves maxPowI at -1
    d = gpu.ctx.to_device(cc)  # keep the buffer alive across the call
    got = gpu.ctx.correlate_peak(d.ptr, n, 3)
    assert got.tolist() == [100, 4000 - n, -1] == [O.correlate_peak(c) for c in cc]


@pytest.mark.parametrize("nr,size", [(4, 65536), (2, 65536), (8, 4096), (1, 1024), (16, 65536)])
def test_graft_parity(gpu, nr, size):
    """One pass of GraftReaders' loop (graft.go:96-125) against the oracle; with one reader the graft
    is the identity up to the fftshift's (-1)^n."""
    rng = np.random.default_rng(nr * size)
    x = (rng.standard_normal((nr, size)) + 1j * rng.standard_normal((nr, size))).astype(np.complex64)
    ctx = gpu.ctx
    dx, dst, freq = ctx.to_device(x), ctx.alloc(x.nbytes), ctx.alloc(x.nbytes)
    ctx.graft(dx.ptr, nr, size, dst.ptr, freq.ptr)
    got = dst.download(np.complex64, x.size)
    assert O.rel_l2(got, O.graft(x)) <= TOL
    if nr == 1:
        sign = np.where(np.arange(size) % 2 == 0, 1.0, -1.0).astype(np.float32)
        assert O.rel_l2(got, x[0] * sign) <= TOL


def test_phase_offsets_parity(gpu):
    """PhaseOffsets (align.go:244-272) at the reference's 65536 samples, 4 receivers with PLL phase
    offsets; element 0 is the reference's Rect(1, 1/n)."""
    n, nchan = 1 << 16, 4
    rng = np.random.default_rng(7)
    base = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    true = np.array([0.0, 0.7, -2.1, 1.5])
    bufs = np.stack([(base * np.exp(-1j * p) + 0.05 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64)
                     for p in true])
    d = gpu.ctx.to_device(bufs)  # keep the buffer alive across the call
    got = gpu.ctx.phase_offsets(d.ptr, nchan, n)
    want = O.phase_offsets(bufs)
    assert np.max(np.abs(got - want)) <= 1e-6
    assert abs(np.angle(got[0]) - 1.0 / n) <= 1e-7
    assert np.max(np.abs(np.angle(got[1:] * np.exp(-1j * true[1:])))) <= 0.05  # recovers the PLL offsets


def test_overlappable_kernel_after_a_kernel_outside_the_scheme(gpu):
    """An overlappable kernel (shift, convert, chain) that follows a kernel OUTSIDE the overlap scheme
    (rotate, add, decimate, the generic FFT kernels, the batched channelizer launch) and reads its output
    must not be launched with the programmatic-serialization attribute: it never waits before its
    loads.  Unsynchronised sequences equal the synchronised ones bit for bit."""
    ctx = gpu.ctx
    fs, m = 2_400_000, 1 << 23  # 64 MiB: many waves
    rng = np.random.default_rng(5)
    x = (rng.standard_normal(m) + 1j * rng.standard_normal(m)).astype(np.complex64)
    y, z = ctx.alloc(m * 8), ctx.alloc(m * 8)

    def run(sync):
        y.upload(x)
        st, st2 = H.NcoState(fs, 0.0), H.NcoState(fs, 0.0)
        ctx.shift(y.ptr, m, 1e5, st)           # scheme kernel: the window is open
        if sync:
            ctx.sync()
        ctx.rotate(y.ptr, m, 0.6 + 0.8j)        # outside the scheme, writes y
        if sync:
            ctx.sync()
        ctx.shift(y.ptr, m, -3e5, st2)          # reads what rotate wrote
        if sync:
            ctx.sync()
        ctx.add(z.ptr, [y.ptr, y.ptr], m)       # outside the scheme, reads y, writes z
        if sync:
            ctx.sync()
        ctx.shift(z.ptr, m, 2e5, st)            # reads what add wrote
        ctx.sync()
        return z.download(np.complex64, m)

    got, want = run(False), run(True)
    assert np.array_equal(bits(got), bits(want))

    # batched channelizer launch (outside the scheme) -> shift in place on one stream's output
    fmt, nfft, D, n, ns = H.FORMAT_I16, 1024, 16, 1 << 20, 8
    shifts = [-(1e6 + 10e3 * s) for s in range(ns)]
    Hf = O.filter_freq(O.lowpass_taps(255, 1 / 32), nfft)
    raws = [O.synth_raw(fmt, n, 8_000_000, -shifts[s], seed=s) for s in range(ns)]
    srcs = [ctx.to_device(r) for r in raws]
    per = n // D
    dsts = [ctx.alloc(per * 8) for _ in range(ns)]

    def run_chz(sync):
        chz = H.Channelizer(ctx, fmt, 8_000_000, shifts, Hf, D)
        st = H.NcoState(500_000, 0.0)
        for _ in range(2):  # the second exec takes the batched kernel
            chz.exec([s.ptr for s in srcs], n, [d.ptr for d in dsts], per)
            if sync:
                ctx.sync()
        ctx.shift(dsts[ns - 1].ptr, per, 1e4, st)
        ctx.sync()
        out = dsts[ns - 1].download(np.complex64, per)
        chz.close()
        return out

    assert np.array_equal(bits(run_chz(False)), bits(run_chz(True)))


# ---------------------------------------------------------------------------------------------
# fused chain
# ---------------------------------------------------------------------------------------------
CHAIN_CASES = [
    # fmt, fs, n, f0, taps, nfft, D
    (H.FORMAT_I8, 20_000_000, 1 << 19, 2.5e6, 255, 1024, 10),    # C2 shape, reduced length
    (H.FORMAT_I16, 61_440_000, 1 << 18, 7.68e6, 4095, 16384, 16),  # C3 shape, reduced length
    (H.FORMAT_U8, 2_400_000, 1 << 17, 300e3, 63, 256, 7),
    (H.FORMAT_I16, 2_000_000, 1 << 17, 250e3, 255, 4096, 3),
    (H.FORMAT_U8, 1_000_000, 1 << 16, 1e5, 31, 64, 1),
    # N = 1024 takes the warp-per-block kernel: odd D = full inverse, even D = folded 512-point inverse
    (H.FORMAT_U8, 2_400_000, 1 << 17, 300e3, 255, 1024, 1),
    (H.FORMAT_I16, 8_000_000, 1 << 17, 1e6, 255, 1024, 5),
    (H.FORMAT_I8, 20_000_000, 1 << 17, 2.5e6, 127, 1024, 2),
    (H.FORMAT_U8, 2_400_000, 1 << 18, 300e3, 255, 1024, 16),
    (H.FORMAT_I16, 61_440_000, 1 << 17, 7.68e6, 255, 1024, 100),
    (H.FORMAT_I8, 20_000_000, 1 << 17, 2.5e6, 255, 1024, 4097),
    # N = 16384 with D a multiple of 16 takes the CTA-per-block kernel with the folded 1024-point inverse
    (H.FORMAT_U8, 2_400_000, 1 << 18, 300e3, 1023, 16384, 16),
    (H.FORMAT_I8, 20_000_000, 1 << 18, 2.5e6, 4095, 16384, 32),
    (H.FORMAT_I16, 61_440_000, 1 << 20, 7.68e6, 4095, 16384, 48),
    (H.FORMAT_I16, 61_440_000, 1 << 17, 7.68e6, 2047, 16384, 8),   # D not a multiple of 16: generic kernel
    # N = K * 1024, K = 2, 4, 8: one CTA of K warps per block (chaink.cu), any decimation factor
    (H.FORMAT_I8, 20_000_000, 1 << 18, 2.5e6, 2047, 2048, 10),
    (H.FORMAT_U8, 2_400_000, 1 << 17, 300e3, 511, 2048, 1),
    (H.FORMAT_I16, 61_440_000, 1 << 18, 7.68e6, 4095, 4096, 16),
    (H.FORMAT_I8, 20_000_000, 1 << 17, 2.5e6, 1023, 4096, 7),
    (H.FORMAT_I16, 8_000_000, 1 << 18, 1e6, 8191, 8192, 12),
    (H.FORMAT_U8, 2_400_000, 1 << 18, 300e3, 255, 8192, 4097),
    (H.FORMAT_I8, 20_000_000, 1 << 17, 2.5e6, 127, 512, 10),       # generic kernel
]


@pytest.mark.parametrize("fmt,fs,n,f0,taps,nfft,D", CHAIN_CASES)
def test_chain_parity(gpu, fmt, fs, n, f0, taps, nfft, D):
    raw = O.synth_raw(fmt, n, fs, f0, seed=nfft + D)
    Hf = O.filter_freq(O.lowpass_taps(taps, 1 / (2 * max(D, 2))), nfft)
    want, ts_want = O.chain(raw, fmt, fs, -f0, Hf, D)
    got, ts = gpu.chain(raw, fmt, fs, -f0, Hf, D)
    assert got.shape == want.shape
    assert ts == ts_want
    err = O.rel_l2(got, want)
    assert err <= TOL, err
    # end-to-end entry point (H2D + kernel + D2H) gives the same samples
    got_h, ts_h = gpu.chain(raw, fmt, fs, -f0, Hf, D, host_path=True)
    assert ts_h == ts and np.array_equal(bits(got_h), bits(got))


def test_chain_equals_unfused_stages(gpu):
    """The fused kernel equals Convert -> Shift -> ConvolveFreq -> Decimate run as separate GPU
    stages (same arithmetic, so nearly bit-equal; assert 1e-6)."""
    fmt, fs, n, f0, nfft, D = H.FORMAT_I8, 20_000_000, 1 << 18, 2.5e6, 1024, 10
    raw = O.synth_raw(fmt, n, fs, f0, seed=77)
    Hf = O.filter_freq(O.lowpass_taps(255, 1 / 20), nfft)
    fused, _ = gpu.chain(raw, fmt, fs, -f0, Hf, D)
    y, _ = gpu.convert_shift(raw, fmt, -f0, fs, 0.0)
    z = gpu.convolution_reader(y, Hf)
    w = gpu.decimate_reader(z, D)
    assert O.rel_l2(fused, w) <= 1e-6


def test_chain_stream_continuity_and_resume(gpu):
    """Two consecutive buffers through one chain == one double-length buffer; ts get/set resumes."""
    fmt, fs, f0, nfft, D = H.FORMAT_I8, 20_000_000, 2.5e6, 1024, 10
    n = 1 << 18
    raw = O.synth_raw(fmt, 2 * n, fs, f0, seed=5)
    Hf = O.filter_freq(O.lowpass_taps(255, 1 / 20), nfft)
    whole, ts_whole = gpu.chain(raw, fmt, fs, -f0, Hf, D)
    a, ts_a = gpu.chain(raw[: 2 * n], fmt, fs, -f0, Hf, D)
    b, ts_b = gpu.chain(raw[2 * n:], fmt, fs, -f0, Hf, D, ts0=ts_a)
    assert ts_b == ts_whole
    # the two runs cut the accumulator into different segment tables, so the fixed-point phase
    # step is rounded at different places (~1e-12 turns): equal to fp32 rounding noise, not bits
    assert O.rel_l2(np.concatenate([a, b]), whole) <= 1e-7


def test_chain_pipelined_host_path(gpu):
    """hzsdr_chain_submit_host / wait_host: several pinned buffers in flight, same samples as the
    synchronous path, ts carried across submissions."""
    fmt, fs, f0, nfft, D = H.FORMAT_I8, 20_000_000, 2.5e6, 1024, 10
    n, nbuf = 1 << 17, 7
    raw = O.synth_raw(fmt, nbuf * n, fs, f0, seed=9)
    Hf = O.filter_freq(O.lowpass_taps(255, 1 / 20), nfft)
    want, ts_want = gpu.chain(raw, fmt, fs, -f0, Hf, D)
    ch = H.Chain(gpu.ctx, fmt, fs, -f0, Hf, D)
    per = ch.out_len(n)
    pin_in = H.PinnedBuffer(nbuf * n * 2)
    pin_out = H.PinnedBuffer(nbuf * per * 8)
    pin_in.view(np.int8)[:] = raw
    for b in range(nbuf):
        got = ch.submit_host(pin_in.ptr + b * n * 2, n, pin_out.ptr + b * per * 8, per)
        assert got == per
    ch.wait_host()
    out = pin_out.view(np.complex64).copy()
    assert ch.ts == ts_want
    assert O.rel_l2(out, want) <= 1e-7
    ch.close()


@pytest.mark.parametrize("nfft,D", [(256, 8), (1024, 8), (1024, 5), (2048, 6), (16384, 16)])  # every chain kernel's LSB variant
def test_chain_pluto_lsb_shift(gpu, nfft, D):
    fs, n = 4_000_000, 1 << 16
    raw12 = (O.synth_raw(O.FORMAT_I16, n, fs, 5e5, seed=12).astype(np.int32) >> 4).astype(np.int16)  # 12-bit LSB aligned
    Hf = O.filter_freq(O.lowpass_taps(63, 1 / 16), nfft)
    want, _ = O.chain(O.shift_lsb_to_msb_bits(raw12, 12), O.FORMAT_I16, fs, -5e5, Hf, D)
    got, _ = gpu.chain(raw12, H.FORMAT_I16, fs, -5e5, Hf, D, lsb_bits=12)
    assert O.rel_l2(got, want) <= TOL


@pytest.mark.parametrize("nfft,D", [(1024, 10), (1024, 3), (4096, 10)])
@pytest.mark.parametrize("shift", [0.0, 9.9e6, -27.3e6])  # no mixing at all; near and beyond the Nyquist rate (the phase wraps)
def test_chain_shift_extremes(gpu, nfft, D, shift):
    fmt, fs, n = H.FORMAT_I8, 20_000_000, 1 << 17
    raw = O.synth_raw(fmt, n, fs, 2.5e6, seed=31)
    Hf = O.filter_freq(O.lowpass_taps(255, 1 / 20), nfft)
    want, ts = O.chain(raw, fmt, fs, shift, Hf, D)
    got, ts_got = gpu.chain(raw, fmt, fs, shift, Hf, D)
    assert ts_got == ts and got.shape == want.shape
    assert O.rel_l2(got, want) <= TOL


def test_chain_rejects_ragged_and_bad_config(gpu):
    Hf = O.filter_freq(O.lowpass_taps(63, 0.1), 256)
    ch = H.Chain(gpu.ctx, H.FORMAT_U8, 1_000_000, 0.0, Hf, 4)
    assert ch.out_len(32768 * 3 + 100) == 3 * 8192
    d = gpu.ctx.alloc(1 << 20)
    with pytest.raises(H.HzsdrError) as ei:
        ch.exec(d.ptr, 32768 + 256, d.ptr, 1 << 16)
    assert ei.value.status == H.ERR_INVALID
    with pytest.raises(H.HzsdrError) as ei:
        ch.exec(d.ptr, 32768, d.ptr, 10)
    assert ei.value.status == H.ERR_DST_TOO_SMALL
    with pytest.raises(H.HzsdrError) as ei:
        H.Chain(gpu.ctx, H.FORMAT_U8, 1_000_000, 0.0, np.zeros(1000, dtype=np.complex64), 4)
    assert ei.value.status == H.ERR_UNSUPPORTED
    with pytest.raises(H.HzsdrError) as ei:
        H.Chain(gpu.ctx, H.FORMAT_C64, 1_000_000, 0.0, Hf, 4)
    assert ei.value.status == H.ERR_FORMAT_UNKNOWN


def test_chain_full_size_c2_properties(gpu):
    """BASELINE config 2 at full size (2^22 samples): size-independent properties -- output
    length, linearity in the input amplitude, and agreement with the CPU chain on a decimate
    block from the middle and from the end of the buffer."""
    fmt, fs, n, f0, nfft, D = H.FORMAT_I8, 20_000_000, 1 << 22, 2.5e6, 1024, 10
    raw = O.synth_raw(fmt, n, fs, f0, seed=2)
    Hf = O.filter_freq(O.lowpass_taps(255, 1 / 20), nfft)
    got, ts = gpu.chain(raw, fmt, fs, -f0, Hf, D)
    assert got.shape == (128 * 3276,)
    _, ts_want = CR.shift_ts(fs, n, 0.0, want_array=False)
    assert ts == ts_want
    # the oracle on the whole buffer takes a few seconds in numpy; check it all
    want, _ = O.chain(raw, fmt, fs, -f0, Hf, D)
    assert O.rel_l2(got, want) <= TOL
    for blk in (0, 64, 127):
        sl = slice(blk * 3276, (blk + 1) * 3276)
        assert O.rel_l2(got[sl], want[sl]) <= TOL
    # linearity: halving the input (exact in i8 for even codes) halves the output
    raw_even = (raw // 2 * 2).astype(np.int8)
    full, _ = gpu.chain(raw_even, fmt, fs, -f0, Hf, D)
    half, _ = gpu.chain((raw_even // 2).astype(np.int8), fmt, fs, -f0, Hf, D)
    assert O.rel_l2(2 * half.astype(np.complex128), full) <= 1e-6


@pytest.mark.parametrize("method", [H.FIR_OVERLAP_SAVE, H.FIR_POLYPHASE])
@pytest.mark.parametrize("ntaps,D", [(255, 10), (63, 4), (4095, 16), (1, 1), (17, 3)])
def test_fir_linear_convolution_streaming(gpu, method, ntaps, D):
    """hzsdr_fir_* (extension): z[n] = sum_k h[k] y[n-k] with history carried across ragged calls,
    out = z[D*i]; both methods against a direct complex128 FIR."""
    rng = np.random.default_rng(ntaps * 31 + D)
    n = 50_000 if ntaps < 1000 else 120_000
    y = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    h = (O.lowpass_taps(ntaps, 0.4 / D) * np.exp(0.3j * np.arange(ntaps))).astype(np.complex64) if ntaps > 1 else np.array([0.5 - 0.25j], np.complex64)
    want = O.fir_overlap_save_reference(y, h)[::D]
    fir = H.Fir(gpu.ctx, h, D, method)
    cuts = [0, 7, 7 + 12_345, 7 + 12_345 + 1, n]  # ragged, including a 1-sample call
    parts = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        src = gpu.ctx.to_device(y[a:b])
        dst = gpu.ctx.alloc(((b - a) // D + 2) * 8)
        got = fir.exec(src.ptr, b - a, dst.ptr, (b - a) // D + 2)
        parts.append(dst.download(np.complex64, got))
    out = np.concatenate(parts)
    assert out.shape == want.shape
    assert O.rel_l2(out, want) <= TOL
    fir.close()


@pytest.mark.parametrize("D", [16, 5])  # even: per-stream split tables, contiguous ranges per CTA; odd: the unpruned batch form
def test_channelizer_matches_per_stream_chains(gpu, D):
    """hzsdr_channelizer_*: many streams in one launch == each stream through its own chain, over
    two consecutive buffers (the first takes the long stream-start segment tables, the second the
    batched kernel), NCO times carried per stream."""
    fmt, fs, nfft, n, ns = H.FORMAT_I16, 8_000_000, 1024, 1 << 16, 12
    shifts = [-1e6 + 173e3 * s for s in range(ns)]
    Hf = O.filter_freq(O.lowpass_taps(255, 1 / 32), nfft)
    raws = [O.synth_raw(fmt, 2 * n, fs, -shifts[s], seed=40 + s) for s in range(ns)]
    chz = H.Channelizer(gpu.ctx, fmt, fs, shifts, Hf, D)
    per = n // 32768 * (32768 // D)
    outs = []
    for half in range(2):
        srcs = [gpu.ctx.to_device(r[half * 2 * n:(half + 1) * 2 * n]) for r in raws]
        dsts = [gpu.ctx.alloc(per * 8) for _ in range(ns)]
        assert chz.exec([s.ptr for s in srcs], n, [d.ptr for d in dsts], per) == per
        outs.append([d.download(np.complex64, per) for d in dsts])
    ts = chz.ts
    for s in range(ns):
        want, ts_want = O.chain(raws[s], fmt, fs, shifts[s], Hf, D)
        got = np.concatenate([outs[0][s], outs[1][s]])
        assert ts[s] == ts_want
        assert O.rel_l2(got, want) <= TOL
    chz.close()


def test_channelizer_host_path_matches_device_path(gpu):
    """hzsdr_channelizer_submit_host (pinned host buffers, streams staged across PCIe in groups,
    H2D / kernel / D2H overlapped) gives the samples and the carried NCO times of the device path,
    over two consecutive buffers, with a stream count that does not divide into the groups."""
    fmt, fs, nfft, D, n, ns = H.FORMAT_I16, 8_000_000, 1024, 16, 1 << 16, 13
    shifts = [-1e6 + 173e3 * s for s in range(ns)]
    Hf = O.filter_freq(O.lowpass_taps(255, 1 / 32), nfft)
    raws = [O.synth_raw(fmt, 2 * n, fs, -shifts[s], seed=70 + s) for s in range(ns)]
    per = n // 32768 * (32768 // D)
    dev, host = H.Channelizer(gpu.ctx, fmt, fs, shifts, Hf, D), H.Channelizer(gpu.ctx, fmt, fs, shifts, Hf, D)
    pin_in = [H.PinnedBuffer(n * 4) for _ in range(ns)]
    pin_out = [H.PinnedBuffer(per * 8) for _ in range(ns)]
    for half in range(2):
        parts = [r[half * 2 * n:(half + 1) * 2 * n] for r in raws]
        srcs = [gpu.ctx.to_device(p) for p in parts]
        dsts = [gpu.ctx.alloc(per * 8) for _ in range(ns)]
        assert dev.exec([s.ptr for s in srcs], n, [d.ptr for d in dsts], per) == per
        for b, p in zip(pin_in, parts):
            b.view(np.int16)[:] = p
        assert host.submit_host([b.ptr for b in pin_in], n, [b.ptr for b in pin_out], per) == per
        gpu.ctx.wait_host()
        for s in range(ns):
            assert np.array_equal(bits(pin_out[s].view(np.complex64)), bits(dsts[s].download(np.complex64, per))), (half, s)
    assert np.array_equal(dev.ts, host.ts)
    dev.close()
    host.close()


def test_beamform_host_path(gpu):
    """hzsdr_beamform_submit_host: channels in one pinned block (one 2-D copy per piece) and in
    separate pinned buffers, several pieces per call, two calls in flight before the wait -- the
    beam equals the device path's bit for bit."""
    nchan, n = 16, 1 << 21  # 64 MiB of raw samples: more than one staging piece
    w = O.beamform_angles(433e6, 30.0, [0.15 * c for c in range(nchan)])
    base = O.synth_raw(O.FORMAT_U8, n, 2_400_000, 1e5, seed=3)
    chans = [np.roll(base, 2 * 1013 * c) for c in range(nchan)]
    want = gpu.beamform(chans, H.FORMAT_U8, w)
    block = H.PinnedBuffer(nchan * n * 2)
    block.view(np.uint8).reshape(nchan, 2 * n)[:] = np.stack(chans)
    separate = [H.PinnedBuffer(n * 2) for _ in range(nchan)]
    for b, c in zip(separate, chans):
        b.view(np.uint8)[:] = c
    out1, out2 = H.PinnedBuffer(n * 8), H.PinnedBuffer(n * 8)
    gpu.ctx.beamform_submit_host(H.FORMAT_U8, [block.ptr + c * n * 2 for c in range(nchan)], w, n, out1.ptr)
    gpu.ctx.beamform_submit_host(H.FORMAT_U8, [b.ptr for b in separate], w, n, out2.ptr)
    gpu.ctx.wait_host()
    assert np.array_equal(bits(out1.view(np.complex64)), bits(want))
    assert np.array_equal(bits(out2.view(np.complex64)), bits(want))
    with pytest.raises(H.HzsdrError):
        gpu.ctx.beamform_submit_host(H.FORMAT_C64, [block.ptr], w[:1], 4, out1.ptr)


def test_chain_full_size_c3_properties(gpu):
    """BASELINE config 3 at full size (i16, 2^24 samples, N = 16384, 4095 taps, x16): output length,
    carried ts bit-equal to the compiled serial loop, the oracle on the first and last 2^19 input
    samples (each FFT block is independent, so a slice of blocks is checkable on its own once the
    NCO time at its start is known), and linearity."""
    fmt, fs, n, f0, nfft, D = H.FORMAT_I16, 61_440_000, 1 << 24, 7.68e6, 16384, 16
    raw = O.synth_raw(fmt, n, fs, f0, seed=3)
    Hf = O.filter_freq(O.lowpass_taps(4095, 1 / 32), nfft)
    got, ts = gpu.chain(raw, fmt, fs, -f0, Hf, D)
    assert got.shape == (n // D,)
    _, ts_want = CR.shift_ts(fs, n, 0.0, want_array=False)
    assert ts == ts_want
    m = 1 << 19
    head, _ = O.chain(raw[: 2 * m], fmt, fs, -f0, Hf, D)
    assert O.rel_l2(got[: m // D], head) <= TOL
    _, ts_tail0 = CR.shift_ts(fs, n - m, 0.0, want_array=False)
    tail, _ = O.chain(raw[2 * (n - m):], fmt, fs, -f0, Hf, D, ts0=ts_tail0)
    assert O.rel_l2(got[-(m // D):], tail) <= TOL
    half, _ = gpu.chain((raw // 2 * 2 // 2).astype(np.int16), fmt, fs, -f0, Hf, D)
    full, _ = gpu.chain((raw // 2 * 2).astype(np.int16), fmt, fs, -f0, Hf, D)
    assert O.rel_l2(2 * half.astype(np.complex128), full) <= 2e-6


def test_channelizer_full_length_c5_streams(gpu):
    """BASELINE config 5's stream shape (i16, 2^20 samples, 255 taps, x16), 6 of the 512 streams:
    three consecutive buffers through the batched kernel against the oracle per stream."""
    fmt, fs, nfft, D, n, ns = H.FORMAT_I16, 61_440_000, 1024, 16, 1 << 20, 6
    shifts = [-(1e6 + 10e3 * s) for s in range(ns)]
    Hf = O.filter_freq(O.lowpass_taps(255, 1 / 32), nfft)
    raws = [O.synth_raw(fmt, 3 * n, fs, -shifts[s], seed=70 + s) for s in range(ns)]
    chz = H.Channelizer(gpu.ctx, fmt, fs, shifts, Hf, D)
    per = n // D
    outs = [[] for _ in range(ns)]
    for part in range(3):
        srcs = [gpu.ctx.to_device(r[part * 2 * n:(part + 1) * 2 * n]) for r in raws]
        dsts = [gpu.ctx.alloc(per * 8) for _ in range(ns)]
        assert chz.exec([s.ptr for s in srcs], n, [d.ptr for d in dsts], per) == per
        for s in range(ns):
            outs[s].append(dsts[s].download(np.complex64, per))
    ts = chz.ts
    for s in range(ns):
        want, ts_want = O.chain(raws[s], fmt, fs, shifts[s], Hf, D)
        assert ts[s] == ts_want
        assert O.rel_l2(np.concatenate(outs[s]), want) <= TOL
    chz.close()


def test_beamform_full_size_c4(gpu):
    """BASELINE config 4 at full size (64 u8 channels x 2^20 samples): against the oracle on the
    whole beam, plus the algebraic property that the beam is linear in the weights."""
    nchan, n = 64, 1 << 20
    w = O.beamform_angles(433e6, 30.0, [0.15 * c for c in range(nchan)])
    base = [O.synth_raw(O.FORMAT_U8, n, 2_400_000, 1e5, seed=c, phase=0.37 * c) for c in range(8)]
    chans = [np.roll(base[c % 8], 2 * 997 * (c // 8)) for c in range(nchan)]  # 64 distinct channels from 8 draws
    got = gpu.beamform(chans, H.FORMAT_U8, w)
    want = O.beamform(chans, O.FORMAT_U8, w)
    assert O.rel_l2(got, want) <= TOL
    g2 = gpu.beamform(chans, H.FORMAT_U8, (2 * w).astype(np.complex64))
    assert O.rel_l2(g2, 2 * got.astype(np.complex128)) <= 1e-6


# ---------------------------------------------------------------------------------------------
# K8 beamform
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("fmt,nchan,n", [(H.FORMAT_U8, 4, 4096), (H.FORMAT_U8, 64, 1 << 16), (H.FORMAT_I16, 7, 10000), (H.FORMAT_U8, 3, 4),
                                          (H.FORMAT_I8, 70, 2048)])
def test_beamform_parity(gpu, fmt, nchan, n):
    d = 0.15
    w = O.beamform_angles(433e6, 30.0, [d * i for i in range(nchan)])
    chans = [O.synth_raw(fmt, n, 2_400_000, 1e5, seed=100 + c, phase=0.37 * c) for c in range(nchan)]
    want = O.beamform(chans, fmt, w)
    got = gpu.beamform(chans, fmt, w)
    err = O.rel_l2(got, want)
    assert err <= TOL, err
    assert err <= 2e-6, err


def test_beamform_unit_weights_is_ordered_sum(gpu):
    """All weights 1 (Multiply skipped, stream/multiply.go:59-62): the result is the ordered fp32
    sum of the converted channels -- bit-exact."""
    n, nchan = 8192, 16
    # i8: the conversion scale 2^-7 folds into the weight exactly, so the kernel's FMA chain is the
    # reference's ordered fp32 sum of exactly converted samples
    chans = [O.synth_raw(O.FORMAT_I8, n, 2_400_000, 1e5, seed=c) for c in range(nchan)]
    w = np.ones(nchan, dtype=np.complex64)
    assert np.array_equal(bits(gpu.beamform(chans, H.FORMAT_I8, w)), bits(O.beamform(chans, O.FORMAT_I8, w)))
    # u8 / i16: the fold rounds differently from the reference's division; tolerance path
    chans = [O.synth_raw(O.FORMAT_U8, n, 2_400_000, 1e5, seed=c) for c in range(nchan)]
    assert O.rel_l2(gpu.beamform(chans, H.FORMAT_U8, w), O.beamform(chans, O.FORMAT_U8, w)) <= 5e-7


# ---------------------------------------------------------------------------------------------
# pinned ring + async H2D
# ---------------------------------------------------------------------------------------------
def test_ring_roundtrip_and_underrun(gpu):
    ctx = gpu.ctx
    slot = 32768
    ring = H.Ring(ctx, H.FORMAT_I8, 4, slot)
    with pytest.raises(H.HzsdrError) as ei:
        ring.read()
    assert ei.value.status == H.ERR_RING_UNDERRUN
    bufs = [O.synth_raw(O.FORMAT_I8, slot, 20_000_000, 1e6, seed=s) for s in range(6)]
    dst = ctx.alloc(slot * 8)
    for i in range(3):
        ring.write(bufs[i])
    for i in range(3):
        p, n = ring.read()
        assert n == slot
        ctx.convert_to_c64(H.FORMAT_I8, p, n, dst.ptr, n)
        ring.read_done()
        assert np.array_equal(bits(dst.download(np.complex64, n)), bits(O.convert_i8_to_c64(bufs[i])))
    # overrun: 5 writes into 4 slots drop the oldest (stream/ring.go:170-186)
    for i in range(1, 6):
        ring.write(bufs[i])
    p, n = ring.read()
    ctx.convert_to_c64(H.FORMAT_I8, p, n, dst.ptr, n)
    ring.read_done()
    assert np.array_equal(bits(dst.download(np.complex64, n)), bits(O.convert_i8_to_c64(bufs[2])))
    ring.close()
