"""CPU restatement (numpy) of hz.tools/sdr's IQ sample chain.  TEST INFRASTRUCTURE ONLY.

This module is the *oracle*: the checker that the CUDA path (libhzsdrcuda.so) is
compared against.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The
product path never does; it fails loudly when the CUDA library is missing.

Every function cites the reference file:line (relative to the hztools/go-sdr tree)
whose arithmetic it restates.  The reference is Go + Go-assembler; no Go toolchain
exists in this image, so the restatement is pinned against the reference's own
known-answer tests (transcribed into ``tests/golden/reference_kats.json``) by
``tests/test_oracle.py``.

Parity status per function:
  * conversion, scale, rotate, add, decimate, downsample, lookup, beamform weights:
    pinned by the reference's KATs (bit-exact where the reference asserts equality).
  * shift: pinned by the reference's up/down round-trip test (1e-4); the forward
    values follow ``stream/shifter.go:73-84`` op-for-op in fp64.
  * FFT / ConvolveFreq / ConvolutionReader values: **parity unpinned** -- the
    reference has no in-tree FFT (``fft/fft.go:45-59`` is an interface, no planner
    is named in ``go.mod:5-9``) and no convolution test.  Convention adopted:
    forward kernel e^{-2 pi i k n / N}, backward e^{+...}, both unnormalised (FFTW
    convention, consistent with ``rtl/kerberos/internal/reader.go:54-56``).
"""
from __future__ import annotations

import math

import numpy as np

try:  # scipy is in the image; numpy.fft is the fallback (also pocketfft)
    import scipy.fft as _fft
except Exception:  # pragma: no cover
    _fft = np.fft

# sdr.SampleFormat ids, iq.go:113-129
FORMAT_C64 = 1
FORMAT_U8 = 2
FORMAT_I16 = 3
FORMAT_I8 = 4

FORMAT_SIZE = {FORMAT_C64: 8, FORMAT_U8: 2, FORMAT_I16: 4, FORMAT_I8: 2}  # iq.go:99-110
FORMAT_DTYPE = {FORMAT_C64: np.complex64, FORMAT_U8: np.uint8, FORMAT_I16: np.int16, FORMAT_I8: np.int8}

DECIMATE_BLOCK = 32 * 1024  # stream/decimate.go:41-42, stream/downsample.go:54-55
CONVERT_BLOCK = 32 * 1024  # stream/convert.go:43-44

TAU = math.pi * 2  # stream/shifter.go:70


class SdrError(Exception):
    """Mirror of the reference's sentinel errors (iq.go:27-39, conv.go:30)."""


class ErrDstTooSmall(SdrError):
    pass


class ErrSampleFormatMismatch(SdrError):
    pass


class ErrSampleFormatUnknown(SdrError):
    pass


# --------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------

def _pairs(raw: np.ndarray, dtype) -> np.ndarray:
    """View an interleaved IQ buffer as (n, 2)."""
    raw = np.ascontiguousarray(raw, dtype=dtype)
    return raw.reshape(-1, 2)


def _to_c64(re: np.ndarray, im: np.ndarray) -> np.ndarray:
    out = np.empty(re.shape[0], dtype=np.complex64)
    out.real = re
    out.imag = im
    return out


def go_complex64_mul(a: np.ndarray, b) -> np.ndarray:
    """Go's complex64 * complex64.

    The gc compiler widens both operands to float64, forms the four products and the
    sum/difference in float64, and narrows each component to float32 (external compiler
    behaviour; reference call sites internal/simd/mult.go:29-33 and
    stream/shifter.go:82).  Products of two fp32 values are exact in fp64, so each
    component is round32(round64(exact)).
    """
    a = np.asarray(a, dtype=np.complex64)
    b = np.asarray(b, dtype=np.complex64)
    ar = a.real.astype(np.float64)
    ai = a.imag.astype(np.float64)
    br = b.real.astype(np.float64)
    bi = b.imag.astype(np.float64)
    re = (ar * br - ai * bi).astype(np.float32)
    im = (ar * bi + ai * br).astype(np.float32)
    return _to_c64(re, im)


# --------------------------------------------------------------------------------------
# a2-a4: integer -> complex64 conversion
# --------------------------------------------------------------------------------------

def convert_u8_to_c64(raw: np.ndarray) -> np.ndarray:
    """iq_u8.go:111-121 (pure Go) == iq_u8_amd64.s:71-89 (SUBPS then DIVPS).

    (float32(b) - 127.5) / 127.5, both ops IEEE fp32.
    """
    p = _pairs(raw, np.uint8).astype(np.float32)
    v = (p - np.float32(127.5)) / np.float32(127.5)
    return _to_c64(v[:, 0], v[:, 1])


def convert_i8_to_c64(raw: np.ndarray) -> np.ndarray:
    """iq_i8.go:107-119: float32(b) / 128."""
    p = _pairs(raw, np.int8).astype(np.float32)
    v = p / np.float32(128)
    return _to_c64(v[:, 0], v[:, 1])


def convert_i16_to_c64(raw: np.ndarray) -> np.ndarray:
    """iq_i16.go:141-145: float32(v) / math.MaxInt16, an IEEE fp32 division."""
    p = _pairs(raw, np.int16).astype(np.float32)
    v = p / np.float32(32767)
    return _to_c64(v[:, 0], v[:, 1])


def shift_lsb_to_msb_bits(raw: np.ndarray, bits: int) -> np.ndarray:
    """iq_i16.go:103-111: left shift every component by 16-bits (wrapping int16)."""
    p = np.ascontiguousarray(raw, dtype=np.int16)
    return (p.astype(np.uint16) << np.uint16(16 - bits)).astype(np.uint16).view(np.int16)


def convert_to_c64(raw: np.ndarray, fmt: int) -> np.ndarray:
    """conv.go:55-93 restricted to dst = C64 (the hot path)."""
    if fmt == FORMAT_U8:
        return convert_u8_to_c64(raw)
    if fmt == FORMAT_I8:
        return convert_i8_to_c64(raw)
    if fmt == FORMAT_I16:
        return convert_i16_to_c64(raw)
    if fmt == FORMAT_C64:  # same format is a copy, conv.go:56-58
        return np.array(raw, dtype=np.complex64, copy=True)
    raise ErrSampleFormatUnknown(fmt)


def convert_buffer(dst_len: int, raw: np.ndarray, fmt: int) -> np.ndarray:
    """ConvertBuffer's length rule, conv.go:60-62: src longer than dst is an error."""
    n = np.asarray(raw).size // (1 if fmt == FORMAT_C64 else 2)
    if n > dst_len:
        raise ErrDstTooSmall()
    return convert_to_c64(raw, fmt)


# --------------------------------------------------------------------------------------
# the rest of ConvertBuffer's 4x4 matrix (conv.go:36-46)
# --------------------------------------------------------------------------------------

def _trunc_wrap(x: np.ndarray, bits: int, signed: bool) -> np.ndarray:
    """Go's float32 -> small integer conversion as the amd64 compiler does it: truncate toward
    zero to int32, keep the low `bits` bits (in-range values are simply truncated)."""
    t = np.trunc(x.astype(np.float64)).astype(np.int64) & ((1 << bits) - 1)
    dt = {(8, False): np.uint8, (8, True): np.int8, (16, True): np.int16}[(bits, signed)]
    return t.astype(np.uint64).astype({8: np.uint8, 16: np.uint16}[bits]).view(dt)


def convert_from_c64(buf: np.ndarray, fmt: int) -> np.ndarray:
    """SamplesC64.ToU8 / ToI16 / ToI8, iq_c64.go:77-117.  fp32 arithmetic, separate roundings
    (gc does not fuse on amd64), truncation toward zero.  Returns integer pairs (n, 2)."""
    buf = np.asarray(buf, dtype=np.complex64)
    x = np.stack([buf.real, buf.imag], axis=1).astype(np.float32)
    if fmt == FORMAT_U8:  # uint8(real*127.5 + 127.5)
        return _trunc_wrap(x * np.float32(127.5) + np.float32(127.5), 8, False)
    if fmt == FORMAT_I16:  # int16(real * math.MaxInt16)
        return _trunc_wrap(x * np.float32(32767), 16, True)
    if fmt == FORMAT_I8:  # int8(real * math.MaxInt8)
        return _trunc_wrap(x * np.float32(127), 8, True)
    raise ErrSampleFormatUnknown(fmt)


def convert_int(raw: np.ndarray, src: int, dst: int) -> np.ndarray:
    """Integer <-> integer conversions: iq_u8.go:73-101, iq_i8.go:73-97, iq_i16.go:116-134,150-162."""
    if src == FORMAT_U8 and dst == FORMAT_I8:   # int8(int16(b) - 128)
        return (np.ascontiguousarray(raw, np.uint8) ^ np.uint8(0x80)).view(np.int8)
    if src == FORMAT_I8 and dst == FORMAT_U8:   # uint8(int16(b) + 128)
        return (np.ascontiguousarray(raw, np.int8).view(np.uint8) ^ np.uint8(0x80))
    if src == FORMAT_U8 and dst == FORMAT_I16:  # int16((int32(b) << 8) - 32768)
        return ((np.ascontiguousarray(raw, np.uint8).astype(np.int32) << 8) - 32768).astype(np.int16)
    if src == FORMAT_I8 and dst == FORMAT_I16:  # int16(b) << 8
        return (np.ascontiguousarray(raw, np.int8).astype(np.int16) << 8).astype(np.int16)
    if src == FORMAT_I16 and dst == FORMAT_U8:  # uint8(uint16(int32(v)+32768) >> 8)
        return (((np.ascontiguousarray(raw, np.int16).astype(np.int32) + 32768) & 0xffff) >> 8).astype(np.uint8)
    if src == FORMAT_I16 and dst == FORMAT_I8:  # int8(v >> 8)
        return (np.ascontiguousarray(raw, np.int16) >> 8).astype(np.int8)
    raise ErrSampleFormatUnknown((src, dst))


def multiply_lut_i8(raw: np.ndarray, m: complex) -> np.ndarray:
    """int8MultiplyReader, stream/multiply.go:180-251: the table is Convert -> Multiply -> Convert of
    the identity (every IQ byte pair), reads are pure lookups."""
    ident = lookup_identity_u8().view(np.int8)
    tab = convert_from_c64(rotate(convert_i8_to_c64(ident.reshape(-1)), m), FORMAT_I8)
    return lookup(tab, raw)


def multiply_lut_u8(raw: np.ndarray, m: complex) -> np.ndarray:
    """uint8MultiplyReader, stream/multiply.go:91-172, bug for bug: the table has 65535 entries
    indexed by x0*255 + x1 (:106-108), so (x0, 255) collides with (x0+1, 0); entries are written in
    loop order real 0..255, imag 0..256 (uint8 wraps 256 to 0)."""
    ubuf = np.zeros((65535, 2), dtype=np.uint8)
    for realv in range(256):
        for imagv in range(257):
            v = (realv & 0xff, imagv & 0xff)
            ubuf[v[0] * 255 + v[1]] = v
    tab = convert_from_c64(rotate(convert_u8_to_c64(ubuf.reshape(-1)), m), FORMAT_U8)
    p = np.ascontiguousarray(raw, np.uint8).reshape(-1, 2).astype(np.int64)
    return tab[p[:, 0] * 255 + p[:, 1]]


# --------------------------------------------------------------------------------------
# a14: 65536-entry lookup table
# --------------------------------------------------------------------------------------

def lookup_index(raw: np.ndarray) -> np.ndarray:
    """iq_lookup_table.go:56-64: index = the IQ pair reinterpreted as a native-endian
    (little-endian on amd64) uint16, i.e. I + 256*Q on the raw bytes."""
    b = np.ascontiguousarray(raw).view(np.uint8).reshape(-1, 2)
    return b[:, 0].astype(np.uint32) | (b[:, 1].astype(np.uint32) << 8)


def lookup(table: np.ndarray, raw: np.ndarray) -> np.ndarray:
    """iq_lookup_table.go:198-251: dst[i] = tab[index(src[i])].

    ``table`` has 65536 samples; for integer formats it is shaped (65536, 2)."""
    return np.asarray(table)[lookup_index(raw)]


def lookup_identity_u8() -> np.ndarray:
    """iq_lookup_table.go:69-77."""
    i = np.arange(65536, dtype=np.uint16)
    return i.view(np.uint8).reshape(-1, 2).copy()


# --------------------------------------------------------------------------------------
# a6: Shift (NCO mixer with a serially-rounded fp64 time accumulator)
# --------------------------------------------------------------------------------------

def shift_ts_serial(sample_rate: int, n: int, ts0: float = 0.0):
    """The accumulator of stream/shifter.go:73-79, literally (pure-Python loop; small n)."""
    inc = 1.0 / float(sample_rate)
    ts = float(ts0)
    out = np.empty(n, dtype=np.float64)
    for j in range(n):
        ts += inc
        if ts > TAU:
            ts -= TAU
        out[j] = ts
    return out, ts


def shift_ts(sample_rate: int, n: int, ts0: float = 0.0):
    """Same values as :func:`shift_ts_serial`, vectorised.

    np.add.accumulate on a contiguous float64 vector is a sequential left-to-right sum,
    i.e. the same chain of rounded additions as the Go loop; the 2*pi wrap is handled by
    restarting the accumulation at each wrap.  Checked against the serial loop in
    tests/test_oracle.py."""
    inc = 1.0 / float(sample_rate)
    out = np.empty(n, dtype=np.float64)
    ts = float(ts0)
    j = 0
    while j < n:
        # upper bound on the number of steps before the wrap
        room = int((TAU - ts) / inc) + 2
        m = min(n - j, max(room, 1))
        seg = np.full(m + 1, inc, dtype=np.float64)
        seg[0] = ts
        acc = np.add.accumulate(seg)[1:]
        over = np.nonzero(acc > TAU)[0]
        if over.size == 0:
            out[j:j + m] = acc
            ts = float(acc[-1])
            j += m
        else:
            k = int(over[0])
            out[j:j + k] = acc[:k]
            ts = float(acc[k]) - TAU  # stream/shifter.go:77-79
            out[j + k] = ts
            j += k + 1
    return out, ts


def shift_segments(sample_rate: int, n: int, ts0: float = 0.0):
    """Closed-form description of the accumulator: a list of (j0, count, base, step) with
    ts[j0+k] = base + (k+1)*step exactly (fp64) for k < count.

    This is the *host logic the CUDA library mirrors* (csrc/nco_segments.h); it lives here
    so the CPU tests can pin it bit-for-bit against the serial loop.  Within one fp64
    binade every `ts += inc` adds inc rounded to that binade's grid, a constant, unless
    inc/ulp(ts) has fractional part exactly one half (round-half-even alternates); those
    binades, binade crossings and the 2*pi wrap are emitted as single real steps.
    """
    inc = 1.0 / float(sample_rate)
    segs = []
    ts = float(ts0)
    j = 0
    while j < n:
        m = 0
        if ts > 0.0:
            u = math.ulp(ts)
            q = inc / u
            if q < 2.0 ** 52 and (q % 1.0) != 0.5:
                step = (ts + inc) - ts
                top = min(math.ldexp(1.0, math.frexp(ts)[1]), TAU)
                if step > 0.0:
                    m = max(0, min(int((top - ts) / step) - 2, n - j))
        if m > 0:
            segs.append((j, m, ts, step))
            ts = ts + m * step
            j += m
        else:
            nxt = ts + inc
            if nxt > TAU:
                nxt -= TAU
            segs.append((j, 1, nxt, 0.0))
            ts = nxt
            j += 1
    return segs, ts


def expand_segments(segs, n: int) -> np.ndarray:
    out = np.empty(n, dtype=np.float64)
    for j0, m, base, step in segs:
        if step == 0.0:
            out[j0:j0 + m] = base
        else:
            out[j0:j0 + m] = base + np.arange(1, m + 1, dtype=np.float64) * step
    return out


def shift_buffer(buf: np.ndarray, freq: float, sample_rate: int, ts0: float = 0.0):
    """stream/shifter.go:73-84.  Returns (shifted c64 buffer, carried ts)."""
    buf = np.asarray(buf, dtype=np.complex64)
    ts, ts_end = shift_ts(sample_rate, buf.shape[0], ts0)
    ang = (TAU * float(freq)) * ts  # ((tau*shift)*ts), left-to-right as Go evaluates it
    rot = _to_c64(np.cos(ang).astype(np.float32), np.sin(ang).astype(np.float32))
    return go_complex64_mul(buf, rot), ts_end


# --------------------------------------------------------------------------------------
# a10-a12: Multiply / Gain / Add
# --------------------------------------------------------------------------------------

def rotate(buf: np.ndarray, m: complex) -> np.ndarray:
    """internal/simd/mult.go:29-33 via iq_c64.go:128-130; stream/multiply.go:59-62 skips m==1."""
    m64 = np.complex64(m)
    buf = np.asarray(buf, dtype=np.complex64)
    if m64 == np.complex64(1):
        return buf.copy()
    return go_complex64_mul(buf, m64)


def scale(buf: np.ndarray, r: float) -> np.ndarray:
    """internal/simd/mult.go:25-27 / mult_simd_amd64.s:47-54: re*=r, im*=r in fp32."""
    buf = np.asarray(buf, dtype=np.complex64)
    r32 = np.float32(r)
    return _to_c64(buf.real * r32, buf.imag * r32)


def add(*bufs: np.ndarray) -> np.ndarray:
    """stream/add.go:115-119,165-168: out = ((0 + b0) + b1) + ... in fp32, reader order."""
    out = np.zeros_like(np.asarray(bufs[0], dtype=np.complex64))
    for b in bufs:
        b = np.asarray(b, dtype=np.complex64)
        out = _to_c64(out.real + b.real, out.imag + b.imag)
    return out


def add_int(*bufs: np.ndarray) -> np.ndarray:
    """stream/add.go:95-113: wrapping integer adds for i8 / i16."""
    dt = np.asarray(bufs[0]).dtype
    out = np.zeros_like(bufs[0])
    with np.errstate(over="ignore"):
        for b in bufs:
            out = (out.astype(np.int64) + np.asarray(b).astype(np.int64)).astype(dt)
    return out


# --------------------------------------------------------------------------------------
# a8-a9: Decimate / Downsample
# --------------------------------------------------------------------------------------

def decimate_buffer(frm: np.ndarray, factor: int, to_len: int | None = None) -> np.ndarray:
    """stream/decimate.go:59-101: to[i] = from[factor*i], i < len(from)/factor.
    ``frm`` is c64 (n,) or integer pairs (n,2).  `offset` is ignored by the reference."""
    n = frm.shape[0]
    m = n // int(factor)
    if to_len is not None and to_len < m:
        raise ErrDstTooSmall()
    return frm[: m * factor : factor].copy()


def decimate_reader(stream: np.ndarray, factor: int, block: int = DECIMATE_BLOCK) -> np.ndarray:
    """stream/decimate.go:34-51 over a whole stream: fixed `block`-sample input blocks
    (ReadFull, trailing partial block dropped: stream/read_transformer.go:121-125), the
    decimation phase restarts at every block."""
    nblk = stream.shape[0] // block
    outs = [decimate_buffer(stream[b * block:(b + 1) * block], factor) for b in range(nblk)]
    if not outs:
        return stream[:0].copy()
    return np.concatenate(outs)


def downsample_buffer(frm: np.ndarray, factor: int, fmt: int = FORMAT_C64) -> np.ndarray:
    """stream/downsample.go:68-127: out[i] = (sum_j from[i*f+j]) / float32(f); the sum is
    sequential in complex64 (fp32 adds), u8 / i16 inputs are converted on the fly."""
    if fmt != FORMAT_C64:
        frm = convert_to_c64(frm, fmt)
    frm = np.asarray(frm, dtype=np.complex64)
    f = int(factor)
    m = frm.shape[0] // f
    x = frm[: m * f].reshape(m, f)
    re = np.zeros(m, dtype=np.float32)
    im = np.zeros(m, dtype=np.float32)
    for j in range(f):  # sequential fp32 accumulation, downsample.go:115-117
        re = re + x[:, j].real
        im = im + x[:, j].imag
    return _to_c64(re / np.float32(f), im / np.float32(f))


def downsample_reader(stream: np.ndarray, factor: int, fmt: int = FORMAT_C64,
                      block: int = DECIMATE_BLOCK) -> np.ndarray:
    """stream/downsample.go:47-64 over a whole stream (32768-sample blocks)."""
    per = 1 if fmt == FORMAT_C64 else 2
    s = np.asarray(stream).reshape(-1) if fmt != FORMAT_C64 else np.asarray(stream)
    n = s.shape[0] // per
    nblk = n // block
    outs = [downsample_buffer(s[b * block * per:(b + 1) * block * per], factor, fmt) for b in range(nblk)]
    if not outs:
        return np.zeros(0, dtype=np.complex64)
    return np.concatenate(outs)


# --------------------------------------------------------------------------------------
# a7: FFT planner contract and ConvolveFreq / ConvolutionReader   (PARITY UNPINNED)
# --------------------------------------------------------------------------------------

def fft_forward(x: np.ndarray) -> np.ndarray:
    """fft.Planner(..., fft.Forward) (fft/fft.go:32-59): unnormalised DFT, e^{-2 pi i kn/N}.
    Computed in complex128 and rounded once to complex64 (the Plan's output buffer type)."""
    return _fft.fft(np.asarray(x, dtype=np.complex64).astype(np.complex128), axis=-1).astype(np.complex64)


def fft_backward(x: np.ndarray) -> np.ndarray:
    """fft.Planner(..., fft.Backward): unnormalised inverse DFT, e^{+2 pi i kn/N}."""
    x = np.asarray(x, dtype=np.complex64).astype(np.complex128)
    n = x.shape[-1]
    return (_fft.ifft(x, axis=-1) * n).astype(np.complex64)


def convolve_freq(src: np.ndarray, freq: np.ndarray) -> np.ndarray:
    """fft/convolution.go:183-191: freq1 = FFT(src); freq1[i] *= freq[i]; dst = IFFT(freq1).
    `src` may be (nblocks, N); intermediates are complex64 buffers as in the reference."""
    f1 = fft_forward(src)
    f1 = go_complex64_mul(f1.reshape(-1), np.broadcast_to(np.asarray(freq, dtype=np.complex64), f1.shape).reshape(-1)).reshape(f1.shape)
    return fft_backward(f1)


def fft_convolve(iq1: np.ndarray, iq2: np.ndarray, cross_correlate: bool = False) -> np.ndarray:
    """fft.Convolve / fft.CrossCorrelate, fft/convolution.go:97-139, with complex64 intermediates."""
    f1, f2 = fft_forward(iq1), fft_forward(iq2)
    if cross_correlate:
        f2 = np.conj(f2)
    prod = go_complex64_mul(f1.reshape(-1), f2.reshape(-1)).reshape(f1.shape)
    return fft_backward(prod)


# ---- coherent-receiver helpers: rtl/kerberos/internal -----------------------------------------

def fftshift_and_scale(data: np.ndarray, scale: float) -> np.ndarray:
    """FFTShiftAndScale, rtl/kerberos/internal/reader.go:50-64: halves exchanged, every component
    divided by `scale` in fp32.  Returns a new array (the reference works in place)."""
    d = np.asarray(data, dtype=np.complex64)
    half = d.shape[-1] // 2
    out = d.copy()
    sc = np.float32(scale)
    lo, hi = d[..., :half], d[..., half:2 * half]
    out[..., :half] = (hi.real / sc) + 1j * (hi.imag / sc)
    out[..., half:2 * half] = (lo.real / sc) + 1j * (lo.imag / sc)
    return out.astype(np.complex64)


def graft(iq_bufs: np.ndarray) -> np.ndarray:
    """One pass of graftReader.do's loop, rtl/kerberos/internal/graft.go:96-125: forward transform of
    every reader's buffer into its slice of freqBuf, FFTShiftAndScale(slice, fftSize), ONE backward
    transform over the concatenation.  iq_bufs: (n_readers, fft_size) complex64."""
    x = np.asarray(iq_bufs, dtype=np.complex64)
    nr, size = x.shape
    freq = fftshift_and_scale(fft_forward(x), float(size))
    return fft_backward(freq.reshape(1, nr * size)).reshape(-1)


def cross_correlate(buf1: np.ndarray, buf2: np.ndarray) -> np.ndarray:
    """CrossCorrelater.run / Correlate, rtl/kerberos/internal/align.go:44-75: conjMult of the two
    spectra (complex64 multiply), backward transform."""
    return fft_convolve(np.asarray(buf1)[None, :], np.asarray(buf2)[None, :], cross_correlate=True)[0]


def correlate_peak(cc: np.ndarray) -> int:
    """checkAlignment's search, align.go:125-146: first index of maximum fp32 power, exact zeros
    skipped, wrapped past n/2; -1 when everything is zero."""
    c = np.asarray(cc, dtype=np.complex64)
    re, im = c.real.astype(np.float32), c.imag.astype(np.float32)
    pw = (re * re + im * im).astype(np.float32)
    valid = ~((re == 0) & (im == 0))
    if not valid.any():
        return -1
    pw = np.where(valid, pw, -np.inf)
    i = int(np.argmax(pw))  # first maximum
    n = c.shape[0]
    return i - n if i > n // 2 else i


def phase_offsets(bufs: np.ndarray) -> np.ndarray:
    """PhaseOffsets, align.go:244-272: mean over i of Phase(complex128(conjMult(b0[i], bj[i]))) per
    channel j, fp64 accumulation, then Rect(1, mean).  The reference overwrites channel 0's SUM with
    1 before dividing by the length (align.go:265), so element 0 is Rect(1, 1/n): reproduced."""
    b = np.asarray(bufs, dtype=np.complex64)
    nchan, n = b.shape
    phases = np.zeros(nchan, dtype=np.float64)
    for j in range(1, nchan):
        m = go_complex64_mul(b[0], np.conj(b[j]))
        phases[j] = float(np.sum(np.arctan2(m.imag.astype(np.float64), m.real.astype(np.float64))))
    phases[0] = 1.0
    phases /= float(n)
    return (np.cos(phases) + 1j * np.sin(phases)).astype(np.complex64)


def convolution_reader(stream: np.ndarray, filt: np.ndarray) -> np.ndarray:
    """stream/convolution.go:57-81: block-circular -- each len(filter)-sample block is
    convolved on its own, no history; the trailing partial block is dropped."""
    n = len(filt)
    nblk = stream.shape[0] // n
    if nblk == 0:
        return np.zeros(0, dtype=np.complex64)
    x = np.asarray(stream[: nblk * n], dtype=np.complex64).reshape(nblk, n)
    return convolve_freq(x, filt).reshape(-1)


def fir_overlap_save_reference(stream: np.ndarray, taps: np.ndarray) -> np.ndarray:
    """OUR extension (no reference counterpart): true linear convolution
    z[n] = sum_k h[k] y[n-k], y[n<0] = 0, computed directly in complex128."""
    y = np.asarray(stream, dtype=np.complex64).astype(np.complex128)
    h = np.asarray(taps).astype(np.complex128)
    import scipy.signal
    z = scipy.signal.fftconvolve(y, h)[: y.shape[0]] if y.shape[0] > 1 << 16 else np.convolve(y, h)[: y.shape[0]]
    return z.astype(np.complex64)


# --------------------------------------------------------------------------------------
# composed chain
# --------------------------------------------------------------------------------------

def chain(raw: np.ndarray, fmt: int, sample_rate: int, shift_hz: float, filt: np.ndarray,
          decim: int, ts0: float = 0.0):
    """ConvertReader -> ShiftReader -> ConvolutionReader -> DecimateReader over one raw
    buffer read from a block boundary.  Returns (decimated c64, carried ts)."""
    x = convert_to_c64(raw, fmt)
    y, ts = shift_buffer(x, shift_hz, sample_rate, ts0)
    z = convolution_reader(y, filt)
    w = decimate_reader(z, decim)
    return w, ts


# --------------------------------------------------------------------------------------
# a13: Beamform
# --------------------------------------------------------------------------------------

SPEED_OF_LIGHT = 299792458.0  # hz.tools/rf v0.0.7 Hz.Wavelength(), pinned by stream/beamform_test.go:115-155


def beamform_angles_2d(frequency_hz: float, angle_deg: float, center, antennas) -> np.ndarray | None:
    """stream/beamform.go:57-107, fp64 host math -> complex64 weights."""
    if len(antennas) == 0:
        return None
    ret = np.empty(len(antennas), dtype=np.complex64)
    wavelength = SPEED_OF_LIGHT / float(frequency_hz)
    for i, ant in enumerate(antennas):
        xd = ant[0] - center[0]
        xy = ant[1] - center[1]
        n_distance = math.sqrt(xd * xd + xy * xy)
        if n_distance == 0:
            ret[i] = 1
            continue
        angle_r = angle_deg * (math.pi / 180)
        n_opposite = ant[1] - center[1]
        n_theta_r = math.asin(n_opposite / n_distance)
        p_theta_r = n_theta_r + angle_r
        p_opposite = math.sin(p_theta_r) * n_distance
        phase_shift = (p_opposite / wavelength) * 360
        phase_shift_r = phase_shift * (math.pi / 180)
        ret[i] = np.complex64(complex(math.cos(phase_shift_r), -math.sin(phase_shift_r)))
    return ret


def beamform_angles(frequency_hz: float, angle_deg: float, distances) -> np.ndarray | None:
    """stream/beamform.go:115-128."""
    if len(distances) == 0:
        return None
    ants = [(d, 0.0) for d in distances]
    return beamform_angles_2d(frequency_hz, angle_deg, ants[0], ants)


def beamform(channels, fmt: int, weights: np.ndarray) -> np.ndarray:
    """stream/beamform.go:148-171: per channel ConvertReader(C64) -> Multiply(w_c) -> Add,
    summed left to right in fp32 starting from 0 (stream/add.go:165-168)."""
    terms = [rotate(convert_to_c64(ch, fmt), w) for ch, w in zip(channels, weights)]
    return add(*terms)


# --------------------------------------------------------------------------------------
# synthetic inputs; testutils/cw.go:31-44 for the CW
# --------------------------------------------------------------------------------------

def cw(n: int, freq: float, sample_rate: int, phase: float = 0.0) -> np.ndarray:
    """testutils/cw.go:31-44."""
    now = np.arange(n, dtype=np.float64) / float(sample_rate)
    a = TAU * float(freq) * now + phase
    return _to_c64(np.cos(a).astype(np.float32), np.sin(a).astype(np.float32))


# synthetic inputs live outside the oracle (go-sdr_b200/python/hzsdr_synth.py) so that the product's
# benchmark arm never imports this module; re-exported here for the tests' convenience
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "..", "go-sdr_b200", "python"))
from hzsdr_synth import filter_freq, lowpass_taps, synth_raw  # noqa: E402,F401


def rel_l2(a: np.ndarray, b: np.ndarray) -> float:
    a = np.asarray(a).astype(np.complex128).reshape(-1)
    b = np.asarray(b).astype(np.complex128).reshape(-1)
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / den) if den > 0 else float(np.linalg.norm(a - b))
