"""GPU parity of the fused polyphase decimator (hzsdr_polyphase_*: raw -> Convert -> Shift -> real-tap FIR ->
keep every D-th sample in one kernel; BASELINE north_star's "fused polyphase FIR+decimate").

The reference has no FIR (its ConvolutionReader is block-circular, stream/convolution.go:57-81), so the
definition is SURVEY.md 2.3b's: the TRUE linear convolution z[n] = sum_k h[k] y[n-k] of the mixed stream
(y[n<0] = 0), out[i] = z[D i] with a continuous decimation phase -- checked against a direct complex128 FIR of
the oracle's Convert + Shift output.  Bar: relative L2 <= 1e-5, carried ts bit-equal to the oracle's."""
import numpy as np
import pytest

import go_sdr_oracle as O
import hzsdr as H
from gpu_impl import GpuImpl

pytestmark = pytest.mark.gpu

TOL = 1e-5


@pytest.fixture(scope="module")
def gpu():
    return GpuImpl()


def oracle_poly(raw, fmt, fs, shift, taps, D, ts0=0.0):
    x = O.convert_to_c64(raw, fmt)
    y, ts = O.shift_buffer(x, shift, fs, ts0)
    z = O.fir_overlap_save_reference(y, taps)
    return z[::D], ts


CASES = [
    # fmt, fs, n, f0, ntaps, D
    (H.FORMAT_I8, 20_000_000, 1 << 18, 2.5e6, 255, 10),      # C2's filter and decimation
    (H.FORMAT_I16, 61_440_000, 1 << 17, 1.0e6, 255, 16),     # C5's: ntaps / D = 16, a whole number of tap pairs
    (H.FORMAT_U8, 2_400_000, 100_003, 300e3, 127, 10),       # u8 (silence is not code 0), ragged length
    (H.FORMAT_I16, 8_000_000, 1 << 16, 1e6, 31, 1),          # no decimation
    (H.FORMAT_I8, 20_000_000, 1 << 16, 2.5e6, 64, 3),        # odd D, taps a multiple of nothing
    (H.FORMAT_I16, 61_440_000, 1 << 16, 7.68e6, 1023, 48),   # D > 32: more phases than lanes
    (H.FORMAT_I8, 20_000_000, 40_000, 2.5e6, 1, 7),          # a single tap: plain decimation of the mixed stream
    (H.FORMAT_I16, 20_000_000, 1 << 16, 2.5e6, 2047, 64),    # long filter: fewer warps per CTA
]


@pytest.mark.parametrize("fmt,fs,n,f0,ntaps,D", CASES)
def test_polyphase_parity(gpu, fmt, fs, n, f0, ntaps, D):
    raw = O.synth_raw(fmt, n, fs, f0, seed=ntaps + D)
    taps = O.lowpass_taps(ntaps, 1 / (2 * max(D, 2))).astype(np.float32) if ntaps > 1 else np.array([0.75], dtype=np.float32)
    want, ts_want = oracle_poly(raw, fmt, fs, -f0, taps, D)
    pp = H.Polyphase(gpu.ctx, fmt, fs, -f0, taps, D)
    total = pp.out_len(n)
    assert total == want.size
    src, dst = gpu.ctx.to_device(raw), gpu.ctx.alloc(max(total, 1) * 8)
    assert pp.exec(src.ptr, n, dst.ptr, total) == total
    got = dst.download(np.complex64, total)
    assert pp.ts == ts_want
    err = O.rel_l2(got, want)
    assert err <= TOL, err
    pp.close()


@pytest.mark.parametrize("ts0", [0.0, 3.9999, 6.2831, 5.0])
def test_polyphase_history_carried_across_calls(gpu, ts0):
    """Ragged consecutive calls through ONE decimator equal one long buffer: the carried raw history, its accumulator
    segments and the decimation phase join the calls seamlessly -- at stream start (calls shorter than the history),
    across a binade edge and across the 2*pi-second wrap."""
    fmt, fs, f0, ntaps, D = H.FORMAT_I8, 20_000_000, 2.5e6, 255, 10
    parts = [100, 57, 1000, 65536, 3, 70001, 32768, 12345]
    n = sum(parts)
    raw = O.synth_raw(fmt, n, fs, f0, seed=78)
    taps = O.lowpass_taps(ntaps, 1 / (2 * D)).astype(np.float32)
    want, ts_want = oracle_poly(raw, fmt, fs, -f0, taps, D, ts0=ts0)
    pp = H.Polyphase(gpu.ctx, fmt, fs, -f0, taps, D)
    pp.ts = ts0
    src = gpu.ctx.to_device(raw)
    dst = gpu.ctx.alloc(want.size * 8)
    done_in = done_out = 0
    for m in parts:
        cnt = pp.out_len(m)
        got = pp.exec(src.ptr + 2 * done_in, m, dst.ptr + 8 * done_out, want.size - done_out)
        assert got == cnt
        done_in += m
        done_out += cnt
    assert done_out == want.size
    assert pp.ts == ts_want
    got = dst.download(np.complex64, want.size)
    err = O.rel_l2(got, want)
    assert err <= TOL, err
    # and piecewise: no call may be much worse than the whole
    edges = np.cumsum([0] + [len(range((-s) % D, m, D)) for s, m in zip(np.cumsum([0] + parts[:-1]), parts)])
    for a, b in zip(edges[:-1], edges[1:]):
        if b - a >= 8:
            assert O.rel_l2(got[a:b], want[a:b]) <= 5 * TOL, (a, b)
    pp.close()


def test_polyphase_pluto_lsb(gpu):
    fmt, fs, f0, ntaps, D, n = H.FORMAT_I16, 61_440_000, 7.68e6, 255, 16, 1 << 16
    raw12 = (O.synth_raw(fmt, n, fs, f0, seed=5).astype(np.int32) >> 4).astype(np.int16)
    taps = O.lowpass_taps(ntaps, 1 / (2 * D)).astype(np.float32)
    want, ts_want = oracle_poly(O.shift_lsb_to_msb_bits(raw12, 12), fmt, fs, -f0, taps, D)
    pp = H.Polyphase(gpu.ctx, fmt, fs, -f0, taps, D, i16_lsb_bits=12)
    total = pp.out_len(n)
    src, dst = gpu.ctx.to_device(raw12), gpu.ctx.alloc(total * 8)
    assert pp.exec(src.ptr, n, dst.ptr, total) == total
    assert pp.ts == ts_want
    assert O.rel_l2(dst.download(np.complex64, total), want) <= TOL
    pp.close()


def test_polyphase_rejects_what_it_cannot_hold(gpu):
    with pytest.raises(H.HzsdrError):
        H.Polyphase(gpu.ctx, H.FORMAT_C64, 1_000_000, 0.0, np.ones(8, dtype=np.float32), 2)  # raw formats only
    with pytest.raises(H.HzsdrError):
        H.Polyphase(gpu.ctx, H.FORMAT_I8, 1_000_000, 0.0, np.ones(65536, dtype=np.float32), 2)  # 32768 taps per phase: no room
