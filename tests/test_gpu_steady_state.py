"""GPU parity of every fused-chain kernel at the NCO states the benchmark actually runs in.

bench.py streams hundreds of seconds of signal through one chain, so the carried accumulator `ts`
(stream/shifter.go:68) lives in the [2,4) and [4,2*pi) binades and crosses the 2*pi-SECOND wrap
(stream/shifter.go:77-79) every few dozen buffers.  The chain kernels have their own mixers (split
tables keyed on the phase step, per-segment slow paths), so the stand-alone shift tests do not cover
them.  Every case here starts a chain at a given `ts0`:

    3.9999      a binade edge (4.0) a few thousand samples into the buffer: the fp64 step changes
    6.2831      the 2*pi-second wrap a few thousand samples into the buffer (the phase jumps)
    5.0 / 2.5   mid-binade steady state: one accumulator segment for the whole buffer

and asserts the carried `ts` bit-equal to the oracle's accumulator and the samples within the
north-star bar (relative L2 <= 1e-5) of oracle.chain(..., ts0=)."""
import numpy as np
import pytest

import cpu_ref as CR
import go_sdr_oracle as O
import hzsdr as H
from gpu_impl import GpuImpl

pytestmark = pytest.mark.gpu

TOL = 1e-5
TS0 = [3.9999, 6.2831, 5.0, 2.5]


@pytest.fixture(scope="module")
def gpu():
    return GpuImpl()


# fmt, fs, n, f0, taps, nfft, D -- one row per chain kernel (and per template form of k_chain1024)
KERNELS = {
    "chain1024_split_c2": (H.FORMAT_I8, 20_000_000, 1 << 18, 2.5e6, 255, 1024, 10),    # SPLIT (even D), C2 rate
    "chain1024_split_c5": (H.FORMAT_I16, 61_440_000, 1 << 18, 1.0e6, 255, 1024, 16),   # SPLIT, C5 rate/format
    "chain1024_unsplit_d5": (H.FORMAT_I8, 20_000_000, 1 << 18, 2.5e6, 255, 1024, 5),   # odd D: unpruned, unsplit
    "chain1024_unsplit_u8": (H.FORMAT_U8, 20_000_000, 1 << 17, 2.5e6, 255, 1024, 1),
    "chain16k_c3": (H.FORMAT_I16, 61_440_000, 1 << 18, 7.68e6, 4095, 16384, 16),       # CTA-per-block, C3 rate
    "chain16k_i8_d32": (H.FORMAT_I8, 20_000_000, 1 << 18, 2.5e6, 4095, 16384, 32),
    "chaink2": (H.FORMAT_I8, 20_000_000, 1 << 18, 2.5e6, 2047, 2048, 10),
    "chaink4": (H.FORMAT_I16, 61_440_000, 1 << 18, 7.68e6, 4095, 4096, 16),
    "chaink8": (H.FORMAT_I16, 20_000_000, 1 << 18, 2.5e6, 8191, 8192, 12),
    "generic512": (H.FORMAT_I8, 20_000_000, 1 << 17, 2.5e6, 127, 512, 10),
    "generic256": (H.FORMAT_U8, 61_440_000, 1 << 17, 7.68e6, 63, 256, 8),
    "generic16k_d8": (H.FORMAT_I16, 61_440_000, 1 << 17, 7.68e6, 2047, 16384, 8),      # N = 16384, D % 16 != 0
}


@pytest.mark.parametrize("ts0", TS0)
@pytest.mark.parametrize("kernel", sorted(KERNELS))
def test_chain_kernels_at_steady_state_ts(gpu, kernel, ts0):
    fmt, fs, n, f0, taps, nfft, D = KERNELS[kernel]
    raw = O.synth_raw(fmt, n, fs, f0, seed=int(ts0 * 1000) % 89 + nfft)
    Hf = O.filter_freq(O.lowpass_taps(taps, 1 / (2 * max(D, 2))), nfft)
    want, ts_want = O.chain(raw, fmt, fs, -f0, Hf, D, ts0=ts0)
    _, ts_serial = CR.shift_ts(fs, n, ts0, want_array=False)  # the literal compiled loop
    assert ts_want == ts_serial
    got, ts = gpu.chain(raw, fmt, fs, -f0, Hf, D, ts0=ts0)
    assert got.shape == want.shape
    assert ts == ts_want, "carried ts must be bit-equal to the reference accumulator"
    err = O.rel_l2(got, want)
    assert err <= TOL, (kernel, ts0, err)
    if ts0 == 6.2831:  # the wrap really is inside this buffer
        assert ts < 1.0


@pytest.mark.parametrize("ts0", TS0)
@pytest.mark.parametrize("kernel", ["chain1024_split_c2", "chain1024_unsplit_d5", "chain16k_c3", "chaink4", "generic512"])
def test_chain_two_buffers_across_the_event(gpu, kernel, ts0):
    """The buffer that contains the binade edge / wrap, then the next one (which starts just after
    it), through ONE chain object: the second buffer reuses whatever tables the first cached."""
    fmt, fs, n, f0, taps, nfft, D = KERNELS[kernel]
    n //= 2
    raw = O.synth_raw(fmt, 2 * n, fs, f0, seed=7 + nfft)
    Hf = O.filter_freq(O.lowpass_taps(taps, 1 / (2 * max(D, 2))), nfft)
    want, ts_want = O.chain(raw, fmt, fs, -f0, Hf, D, ts0=ts0)
    ch = H.Chain(gpu.ctx, fmt, fs, -f0, Hf, D)
    ch.ts = ts0
    per = ch.out_len(n)
    src = gpu.ctx.to_device(raw)
    dst = gpu.ctx.alloc(2 * per * 8)
    sb = 2 if fmt != H.FORMAT_I16 else 4
    assert ch.exec(src.ptr, n, dst.ptr, per) == per
    assert ch.exec(src.ptr + n * sb, n, dst.ptr + per * 8, per) == per
    got = dst.download(np.complex64, 2 * per)
    assert ch.ts == ts_want
    assert O.rel_l2(got, want) <= TOL
    assert O.rel_l2(got[per:], want[per:]) <= TOL
    ch.close()


@pytest.mark.parametrize("D", [16, 5])  # even: batched SPLIT kernel with per-stream tables; odd: the unpruned batch form
def test_channelizer_at_steady_state_ts(gpu, D):
    """k_chain1024<batch>: every stream starts at its own carried time (hzsdr_channelizer_set_ts) --
    binade edges, the wrap, mid-binade -- at C5's rate and format, two consecutive buffers."""
    fmt, fs, nfft, n = H.FORMAT_I16, 61_440_000, 1024, 1 << 17
    ts0 = [3.9999, 6.2831, 5.0, 2.5, 1.99999, 6.28318, 0.7, 4.0, 3.99999999, 6.2, 1.0, 0.0]
    ns = len(ts0)
    shifts = [-(1e6 + 10e3 * s) for s in range(ns)]
    Hf = O.filter_freq(O.lowpass_taps(255, 1 / 32), nfft)
    raws = [O.synth_raw(fmt, 2 * n, fs, -shifts[s], seed=90 + s) for s in range(ns)]
    chz = H.Channelizer(gpu.ctx, fmt, fs, shifts, Hf, D)
    chz.set_ts(ts0)
    assert np.array_equal(chz.ts, np.array(ts0))
    per = n // 32768 * (32768 // D)
    outs = []
    for half in range(2):
        srcs = [gpu.ctx.to_device(r[half * 2 * n:(half + 1) * 2 * n]) for r in raws]
        dsts = [gpu.ctx.alloc(per * 8) for _ in range(ns)]
        assert chz.exec([s.ptr for s in srcs], n, [d.ptr for d in dsts], per) == per
        outs.append([d.download(np.complex64, per) for d in dsts])
    ts = chz.ts
    for s in range(ns):
        want, ts_want = O.chain(raws[s], fmt, fs, shifts[s], Hf, D, ts0=ts0[s])
        got = np.concatenate([outs[0][s], outs[1][s]])
        assert ts[s] == ts_want, (s, ts0[s])
        err = O.rel_l2(got, want)
        assert err <= TOL, (s, ts0[s], err)
    chz.close()


@pytest.mark.parametrize("workload", ["c2", "c3"])
def test_long_run_last_buffer_matches_oracle(gpu, workload):
    """What bench.py times: many consecutive full-size buffers through ONE chain (C2: 48 x 2^22
    samples = 10 s of signal, so `ts` crosses the 2*pi-second wrap and three binades; C3: 26 x 2^24 = 7.1 s).
    The carried ts is checked after every buffer against the compiled serial accumulator, the
    samples of the buffers around the wrap and of the last buffer against the oracle started from
    that accumulator value."""
    if workload == "c2":
        fmt, fs, n, f0, taps, nfft, D, nbuf = H.FORMAT_I8, 20_000_000, 1 << 22, 2.5e6, 255, 1024, 10, 48
    else:
        fmt, fs, n, f0, taps, nfft, D, nbuf = H.FORMAT_I16, 61_440_000, 1 << 24, 7.68e6, 4095, 16384, 16, 26
    Hf = O.filter_freq(O.lowpass_taps(taps, 1 / (2 * D)), nfft)
    raws = [O.synth_raw(fmt, n, fs, f0, seed=500 + i) for i in range(3)]
    srcs = [gpu.ctx.to_device(r) for r in raws]
    ch = H.Chain(gpu.ctx, fmt, fs, -f0, Hf, D)
    per = ch.out_len(n)
    dst = gpu.ctx.alloc(per * 8)
    ts_before = []
    ts_serial = 0.0
    wrapped = False
    check = []  # (buffer index, ts at its start, its output)
    for b in range(nbuf):
        ts_before.append(ch.ts)
        assert ch.exec(srcs[b % 3].ptr, n, dst.ptr, per) == per
        _, ts_serial = CR.shift_ts(fs, n, ts_serial, want_array=False)
        assert ch.ts == ts_serial, b
        crossed = ch.ts < ts_before[-1]
        wrapped |= crossed
        edge = any(ts_before[-1] < e <= ch.ts for e in (1.0, 2.0, 4.0))  # the fp64 step changes inside this buffer
        if crossed or edge or b == nbuf - 1:
            check.append((b, ts_before[-1], dst.download(np.complex64, per)))
    assert wrapped, "the run was meant to cross the 2*pi-second wrap"
    assert check
    m = 1 << 19  # oracle on the head and the tail of each checked buffer (FFT blocks are independent)
    for b, ts0, got in check:
        raw = raws[b % 3]
        head, _ = O.chain(raw[: 2 * m], fmt, fs, -f0, Hf, D, ts0=ts0)
        assert O.rel_l2(got[: head.size], head) <= TOL, (b, "head")
        _, ts_tail = CR.shift_ts(fs, n - m, ts0, want_array=False)
        tail, _ = O.chain(raw[2 * (n - m):], fmt, fs, -f0, Hf, D, ts0=ts_tail)
        assert O.rel_l2(got[-tail.size:], tail) <= TOL, (b, "tail")
        if workload == "c2":  # and the whole buffer (4 s of numpy for 2^22 samples)
            want, _ = O.chain(raw, fmt, fs, -f0, Hf, D, ts0=ts0)
            assert O.rel_l2(got, want) <= TOL, (b, "whole")
    ch.close()


# fmt, fs, f0, D, lsb_bits, ts0 -- hzsdr_chain_exec_batch as ONE launch of the batched kernel (tables in Tensor Memory)
BATCHES = {
    "c2_wrap": (H.FORMAT_I8, 20_000_000, 2.5e6, 10, 0, 6.2),        # the 2*pi-second wrap inside buffer 1
    "c2_start": (H.FORMAT_I8, 20_000_000, 2.5e6, 10, 0, 0.0),       # stream start: buffer 0 needs the long segment table
    "c5_binade": (H.FORMAT_I16, 61_440_000, 1.0e6, 16, 0, 3.99),    # binade edge at 4.0
    "u8_d4": (H.FORMAT_U8, 20_000_000, 2.5e6, 4, 0, 5.0),           # D/2 even: the padded stage-C layout
    "pluto_lsb": (H.FORMAT_I16, 61_440_000, 7.68e6, 16, 12, 2.5),
}


@pytest.mark.parametrize("case", sorted(BATCHES))
def test_chain_exec_batch_one_launch(gpu, case):
    """20 consecutive 2^20-sample buffers through hzsdr_chain_exec_batch (enough blocks for the single batched
    launch) against (a) the same buffers one hzsdr_chain_exec at a time and (b) the oracle, started from the
    serial accumulator's value, on the buffers around accumulator events and on the last one."""
    fmt, fs, f0, D, lsb, ts0 = BATCHES[case]
    n, nbuf, taps, nfft = 1 << 20, 20, 255, 1024
    Hf = O.filter_freq(O.lowpass_taps(taps, 1 / (2 * D)), nfft)
    raws = [O.synth_raw(fmt, n, fs, f0, seed=700 + i) for i in range(4)]
    if lsb:
        raws = [(r.astype(np.int32) >> (16 - lsb)).astype(np.int16) for r in raws]
    srcs = [gpu.ctx.to_device(r) for r in raws]
    ch = H.Chain(gpu.ctx, fmt, fs, -f0, Hf, D, i16_lsb_bits=lsb)
    one = H.Chain(gpu.ctx, fmt, fs, -f0, Hf, D, i16_lsb_bits=lsb)
    ch.ts = one.ts = ts0
    per = ch.out_len(n)
    outs = [gpu.ctx.alloc(per * 8) for _ in range(nbuf)]
    ref = gpu.ctx.alloc(per * 8)
    packed = H.Chain.pack_batch([srcs[b % 4].ptr for b in range(nbuf)], [o.ptr for o in outs])
    assert ch.exec_batch(packed, n, per) == per
    ts_start, ts = [], ts0
    for b in range(nbuf):
        ts_start.append(ts)
        _, ts = CR.shift_ts(fs, n, ts, want_array=False)
    assert ch.ts == ts
    worst = 0.0
    for b in range(nbuf):
        assert one.exec(srcs[b % 4].ptr, n, ref.ptr, per) == per
        a, r = outs[b].download(np.complex64, per), ref.download(np.complex64, per)
        worst = max(worst, O.rel_l2(a, r))
    assert one.ts == ts
    assert worst <= 2e-6, worst  # two CUDA paths of the same arithmetic (the split tables are built in different places)
    events = [b for b in range(nbuf)
              if b in (0, nbuf - 1) or ts_start[b] > (ts_start[b + 1] if b + 1 < nbuf else ts)
              or any(ts_start[b] < e <= (ts_start[b + 1] if b + 1 < nbuf else ts) for e in (1.0, 2.0, 4.0))]
    for b in events:
        raw = O.shift_lsb_to_msb_bits(raws[b % 4], lsb) if lsb else raws[b % 4]
        want, _ = O.chain(raw, fmt, fs, -f0, Hf, D, ts0=ts_start[b])
        assert O.rel_l2(outs[b].download(np.complex64, per), want) <= TOL, (case, b)
    ch.close()
    one.close()


def test_chain_exec_batch_orders_conflicting_buffers(gpu):
    """Two buffers of one batch that write the SAME destination: the batch must behave like the calls one by one
    (the later buffer wins), i.e. the library splits the batch instead of letting one kernel race with itself."""
    fmt, fs, f0, D = H.FORMAT_I8, 20_000_000, 2.5e6, 10
    n, nbuf = 1 << 19, 12
    Hf = O.filter_freq(O.lowpass_taps(255, 1 / (2 * D)), 1024)
    raws = [O.synth_raw(fmt, n, fs, f0, seed=900 + i) for i in range(nbuf)]
    srcs = [gpu.ctx.to_device(r) for r in raws]
    ch = H.Chain(gpu.ctx, fmt, fs, -f0, Hf, D)
    one = H.Chain(gpu.ctx, fmt, fs, -f0, Hf, D)
    ch.ts = one.ts = 2.5
    per = ch.out_len(n)
    outs = [gpu.ctx.alloc(per * 8) for _ in range(nbuf)]
    refs = [gpu.ctx.alloc(per * 8) for _ in range(nbuf)]
    alias = {7: 3, 9: 3}  # buffers 3, 7 and 9 share a destination
    dst = [outs[alias.get(b, b)].ptr for b in range(nbuf)]
    rdst = [refs[alias.get(b, b)].ptr for b in range(nbuf)]
    for _ in range(3):  # back to back: the second and third call also conflict with what is still in flight
        assert ch.exec_batch(H.Chain.pack_batch([s.ptr for s in srcs], dst), n, per) == per
        for b in range(nbuf):
            assert one.exec(srcs[b].ptr, n, rdst[b], per) == per
            gpu.ctx.sync()
    assert ch.ts == one.ts
    for b in range(nbuf):
        if b in alias:
            continue
        a, r = outs[b].download(np.complex64, per), refs[b].download(np.complex64, per)
        assert O.rel_l2(a, r) <= 2e-6, b
    ch.close()
    one.close()


def test_ring_to_chain_end_to_end(gpu):
    """The driver hand-off: a PRODUCER THREAD writes raw buffers into pinned ring slots
    (UnsafeRingBuffer.WritePeekUnsafePointer / WritePoke, stream/ring.go:344-392), the owner thread drains them with
    hzsdr_chain_submit_ring -- async H2D on the ring's copy stream, fused kernel, async D2H -- with no host wait in
    between.  24 buffers through a 4-slot ring (the producer laps the ring six times and blocks on slots whose copy is
    still pending); every output equals the oracle's for the stream position it was read at, ts bit-equal."""
    import ctypes as C
    import threading
    import time

    fmt, fs, f0, D = H.FORMAT_I8, 20_000_000, 2.5e6, 10
    n, nbuf, slots = 1 << 18, 24, 4
    Hf = O.filter_freq(O.lowpass_taps(255, 1 / (2 * D)), 1024)
    raws = [O.synth_raw(fmt, n, fs, f0, seed=40 + (i % 5)) for i in range(nbuf)]
    ring = H.Ring(gpu.ctx, fmt, slots, n)
    ch = H.Chain(gpu.ctx, fmt, fs, -f0, Hf, D)
    ch.ts = 6.2  # the 2*pi-second wrap falls inside the run
    per = ch.out_len(n)
    outs = [H.PinnedBuffer(per * 8) for _ in range(nbuf)]
    errors = []
    free_slots = threading.Semaphore(slots)  # the driver's pacing: an overrun would drop the oldest slot (ring.go:170-186)

    def producer():
        try:
            for i in range(nbuf):
                free_slots.acquire()
                p = ring.write_peek()
                C.memmove(p, raws[i].ctypes.data, raws[i].nbytes)
                ring.write_poke(n)
                if i % 7 == 3:
                    time.sleep(0.002)  # let the reader run dry now and then
        except Exception as e:  # pragma: no cover
            errors.append(e)

    t = threading.Thread(target=producer)
    t.start()
    done, underruns = 0, 0
    deadline = time.time() + 60
    while done < nbuf and time.time() < deadline:
        try:
            assert ch.submit_ring(ring, outs[done].ptr, per) == per
            done += 1
            free_slots.release()
        except H.HzsdrError as e:
            assert e.status == H.ERR_RING_UNDERRUN
            underruns += 1
            time.sleep(0.0002)
    t.join()
    ch.wait_host()
    assert not errors and done == nbuf
    ts = 6.2
    for i in range(nbuf):
        ts0 = ts
        _, ts = CR.shift_ts(fs, n, ts, want_array=False)
        if i in (0, 1, 2, 3, 7, nbuf - 1) or ts < ts0:
            want, _ = O.chain(raws[i], fmt, fs, -f0, Hf, D, ts0=ts0)
            got = outs[i].view(np.complex64)[:per]
            assert O.rel_l2(got, want) <= TOL, i
    assert ch.ts == ts
    ring.close()
    ch.close()


@pytest.mark.parametrize("fmt,ts0", [(H.FORMAT_U8, 0.0), (H.FORMAT_U8, 6.2829), (H.FORMAT_I16, 3.9999), (H.FORMAT_I8, 5.0)])
def test_convert_shift_batch(gpu, fmt, ts0):
    """hzsdr_convert_shift_batch (C1: rtl u8 buffers -> ConvertBuffer -> ShiftReader, many buffers per launch) against
    the oracle's Convert + Shift buffer by buffer: carried ts bit-equal to the serial accumulator, rel-L2 <= 1e-5;
    70 buffers = two launches, the first buffer from ts = 0 has dozens of accumulator segments."""
    fs, f0, n, nbuf = 2_400_000, 300e3, 1 << 16, 70
    raws = [O.synth_raw(fmt, n, fs, f0, seed=300 + i) for i in range(5)]
    srcs = [gpu.ctx.to_device(r) for r in raws]
    outs = [gpu.ctx.alloc(n * 8) for _ in range(nbuf)]
    st = H.NcoState(fs, ts0)
    packed = H.Chain.pack_batch([srcs[b % 5].ptr for b in range(nbuf)], [o.ptr for o in outs])
    gpu.ctx.convert_shift_batch(fmt, packed, n, n, -f0, st)
    ts = ts0
    for b in range(nbuf):
        want, ts = O.shift_buffer(O.convert_to_c64(raws[b % 5], fmt), -f0, fs, ts)
        if b < 4 or b % 9 == 0 or b == nbuf - 1:
            assert O.rel_l2(outs[b].download(np.complex64, n), want) <= TOL, b
    assert st.ts == ts
