"""Throughput of the fused polyphase decimator (device-resident, CUDA events): C2's filter (255 taps, /10), C5's (/16),
a short filter, against the FFT chain's numbers from bench.py.  No oracle, no parity claims here (tests/ has those)."""
import os, sys
sys.path[:0] = ['/root/repo/go-sdr_b200/python', '/root/repo']
import numpy as np, torch
import hzsdr as H, hzsdr_synth as Y, bench
ctx = H.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream)
for name, fmt, fs, f0, ntaps, D, n in [("c2 i8 255/10", H.FORMAT_I8, 20_000_000, 2.5e6, 255, 10, 1 << 24),
                                        ("c5 i16 255/16", H.FORMAT_I16, 61_440_000, 1e6, 255, 16, 1 << 24),
                                        ("i8 127/10", H.FORMAT_I8, 20_000_000, 2.5e6, 127, 10, 1 << 24),
                                        ("i8 63/8", H.FORMAT_I8, 20_000_000, 2.5e6, 63, 8, 1 << 24),
                                        ("c2 per 2^22 buffer", H.FORMAT_I8, 20_000_000, 2.5e6, 255, 10, 1 << 22)]:
    taps = bench.lowpass(ntaps, 1 / (2 * D)) if hasattr(bench, 'lowpass') else np.hamming(ntaps).astype(np.float32) / ntaps
    pp = H.Polyphase(ctx, fmt, fs, -f0, np.asarray(taps, dtype=np.float32), D)
    nbuf = 8 if n >= 1 << 24 else 32
    sb = 2 if fmt != H.FORMAT_I16 else 4
    srcs = [ctx.to_device(Y.synth_raw(fmt, n, fs, f0, seed=i)) for i in range(2)]
    pool = []
    for i in range(nbuf):
        d = ctx.alloc(n * sb); H._check(H.load().hzsdr_copy(ctx.h, d.ptr, srcs[i & 1].ptr, n * sb)); pool.append(d)
    per = n // D + 1
    outs = [ctx.alloc(per * 8) for _ in range(nbuf)]
    for i in range(nbuf): pp.exec(pool[i].ptr, n, outs[i].ptr, per)
    ctx.sync()
    best = 0
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(4):
            for i in range(nbuf): pp.exec(pool[i].ptr, n, outs[i].ptr, per)
        e1.record(stream); ctx.sync(); torch.cuda.synchronize()
        best = max(best, 4 * nbuf * n / e0.elapsed_time(e1) / 1e6)
    print(f"{name:22s} {best:8.1f} Gsamples/s")
    pp.close()
