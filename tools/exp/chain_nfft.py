#!/usr/bin/env python
"""Device-resident throughput of the fused chain (i8, 2^22-sample buffers, decimate x D) as a function of
the filter length N: which lengths have a specialised kernel and what the generic one costs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "go-sdr_b200", "python"))
import torch
import hzsdr as H
import hzsdr_synth as S

ctx = H.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", 0))
n, nbuf = 1 << 22, 16
host = S.synth_raw(4, n, 20_000_000, 2.5e6, 1)
src = [ctx.to_device(host) for _ in range(nbuf)]
for nfft, D in ((256, 10), (512, 10), (1024, 10), (1024, 5), (2048, 10), (4096, 16), (8192, 16), (16384, 16)):
    filt = S.filter_freq(S.lowpass_taps(nfft // 4 - 1, 0.05), nfft)
    chain = H.Chain(ctx, 4, 20_000_000, -2.5e6, filt, D)
    per = chain.out_len(n)
    outs = [ctx.alloc(per * 8) for _ in range(nbuf)]
    def step():
        for i in range(nbuf):
            chain.exec(src[i].ptr, n, outs[i].ptr, per)
    for _ in range(2): step()
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(5): step()
    e1.record(stream); ctx.sync(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"N={nfft:6d} D={D:3d}: {nbuf * n / ms / 1e6:8.1f} Gsamples/s", flush=True)
    del outs, chain
