#!/usr/bin/env python
"""Device-resident throughput of the C2-shaped fused chain as a function of the buffer length
(launch granularity): total samples fixed at 2^28 per timed pass.  usage: python tools/exp/chain_sizes.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "go-sdr_b200", "python"))
import numpy as np
import torch
import hzsdr as H
if os.environ.get('HZSDR_LIB_OVERRIDE'): H.LIB_PATH = os.environ['HZSDR_LIB_OVERRIDE']
import hzsdr_synth as S

ctx = H.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", 0))
filt = S.filter_freq(S.lowpass_taps(255, 0.05), 1024)
TOTAL = 1 << 28
for fmt, raw in ((4, 2),):
    for lg in (22, 26):
        n = 1 << lg
        nbuf = TOTAL // n
        chain = H.Chain(ctx, fmt, 20_000_000, -2.5e6, filt, 10)
        per_out = chain.out_len(n)
        host = S.synth_raw(fmt, n, 20_000_000, 2.5e6, 1)
        src = [ctx.to_device(host) for _ in range(min(nbuf, 64))]
        outs = [ctx.alloc(per_out * 8) for _ in range(min(nbuf, 64))]
        def step():
            for i in range(nbuf):
                chain.exec(src[i % len(src)].ptr, n, outs[i % len(outs)].ptr, per_out)
        for _ in range(2): step()
        ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(5): step()
        e1.record(stream); ctx.sync(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"fmt {fmt} n=2^{lg} x {nbuf:4d} launches: {TOTAL / ms / 1e6:8.1f} Gsamples/s  ({ms * 1e3 / nbuf:8.1f} us/launch)", flush=True)
        del src, outs, chain
