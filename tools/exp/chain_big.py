#!/usr/bin/env python
"""Three launches of the C2-shaped fused chain on one 2^26-sample buffer (steady-state kernel for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "go-sdr_b200", "python"))
import hzsdr as H
import hzsdr_synth as S
ctx = H.Context(0)
filt = S.filter_freq(S.lowpass_taps(255, 0.05), 1024)
n = 1 << 26
chain = H.Chain(ctx, 4, 20_000_000, -2.5e6, filt, 10)
per_out = chain.out_len(n)
src = ctx.to_device(S.synth_raw(4, n, 20_000_000, 2.5e6, 1))
out = ctx.alloc(per_out * 8)
for _ in range(3):
    chain.exec(src.ptr, n, out.ptr, per_out)
ctx.sync()
print("ok")
