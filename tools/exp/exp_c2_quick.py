import sys, os, time
sys.path[:0] = ['/root/repo/go-sdr_b200/python', '/root/repo']
import numpy as np, torch
import hzsdr as H, bench
w = bench.WORKLOADS['c2']
ctx = H.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream)
filt = bench.filter_for(w)
n, nbuf = w['n'], 64
bufs = bench.make_buffers(w, nbuf, 2, distinct=4)
src = [ctx.to_device(b) for b in bufs[:4]]
pool = []
for i in range(nbuf):
    if i < 4: pool.append(src[i]); continue
    d = ctx.alloc(n * 2); H._check(H.load().hzsdr_copy(ctx.h, d.ptr, src[i % 4].ptr, n * 2)); pool.append(d)
ch = H.Chain(ctx, w['fmt'], w['fs'], -w['f0'], filt, w['D'])
per = ch.out_len(n)
outs = [ctx.alloc(per * 8) for _ in range(nbuf)]
K = int(os.environ.get('K', '64'))  # buffers per hzsdr_chain_exec_batch call
packs = [H.Chain.pack_batch([p.ptr for p in pool[i:i + K]], [o.ptr for o in outs[i:i + K]]) for i in range(0, nbuf, K)]
def step():
    for pk in packs: ch.exec_batch(pk, n, per)
for _ in range(5): step()
ctx.sync()
best = 0
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(40): step()
    e1.record(stream); ctx.sync(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    best = max(best, 40 * nbuf * n / ms / 1e6)
print(os.environ.get('HZSDR_LIB', 'default'), 'K', K, 'Gsamples/s', round(best, 1))
