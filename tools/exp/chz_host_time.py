"""Host time per hzsdr_channelizer_exec call against the GPU time of the launch it enqueues (C5 shape)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "go-sdr_b200", "python"))
import numpy as np, torch
import hzsdr as H, hzsdr_synth as O
sys.path.insert(0, ROOT)
import bench
w = bench.WORKLOADS["c5"]
ctx = H.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream)
n, ns = w["n"], w["streams"]
filt = bench.filter_for(w)
base = [ctx.to_device(O.synth_raw(w["fmt"], n, w["fs"], w["f0"] + 10e3 * i, seed=i)) for i in range(4)]
srcs = []
for i in range(ns):
    d = ctx.alloc(n * w["raw"]); H._check(H.load().hzsdr_copy(ctx.h, d.ptr, base[i % 4].ptr, n * w["raw"])); srcs.append(d)
chz = H.Channelizer(ctx, w["fmt"], w["fs"], [-(w["f0"] + 10e3 * s) for s in range(ns)], filt, w["D"])
per = n // 32768 * (32768 // w["D"])
dsts = [ctx.alloc(per * 8) for _ in range(ns)]
sp, dp = [x.ptr for x in srcs], [x.ptr for x in dsts]
for _ in range(4): chz.exec(sp, n, dp, per)
ctx.sync()
for reps in (1, 20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); t0 = time.perf_counter()
    for _ in range(reps): chz.exec(sp, n, dp, per)
    t1 = time.perf_counter(); e1.record(stream); ctx.sync(); torch.cuda.synchronize()
    print(f"reps {reps}: host {1e3*(t1-t0)/reps:.3f} ms per call, gpu {e0.elapsed_time(e1)/reps:.3f} ms per call")
