import sys, time, os
import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0]=[R+'/go-sdr_b200/python']
import numpy as np, hzsdr as H, hzsdr_synth as O  # host enqueue cost per chain launch
ctx=H.Context(0)
n=1<<22; fs=20_000_000
raw=O.synth_raw(4,n,fs,2.5e6,seed=1)
filt=O.filter_freq(O.lowpass_taps(255,1/20),1024)
ch=H.Chain(ctx,4,fs,-2.5e6,filt,10)
per=ch.out_len(n)
srcs=[ctx.to_device(raw) for _ in range(16)]
outs=[ctx.alloc(per*8) for _ in range(16)]
for rep in range(3):
    ctx.sync(); t0=time.perf_counter()
    for k in range(20):
        for i in range(16): ch.exec(srcs[i].ptr,n,outs[i].ptr,per)
    t1=time.perf_counter(); ctx.sync(); t2=time.perf_counter()
    print("host enqueue per call %.2f us; total per call %.2f us"%((t1-t0)/320*1e6,(t2-t0)/320*1e6))
