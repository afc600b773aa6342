#!/usr/bin/env python
"""A few launches of one workload's dominant kernel, for ncu (one GPU, short):
    ncu --set full --clock-control none --import-source on -k regex:<kernel> -s <skip> -c 1 -o gpurun_out/x \\
        python tools/prof_run.py c1|c2|c2b|c3|c3os|c5|c4|c4rs|poly [launches]
No oracle, no timing claims: numbers printed under a profiler are never bench values."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "go-sdr_b200", "python"), ROOT]

import numpy as np  # noqa: E402

import bench  # noqa: E402
import hzsdr as H  # noqa: E402
import hzsdr_synth as Y  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    launches = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    ctx = H.Context(0)
    if name in ("c2", "c3", "c3os"):
        w = bench.WORKLOADS[name]
        n = w["n"]
        filt = bench.filter_for(w)
        ch = H.Chain(ctx, w["fmt"], w["fs"], -w["f0"], filt, w["D"], overlap_save_taps=w["taps"] if w.get("overlap_save") else 0)
        per = ch.out_len(n)
        raws = [ctx.to_device(Y.synth_raw(w["fmt"], n, w["fs"], w["f0"], seed=i)) for i in range(2)]
        outs = [ctx.alloc(per * 8) for _ in range(2)]
        for i in range(launches):
            ch.exec(raws[i & 1].ptr, n, outs[i & 1].ptr, per)
    elif name == "c2b":  # hzsdr_chain_exec_batch: 64 consecutive C2 buffers = one launch of the batched kernel (bench.py's step)
        w = bench.WORKLOADS["c2"]
        n, nbuf = w["n"], 64
        filt = bench.filter_for(w)
        ch = H.Chain(ctx, w["fmt"], w["fs"], -w["f0"], filt, w["D"])
        per = ch.out_len(n)
        base = [ctx.to_device(Y.synth_raw(w["fmt"], n, w["fs"], w["f0"], seed=i)) for i in range(2)]
        pool = []
        for i in range(nbuf):
            d = ctx.alloc(n * 2)
            H._check(H.load().hzsdr_copy(ctx.h, d.ptr, base[i & 1].ptr, n * 2))
            pool.append(d)
        outs = [ctx.alloc(per * 8) for _ in range(nbuf)]
        packed = H.Chain.pack_batch([p.ptr for p in pool], [o.ptr for o in outs])
        for _ in range(max(3, launches // 4)):
            ch.exec_batch(packed, n, per)
    elif name == "c1":  # hzsdr_convert_shift_batch: 64 rtl u8 buffers of 2^20 samples = one launch (bench.py's C1 step is four)
        w = bench.WORKLOADS["c1"]
        n, nbuf = w["n"], 64
        base = [ctx.to_device(Y.synth_raw(w["fmt"], n, w["fs"], w["f0"], seed=i)) for i in range(2)]
        srcs = []
        for i in range(nbuf):
            d = ctx.alloc(n * 2)
            H._check(H.load().hzsdr_copy(ctx.h, d.ptr, base[i & 1].ptr, n * 2))
            srcs.append(d)
        dsts = [ctx.alloc(n * 8) for _ in range(nbuf)]
        st = H.NcoState(w["fs"], 0.0)
        packed = H.Chain.pack_batch([s_.ptr for s_ in srcs], [d.ptr for d in dsts])
        for _ in range(max(3, launches // 4)):
            ctx.convert_shift_batch(w["fmt"], packed, n, n, -w["f0"], st)
    elif name == "poly":  # the fused polyphase decimator on C2's filter and decimation, 2^24 samples per call
        w = bench.WORKLOADS["c2"]
        n = 1 << 24
        taps = np.hamming(255).astype(np.float32) / 255
        pp = H.Polyphase(ctx, w["fmt"], w["fs"], -w["f0"], taps, w["D"])
        src = ctx.to_device(Y.synth_raw(w["fmt"], n, w["fs"], w["f0"], seed=1))
        per = n // w["D"] + 1
        out = ctx.alloc(per * 8)
        for _ in range(max(3, launches // 4)):
            pp.exec(src.ptr, n, out.ptr, per)
    elif name == "c5":
        w = bench.WORKLOADS["c5"]
        n, ns = w["n"], 512
        filt = bench.filter_for(w)
        base = [ctx.to_device(Y.synth_raw(w["fmt"], n, w["fs"], w["f0"], seed=i)) for i in range(2)]
        srcs = []
        for i in range(ns):
            d = ctx.alloc(n * 4)
            H._check(H.load().hzsdr_copy(ctx.h, d.ptr, base[i & 1].ptr, n * 4))
            srcs.append(d)
        per = n // 16
        dsts = [ctx.alloc(per * 8) for _ in range(ns)]
        chz = H.Channelizer(ctx, w["fmt"], w["fs"], [-(w["f0"] + 1e4 * s) for s in range(ns)], filt, w["D"])
        for _ in range(max(3, launches // 4)):
            chz.exec([s.ptr for s in srcs], n, [d.ptr for d in dsts], per)
    elif name in ("c4", "c4rs"):
        w = bench.WORKLOADS["c4"]
        n, nchan, nbuf = w["n"], w["channels"], 8
        weights = H.beamform_angles(433e6, 30.0, [0.15 * c for c in range(nchan)])
        base = [ctx.to_device(Y.synth_raw(w["fmt"], n, w["fs"], w["f0"], seed=c)) for c in range(4)]
        chans = []
        for c in range(nchan):
            d = ctx.alloc(n * 2)
            H._check(H.load().hzsdr_copy(ctx.h, d.ptr, base[c & 3].ptr, n * 2))
            chans.append(d)
        if name == "c4":
            out = ctx.alloc(n * 8)
            for _ in range(launches):
                ctx.beamform(w["fmt"], [c.ptr for c in chans], weights, n, out.ptr)
        else:  # the fused reduce-scatter kernel with a single rank: the peer stores land locally (8 of the 64 channels)
            grp = H.BeamGroup(ctx, 1, 0, n, max_batch=nbuf)
            slices = [ctx.alloc(n * 8) for _ in range(nbuf)]
            packed = H.BeamGroup.pack_batch([[c.ptr for c in chans[:8]] for _ in range(nbuf)], weights[:8], [s.ptr for s in slices])
            for _ in range(max(3, launches // 2)):
                grp.exec_batch_packed(w["fmt"], packed)
            grp.join()
            ctx.sync()
            grp.close()
    ctx.sync()
    print("done", name)


if __name__ == "__main__":
    main()
