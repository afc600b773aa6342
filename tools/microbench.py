#!/usr/bin/env python
"""Per-kernel HBM roofline of the elementwise kernels (K1, K2, K3/K4/K5, K7, K8): algorithmic bytes /
CUDA-event time / measured copy peak.  Run on a B200: python tools/microbench.py > profiles/...json"""
import json
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R + "/go-sdr_b200/python"]
import numpy as np
import torch

import hzsdr as H
import hzsdr_synth as O  # input generators only; nothing under oracle/ is used here

ctx = H.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream)
peak = json.load(open(R + "/MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists(R + "/MEASURED_PEAKS.json") else 6650.0
N = 1 << 26  # samples per launch: >= 128 MiB touched, larger than L2


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    ctx.sync()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps / 1e3


out = []


def report(name, bytes_per_sample, secs, n=N):
    gbs = bytes_per_sample * n / secs / 1e9
    out.append({"kernel": name, "samples": n, "us": secs * 1e6, "Msamples_s": n / secs / 1e6, "bytes_per_sample": bytes_per_sample,
                "GBs": gbs, "frac_of_measured_peak": gbs / peak})
    print(f"{name:34s} {secs * 1e6:9.1f} us  {n / secs / 1e9:8.1f} Gsamples/s  {gbs:7.0f} GB/s  {100 * gbs / peak:5.1f}% of {peak:.0f}", file=sys.stderr)


rng = np.random.default_rng(0)
raw8 = rng.integers(0, 255, size=2 * (1 << 22), endpoint=True).astype(np.uint8)
d_u8 = ctx.alloc(2 * N)
for k in range(N // (1 << 22)):
    d_u8.upload(raw8, k * raw8.nbytes)
d_i16 = ctx.alloc(4 * N)
H._check(H.load().hzsdr_copy(ctx.h, d_i16.ptr, d_u8.ptr, 2 * N))
H._check(H.load().hzsdr_copy(ctx.h, d_i16.ptr + 2 * N, d_u8.ptr, 2 * N))
c64 = ctx.alloc(8 * N)
c64b = ctx.alloc(8 * N)

timeit(lambda: ctx.scale(c64.ptr, N, 1.0), reps=50)  # let the clocks ramp before the first measurement
report("convert u8->c64 (K1)", 10, timeit(lambda: ctx.convert_to_c64(H.FORMAT_U8, d_u8.ptr, N, c64.ptr, N)))
report("convert i8->c64 (K1)", 10, timeit(lambda: ctx.convert_to_c64(H.FORMAT_I8, d_u8.ptr, N, c64.ptr, N)))
report("convert i16->c64 (K1)", 12, timeit(lambda: ctx.convert_to_c64(H.FORMAT_I16, d_i16.ptr, N, c64.ptr, N)))
st = H.NcoState(20_000_000, 1.0)
report("shift c64 in place (K2)", 16, timeit(lambda: ctx.shift(c64.ptr, N, -2.5e6, st)))
report("convert+shift u8 fused (K1+K2)", 10, timeit(lambda: ctx.convert_shift(H.FORMAT_U8, d_u8.ptr, N, c64.ptr, N, -2.5e6, st)))
report("convert+shift i16 fused (K1+K2)", 12, timeit(lambda: ctx.convert_shift(H.FORMAT_I16, d_i16.ptr, N, c64.ptr, N, -2.5e6, st)))
report("rotate c64 in place (K3)", 16, timeit(lambda: ctx.rotate(c64.ptr, N, 0.6 + 0.8j)))
report("scale c64 in place (K4)", 16, timeit(lambda: ctx.scale(c64.ptr, N, 0.5)))
n4 = N // 4
srcs = [c64.ptr + k * n4 * 8 for k in range(4)]
report("add 4 x c64 (K5)", 8 * 4 + 8, timeit(lambda: ctx.add(c64b.ptr, srcs, n4)), n=n4)
report("decimate c64 x10, 32768 blocks (K7)", 8 * 0.1 * 2, timeit(lambda: ctx.decimate(H.FORMAT_C64, c64.ptr, N, c64b.ptr, N, 10, 32768)))
report("decimate c64 x2 (K7)", 8 + 4, timeit(lambda: ctx.decimate(H.FORMAT_C64, c64.ptr, N, c64b.ptr, N, 2, 32768)))
report("downsample c64 x4 (K7)", 8 + 2, timeit(lambda: ctx.downsample(H.FORMAT_C64, c64.ptr, N, c64b.ptr, N, 4, 32768)))
nb = 1 << 20
chans = [d_u8.ptr + c * 2 * nb for c in range(64)]
w = H.beamform_angles(433e6, 30.0, [0.15 * c for c in range(64)])
# a streaming pipeline rotates its output buffers: consecutive launches then touch disjoint memory and may
# overlap; writing the same destination every time serialises them
beam_out = [c64b.ptr + k * nb * 8 for k in range(8)]
turn = [0]


def beam():
    turn[0] += 1
    ctx.beamform(H.FORMAT_U8, chans, w, nb, beam_out[turn[0] % 8])


report("beamform 64 x u8 (K8), per out sample", 64 * 2 + 8, timeit(beam, reps=40), n=nb)
filt = ctx.to_device(O.filter_freq(O.lowpass_taps(255, 1 / 20), 1024))
report("convolve_freq N=1024 (K6b), same destination", 16, timeit(lambda: ctx.convolve_freq(c64.ptr, c64b.ptr, filt.ptr, 1024, N // 1024)))
half = N // 2


def conv_rotating():
    turn[0] += 1
    ctx.convolve_freq(c64.ptr, c64b.ptr + (turn[0] & 1) * half * 8, filt.ptr, 1024, half // 1024)


report("convolve_freq N=1024 (K6b), rotating destinations", 16, timeit(conv_rotating, reps=40), n=half)
for n in (256, 1024, 4096, 16384):
    plan = H.FftPlan(ctx, n, n, H.FFT_FORWARD)
    report(f"fft forward N={n} (K6a)", 16, timeit(lambda: plan.transform(c64.ptr, c64b.ptr, N // n)))
print(json.dumps(out, indent=1))
