#!/usr/bin/env python
"""bench.py -- Msamples/s through Convert -> Shift -> FFT-convolution -> Decimate on B200.

    python bench.py --gpus N --steps K --warmup W            # our arm (libhzsdrcuda.so)
    python bench.py --impl reference --gpus N --steps K ...   # CPU arm: the reference's algorithm
                                                              #   on the box's host cores

Workload (BASELINE.json configs[1], "C2"): HackRF-format i8 IQ at 20 Msps in 2^22-sample buffers
-> sdr.ConvertBuffer(C64) -> stream.ShiftReader(-2.5 MHz) -> stream.ConvolutionReader(255-tap
lowpass as a 1024-bin frequency-domain filter; block-circular, the reference's semantics) ->
stream.DecimateReader(x10).  One *step* = one pass of that chain over a batch of `--buffers`
consecutive 2^22-sample buffers of one stream (default 64 = 2^28 samples, 512 MiB of raw input:
larger than the 126 MB L2, so every step streams from HBM).  With N GPUs every rank runs its own
independent stream (weak scaling, no data-path collective).

The JSON line carries the device-resident number (`value`), the end-to-end number through the
pipelined host API with pinned H2D/D2H inside the timed region (`e2e`), the roofline of the fused
kernel (`roofline`, algorithmic bytes / CUDA-event time / measured HBM peak) and the CPU baseline
(`cpu_baseline`).  Other workloads (--workload c1|c3|c4|c5|convert|...) print the same shape of
line for the secondary configs; they are reported in profiles/ and DESIGN.md, not the headline.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(ROOT, "go-sdr_b200", "python"), os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

METRIC = "Msamples/s through Convert->Shift->FFT-FIR->Decimate"
UNIT = "Msamples/s"

# ---- workload definitions -----------------------------------------------------
WORKLOADS = {
    # name: fmt, fs, buffer samples, f0, taps, nfft, decimate, raw bytes/sample
    "c2": dict(fmt=4, fs=20_000_000, n=1 << 22, f0=2.5e6, taps=255, nfft=1024, D=10, raw=2,
               desc="i8 20 Msps, 2^22-sample buffers -> Convert -> Shift(-2.5 MHz) -> 255-tap FFT convolution "
                    "(N=1024, block-circular = reference semantics) -> Decimate x10"),
    "c3": dict(fmt=3, fs=61_440_000, n=1 << 24, f0=7.68e6, taps=4095, nfft=16384, D=16, raw=4,
               desc="i16 61.44 Msps, 2^24-sample buffers -> Convert -> Shift(-7.68 MHz) -> 4095-tap FFT convolution "
                    "(N=16384, block-circular) -> Decimate x16"),
    # secondary workloads (not the headline): reported in DESIGN.md / profiles
    "c1": dict(kind="convert_shift", fmt=2, fs=2_400_000, n=1 << 20, f0=300e3, raw=2, buffers=256,
               desc="rtl u8 2.4 Msps, 2^20-sample buffers -> fused Convert + Shift(-300 kHz) (hzsdr_convert_shift)"),
    "c5": dict(kind="channelizer", fmt=3, fs=61_440_000, n=1 << 20, f0=1e6, taps=255, nfft=1024, D=16, raw=4, streams=512, buffers=1,
               desc="channelizer fan-out: 512 independent i16 streams x 2^20 samples, each Convert -> Shift(own f) -> 255-tap FFT "
                    "convolution (N=1024) -> Decimate x16; streams sharded across the GPUs, no collective; ONE launch per step"),
    "c4": dict(kind="beamform", fmt=2, fs=2_400_000, n=1 << 20, f0=100e3, raw=2, channels=64, buffers=8,
               desc="64 coherent u8 channels x 2^20 samples -> Convert -> steering Multiply -> Beamform sum; channels "
                    "sharded across the GPUs, ONE NCCL reduce of the partial beams onto rank 0"),
}


def peaks() -> dict:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return {"hbm_gbs": float(p["hbm_gbs"]), "source": "MEASURED_PEAKS.json (measured)"}
    return {"hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def ncu_traffic(workload: str):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as fh:
            return json.load(fh).get(workload)
    except Exception:
        return None


# ---- clocks sampler -----------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons during the timed region (NVML; nvidia-smi fallback)."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting",
               0x10: "sync_boost"}

    def __init__(self, device: int):
        self.device = device
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(device)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nvml = None

    def _loop(self):
        while not self._stop.is_set():
            try:
                if self._nvml:
                    n = self._nvml
                    self.samples.append(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM))
                    mask = n.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                    for bit, name in self.REASONS.items():
                        if mask & bit:
                            self.reasons.add(name)
                else:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.device), "--query-gpu=clocks.sm,clocks.max.sm,"
                                          "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                                          "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                    f = [x.strip() for x in out.strip().split(",")]
                    self.samples.append(int(f[0]))
                    self.max_mhz = int(f[1])
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:]):
                        if v.lower().startswith("active"):
                            self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.005 if self._nvml else 0.05)

    def __enter__(self):
        self._thr = threading.Thread(target=self._loop, daemon=True)
        self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._thr.join(timeout=2)

    def summary(self) -> dict:
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---- synthetic data -----------------------------------------------------------------------------
def make_buffers(w: dict, nbuf: int, seed: int, distinct: int = 8):
    """`nbuf` consecutive raw buffers of one CW+noise stream.  Noise is drawn for `distinct`
    buffers and reused round-robin (the values do not affect timing; generation time does)."""
    import hzsdr_synth as Y
    base = [Y.synth_raw(w["fmt"], w["n"], w["fs"], w["f0"], seed=seed * 1000 + i) for i in range(min(distinct, nbuf))]
    return [base[i % len(base)] for i in range(nbuf)]


def filter_for(w: dict):
    import hzsdr_synth as Y
    return Y.filter_freq(Y.lowpass_taps(w["taps"], 1.0 / (2 * w["D"])), w["nfft"])


# ---- CPU arm: the reference's algorithm on host cores -------------------------------------------
class CpuChain:
    """One stream of the reference chain on one host thread: the oracle's C twin for
    Convert/Shift/Decimate (oracle/cpu_ref.c, the reference's loop structure and SSE width) and
    scipy's pocketfft (complex64) as the user-supplied fft.Planner -- 3x faster than the C twin's
    plain radix-2 stand-in, so the faster (fairer) of the two is what gets timed."""

    def __init__(self, w: dict, filt: np.ndarray):
        import scipy.fft as sf

        import cpu_ref as CR
        self.w, self.filt, self.sf, self.CR = w, filt, sf, CR
        self.ts = 0.0

    def run(self, raw: np.ndarray) -> int:
        w, CR, sf = self.w, self.CR, self.sf
        x = CR.convert_to_c64(raw, w["fmt"])
        y, self.ts = CR.shift_buffer(x, -w["f0"], w["fs"], self.ts)
        nblk = y.size // w["nfft"]
        F = sf.fft(y[: nblk * w["nfft"]].reshape(nblk, w["nfft"]), axis=-1)
        F *= self.filt
        z = sf.ifft(F, axis=-1, norm="forward").reshape(-1)
        lz = (z.size // 32768) * 32768
        out = z[:lz].reshape(-1, 32768)[:, : (32768 // w["D"]) * w["D"] : w["D"]]
        return int(np.ascontiguousarray(out).size)


def cpu_stage_rates(w: dict, raw: np.ndarray, reps: int = 3) -> dict:
    """the reference's stages one by one on ONE host thread (Msamples/s of stage input,
    best of `reps`), and what one stream gets with a thread per stage as the reference's
    ReadTransformers run them (the slowest stage sets the pace)."""
    import scipy.fft as sf

    import cpu_ref as CR
    filt = filter_for(w)
    n = raw.shape[0]

    def best(fn):
        t = []
        for _ in range(reps):
            t0 = time.perf_counter()
            out = fn()
            t.append(time.perf_counter() - t0)
        return out, n / min(t) / 1e6

    x, r_conv = best(lambda: CR.convert_to_c64(raw, w["fmt"]))
    (y, _), r_shift = best(lambda: CR.shift_buffer(x, -w["f0"], w["fs"], 0.0))
    nblk = y.size // w["nfft"]

    def convolve():
        F = sf.fft(y[: nblk * w["nfft"]].reshape(nblk, w["nfft"]), axis=-1)
        F *= filt
        return sf.ifft(F, axis=-1, norm="forward").reshape(-1)
    z, r_fir = best(convolve)
    lz = (z.size // 32768) * 32768
    _, r_dec = best(lambda: np.ascontiguousarray(z[:lz].reshape(-1, 32768)[:, : (32768 // w["D"]) * w["D"] : w["D"]]))
    rates = {"convert": r_conv, "shift": r_shift, "convolution": r_fir, "decimate": r_dec}
    rates["one_stream_thread_per_stage"] = min(rates.values())
    rates["one_stream_one_thread"] = 1.0 / sum(1.0 / v for k, v in rates.items() if k != "one_stream_thread_per_stage")
    return rates


def cpu_throughput(w: dict, threads: int, reps: int, bufs) -> tuple[float, float]:
    """All `threads` host threads each push `reps` buffers through their own stream.  Returns
    (Msamples/s aggregate, seconds)."""
    filt = filter_for(w)
    chains = [CpuChain(w, filt) for _ in range(threads)]

    def work(i):
        for r in range(reps):
            chains[i].run(bufs[(i + r) % len(bufs)])

    with cf.ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(lambda i: chains[i].run(bufs[i % len(bufs)]), range(threads)))  # warm caches/plans
        t0 = time.perf_counter()
        list(ex.map(work, range(threads)))
        dt = time.perf_counter() - t0
    return threads * reps * w["n"] / dt / 1e6, dt


def run_reference(args, w: dict) -> dict:
    import cpu_ref as CR
    CR.build()
    threads = os.cpu_count() or 1
    bufs = make_buffers(w, min(threads, 8), seed=2, distinct=min(threads, 8))
    filt = filter_for(w)
    chains = [CpuChain(w, filt) for _ in range(threads)]
    with cf.ThreadPoolExecutor(max_workers=threads) as ex:
        step = lambda: list(ex.map(lambda i: chains[i].run(bufs[i % len(bufs)]), range(threads)))  # noqa: E731
        for _ in range(args.warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        dt = time.perf_counter() - t0
    samples_per_step = threads * w["n"]
    value = samples_per_step * args.steps / dt / 1e6
    sample = f"{threads} host threads x 1 buffer of {w['n']} samples per step (each thread its own stream)"
    return {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload + ": " + w["desc"], "buffers_per_step": threads, "l2": "inputs cycle through "
                   f"{len(bufs)} distinct buffers; CPU arm"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "note": "C restatement of the reference loops (oracle/cpu_ref.c) + scipy pocketfft as the "
                                 "Planner; the Go reference cannot be built here (no Go toolchain; GOMAXPROCS n/a)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


# ---- our arm ------------------------------------------------------------------------------------
def run_ours(args, w: dict) -> dict | None:
    import torch

    import hzsdr as H

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(local)

    ctx = H.Context(local)  # raises without a B200: there is no CPU fallback
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))
    filt = filter_for(w)
    nbuf, n = args.buffers, w["n"]
    host_bufs = make_buffers(w, nbuf, seed=2 + rank)

    chain = H.Chain(ctx, w["fmt"], w["fs"], -w["f0"], filt, w["D"])
    per_out = chain.out_len(n)
    src = [ctx.to_device(b) for b in host_bufs[: min(nbuf, 8)]]
    # device-resident pool: nbuf distinct device buffers (copies of the distinct host buffers)
    pool = []
    for i in range(nbuf):
        if i < len(src):
            pool.append(src[i])
        else:
            d = ctx.alloc(n * w["raw"])
            H._check(H.load().hzsdr_copy(ctx.h, d.ptr, src[i % len(src)].ptr, n * w["raw"]))
            pool.append(d)
    outs = [ctx.alloc(per_out * 8) for _ in range(nbuf)]
    ctx.sync()

    def step_device():
        for i in range(nbuf):
            chain.exec(pool[i].ptr, n, outs[i].ptr, per_out)

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    # ---- device-resident timing ----
    for _ in range(args.warmup):
        step_device()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        ev0.record(stream)
        for _ in range(args.steps):
            step_device()
        ev1.record(stream)
        ctx.sync()
        torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    barrier()
    launches = args.steps * nbuf
    samples_per_step = nbuf * n * world
    value = samples_per_step * args.steps / (ms / 1e3) / 1e6

    # ---- end to end: pinned host buffers, H2D + kernel + D2H per buffer, pipelined ----
    e2e_chain = H.Chain(ctx, w["fmt"], w["fs"], -w["f0"], filt, w["D"])
    pin_in = H.PinnedBuffer(nbuf * n * w["raw"])
    pin_out = H.PinnedBuffer(nbuf * per_out * 8)
    view = pin_in.view(H.NP_DTYPE[w["fmt"]])
    for i in range(nbuf):
        view[i * 2 * n:(i + 1) * 2 * n] = host_bufs[i]

    def step_e2e():
        for i in range(nbuf):
            e2e_chain.submit_host(pin_in.ptr + i * n * w["raw"], n, pin_out.ptr + i * per_out * 8, per_out)
        e2e_chain.wait_host()  # the step's result is in host memory

    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    barrier()
    e2e_steps = max(2, min(args.steps, 20))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = samples_per_step * e2e_steps / e2e_s / 1e6
    barrier()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return None

    pk = peaks()
    alg_bytes = n * w["raw"] + per_out * 8
    launch_s = (ms / 1e3) / launches
    achieved = alg_bytes / launch_s / 1e9
    log2n = w["nfft"].bit_length() - 1
    flop_per_sample = 2 * 5 * log2n + 6 + 30  # FFT pair + pointwise + convert/NCO
    fp32_tflops = flop_per_sample * (value / world) * 1e6 / 1e12

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload + ": " + w["desc"], "buffers_per_step": nbuf, "samples_per_buffer": n,
                   "parallelism": f"{world} independent stream(s), one per GPU, no collective",
                   "l2": f"each step streams {nbuf} distinct device buffers = {nbuf * n * w['raw'] >> 20} MiB of raw input "
                         "(> 126 MB L2); no explicit flush",
                   "timing": "CUDA events on the library's stream, max over ranks"},
        "clocks": clocks.summary(),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": nbuf * n * w["raw"],
                "d2h_bytes_per_step": nbuf * per_out * 8, "steps": e2e_steps,
                "api": "hzsdr_chain_submit_host/wait_host: pinned H2D -> fused kernel -> pinned D2H, 3-deep pipeline",
                "pcie_gbs": (nbuf * n * w["raw"] + nbuf * per_out * 8) * e2e_steps / e2e_s / 1e9 / world},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                     "frac": achieved / pk["hbm_gbs"], "traffic": ncu_traffic(args.workload),
                     "kernel": {1024: "hz::k_chain1024", 16384: "hz::k_chain16k"}.get(w["nfft"], f"hz::k_chain<{w['nfft']}>") + f"<fmt {w['fmt']}>", "algorithmic_bytes_per_launch": alg_bytes,
                     "bytes_per_sample": alg_bytes / n, "launch_us": launch_s * 1e6, "peak_source": pk["source"],
                     "note": "the fused chain is FP32-issue-bound, not HBM-bound (SURVEY.md 8(d)); fp32 figures alongside",
                     "fp32_tflops_est": fp32_tflops, "fp32_frac_of_74": fp32_tflops / 74.0},
    }

    if world == 1 and not args.no_cpu_baseline:
        import cpu_ref as CR
        CR.build()
        threads = os.cpu_count() or 1
        reps = args.cpu_reps
        v, dt = cpu_throughput(w, threads, reps, host_bufs[: min(8, nbuf)])
        line["cpu_baseline"] = {
            "value": v, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{threads} host threads x {reps} buffers of {n} samples each (own stream per thread), {dt:.1f} s",
            "stages_one_thread": cpu_stage_rates(w, host_bufs[0]),
            "note": "oracle/cpu_ref.c loops + scipy pocketfft as the Planner; Go reference not buildable here (no Go); "
                    "stages_one_thread: Msamples/s of each stage alone on one thread, and of one stream with a thread per "
                    "stage (the reference's goroutine-per-ReadTransformer layout) / with everything on one thread"}
    else:
        line["cpu_baseline"] = None
    if dist is not None:
        dist.destroy_process_group()
    return line


def _dist_setup():
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return torch, dist, rank, world, local


def _time_region(torch, dist, ctx, stream, local, steps, step_fn):
    """W warm-ups are the caller's; barrier + sync on both sides, CUDA events on the library's
    stream, max over ranks.  Returns (ms total, clocks summary)."""
    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        ev0.record(stream)
        for _ in range(steps):
            step_fn()
        ev1.record(stream)
        ctx.sync()
        torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    barrier()
    return ms, clocks.summary()


def run_convert_shift(args, w: dict) -> dict | None:
    """C1 on the GPU: fused u8 -> complex64 -> NCO mix, HBM-bound (2 + 8 B per sample)."""
    import hzsdr as H
    import hzsdr_synth as O
    torch, dist, rank, world, local = _dist_setup()
    ctx = H.Context(local)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))
    n, nbuf = w["n"], args.buffers
    raw = [O.synth_raw(w["fmt"], n, w["fs"], w["f0"], seed=rank * 100 + i) for i in range(4)]
    src = [ctx.to_device(raw[i % 4]) for i in range(nbuf)]
    dst = [ctx.alloc(n * 8) for _ in range(nbuf)]
    st = H.NcoState(w["fs"], 0.0)

    def step():
        for i in range(nbuf):
            ctx.convert_shift(w["fmt"], src[i].ptr, n, dst[i].ptr, n, -w["f0"], st)
    for _ in range(args.warmup):
        step()
    ms, clocks = _time_region(torch, dist, ctx, stream, local, args.steps, step)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return None
    pk = peaks()
    launches = args.steps * nbuf
    alg = n * (w["raw"] + 8)
    achieved = alg / ((ms / 1e3) / launches) / 1e9
    line = {"metric": "Msamples/s through fused Convert->Shift", "value": nbuf * n * world * args.steps / (ms / 1e3) / 1e6, "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "c1: " + w["desc"], "buffers_per_step": nbuf,
                       "l2": f"{nbuf} distinct buffer pairs = {nbuf * alg >> 20} MiB per step (> 126 MB L2)"},
            "clocks": clocks, "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"],
                         "traffic": ncu_traffic("c1"), "kernel": "hz::k_shift<U8, 4>", "algorithmic_bytes_per_launch": alg,
                         "peak_source": pk["source"]}}
    if dist is not None:
        dist.destroy_process_group()
    return line


def run_channelizer(args, w: dict) -> dict | None:
    """C5: the rank's share of 512 independent streams through hzsdr_channelizer_exec (weak in the
    sense of BASELINE: total streams fixed at 512, sharded -> strong scaling of a fixed job)."""
    import hzsdr as H
    import hzsdr_synth as O
    import hzsdr_shard as S
    torch, dist, rank, world, local = _dist_setup()
    ctx = H.Context(local)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))
    n = w["n"]
    mine = S.stream_shard(w["streams"], world, rank)
    shifts = [-(w["f0"] + 10e3 * s) for s in mine]
    filt = filter_for(w)
    base = [ctx.to_device(O.synth_raw(w["fmt"], n, w["fs"], w["f0"] + 10e3 * i, seed=i)) for i in range(4)]
    srcs = []
    for i, _ in enumerate(mine):
        d = ctx.alloc(n * w["raw"])
        H._check(H.load().hzsdr_copy(ctx.h, d.ptr, base[i % 4].ptr, n * w["raw"]))
        srcs.append(d)
    chz = H.Channelizer(ctx, w["fmt"], w["fs"], shifts, filt, w["D"])
    per = n // 32768 * (32768 // w["D"])
    dsts = [ctx.alloc(per * 8) for _ in mine]
    sp, dp = [x.ptr for x in srcs], [x.ptr for x in dsts]

    def step():
        chz.exec(sp, n, dp, per)
    for _ in range(args.warmup + 1):  # the first buffer of a stream takes the long segment tables
        step()
    ms, clocks = _time_region(torch, dist, ctx, stream, local, args.steps, step)

    # end to end: every stream's raw samples start in pinned host memory and its decimated output
    # ends there (hzsdr_channelizer_submit_host: groups of streams staged across PCIe, H2D / kernel
    # / D2H overlapped); wall clock around submit + wait, max over ranks
    e2e_steps = max(3, min(10, args.steps))
    raw_host = [O.synth_raw(w["fmt"], n, w["fs"], w["f0"] + 10e3 * i, seed=i) for i in range(4)]
    pin_in = [H.PinnedBuffer(n * w["raw"]) for _ in mine]
    pin_out = [H.PinnedBuffer(per * 8) for _ in mine]
    for i, b in enumerate(pin_in):
        b.view(np.uint8)[:] = raw_host[i % 4].view(np.uint8).reshape(-1)
    hp, op = [b.ptr for b in pin_in], [b.ptr for b in pin_out]

    def e2e_step():
        chz.submit_host(hp, n, op, per)
    e2e_step()
    ctx.wait_host()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    ctx.wait_host()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    h2d, d2h = len(mine) * n * w["raw"], len(mine) * per * 8
    e2e = {"value": w["streams"] * n * e2e_steps / e2e_s / 1e6, "unit": UNIT, "h2d_bytes_per_step": h2d * world,
           "d2h_bytes_per_step": d2h * world, "steps": e2e_steps,
           "api": "hzsdr_channelizer_submit_host + hzsdr_ctx_wait_host: pinned H2D in stream groups -> batched kernel -> pinned D2H",
           "pcie_gbs": (h2d + d2h) * e2e_steps / e2e_s / 1e9}
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return None
    pk = peaks()
    alg = len(mine) * (n * w["raw"] + per * 8)
    achieved = alg / ((ms / 1e3) / args.steps) / 1e9
    line = {"metric": METRIC + " (512-stream channelizer)", "value": w["streams"] * n * args.steps / (ms / 1e3) / 1e6, "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "c5: " + w["desc"], "streams_per_gpu": len(mine), "parallelism": "streams s mod G, no collective",
                       "l2": f"{alg >> 20} MiB touched per GPU per step (> 126 MB L2 up to 8 GPUs)"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"],
                         "traffic": ncu_traffic("c5"), "kernel": "hz::k_chain1024<I16, batch>", "algorithmic_bytes_per_launch": alg,
                         "peak_source": pk["source"], "note": "FP32-issue-bound like C2"}}
    if dist is not None:
        dist.destroy_process_group()
    return line


def run_beamform(args, w: dict) -> dict | None:
    """C4: channels sharded across ranks, one NCCL reduce of the partial beams (strong scaling)."""
    import hzsdr as H
    import hzsdr_synth as O
    import hzsdr_shard as S
    torch, dist, rank, world, local = _dist_setup()
    ctx = H.Context(local)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))
    n, nbuf, nchan = w["n"], args.buffers, w["channels"]
    mine = S.channel_shard(nchan, world, rank)
    weights = H.beamform_angles(433e6, 30.0, [0.15 * c for c in range(nchan)])
    base = [O.synth_raw(w["fmt"], n, w["fs"], w["f0"], seed=c, phase=0.37 * c) for c in range(min(4, max(1, len(mine))))]
    chans = [[ctx.to_device(base[(b + c) % len(base)]) for c in range(len(mine))] for b in range(nbuf)]
    outs = [ctx.alloc(n * 8) for _ in range(nbuf)]
    comm = grp = None
    fused = world > 1 and args.beam_mode == "fused"
    if fused:
        grp = H.BeamGroup(ctx, world, rank, n)
        handles = [None] * world
        dist.all_gather_object(handles, grp.handle)
        grp.connect(handles)
        slices = [ctx.alloc(n // world * 8) for _ in range(nbuf)]
        packed = [H.BeamGroup.pack([c.ptr for c in chans[b]], weights[mine.start:mine.stop]) for b in range(nbuf)]
    elif world > 1:
        uid = [H.Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        comm = H.Comm(ctx, world, rank, uid[0])

    def step():
        for b in range(nbuf):
            if fused:
                grp.exec_packed(w["fmt"], packed[b], slices[b].ptr)
                if b == nbuf - 1:
                    grp.join()  # the step's slices are complete on the library's stream
                continue
            if len(mine):
                ctx.beamform(w["fmt"], [c.ptr for c in chans[b]], weights[mine.start:mine.stop], n, outs[b].ptr)
            else:
                H._check(H.load().hzsdr_dev_memset(ctx.h, outs[b].ptr, 0, n * 8))
            if comm is not None:
                comm.reduce_c64(outs[b].ptr, n, 0)
    for _ in range(args.warmup):
        step()
    ms, clocks = _time_region(torch, dist, ctx, stream, local, args.steps, step)
    e2e = None
    if world == 1:
        # end to end on one GPU: 64 raw channels in one pinned block -> hzsdr_beamform_submit_host (time
        # slices staged across PCIe, H2D / kernel / D2H overlapped) -> the beam in pinned host memory
        e2e_steps = max(3, min(10, args.steps))
        block = H.PinnedBuffer(nchan * n * w["raw"])
        rows = block.view(np.uint8).reshape(nchan, n * w["raw"])
        for c in range(nchan):
            rows[c] = base[c % len(base)].view(np.uint8).reshape(-1)
        beam = [H.PinnedBuffer(n * 8) for _ in range(2)]
        cp = [block.ptr + c * n * w["raw"] for c in range(nchan)]

        def e2e_step(i):
            ctx.beamform_submit_host(w["fmt"], cp, weights, n, beam[i & 1].ptr)
        e2e_step(0)
        ctx.wait_host()
        t0 = time.perf_counter()
        for i in range(e2e_steps * nbuf):
            e2e_step(i)
        ctx.wait_host()
        e2e_s = time.perf_counter() - t0
        h2d, d2h = nchan * n * w["raw"] * nbuf, n * 8 * nbuf
        e2e = {"value": nchan * n * nbuf * e2e_steps / e2e_s / 1e6, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "steps": e2e_steps, "api": "hzsdr_beamform_submit_host + hzsdr_ctx_wait_host: pinned H2D in time slices -> kernel -> pinned D2H",
               "pcie_gbs": (h2d + d2h) * e2e_steps / e2e_s / 1e9}
    if comm is not None:
        comm.close()
    if grp is not None:
        ctx.sync()
        dist.barrier()
        grp.close()
    if rank != 0:
        dist.destroy_process_group()
        return None
    pk = peaks()
    launches = args.steps * nbuf
    alg = len(mine) * n * w["raw"] + n * 8  # per launch on one GPU
    achieved = alg / ((ms / 1e3) / launches) / 1e9
    line = {"metric": "Msamples/s (channel-samples) through Convert->Multiply->Beamform", "unit": UNIT,
            "value": nchan * n * nbuf * args.steps / (ms / 1e3) / 1e6, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "c4: " + w["desc"], "buffers_per_step": nbuf, "channels_per_gpu": len(mine),
                       "collective": "none (1 GPU)" if world == 1 else (
                           "reduce-scatter fused into the beamform kernel: peer stores over NVLink into the owner rank's staging "
                           "slot + flag, then a local ordered sum (hzsdr_beam_group_*); result stays sliced across the GPUs"
                           if fused else "ncclReduce(sum, fp32, 2*2^20 floats) per buffer onto rank 0, in the timed region"),
                       "l2": f"{nbuf} distinct buffer sets per step = {nbuf * alg >> 20} MiB per GPU"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"],
                         "traffic": ncu_traffic("c4"), "kernel": "hz::k_beamform<U8>", "algorithmic_bytes_per_launch": alg,
                         "peak_source": pk["source"],
                         "note": "includes the reduce when n_gpus > 1; the kernel-only roofline is the 1-GPU line"}}
    if dist is not None:
        dist.destroy_process_group()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--buffers", type=int, default=0, help="buffers per step (default: 64 for c2, 16 for c3)")
    ap.add_argument("--cpu-reps", type=int, default=24)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--beam-mode", default="fused", choices=["fused", "nccl"], help="c4 at N>1: fused peer-memory reduce-scatter or ncclReduce")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    w = WORKLOADS[args.workload]
    if args.buffers <= 0:
        args.buffers = w.get("buffers", max(1, (1 << 28) // w["n"]))

    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return 0
        print(json.dumps(run_reference(args, w)), flush=True)
        return 0

    kind = w.get("kind", "chain")
    line = {"chain": run_ours, "beamform": run_beamform, "convert_shift": run_convert_shift,
            "channelizer": run_channelizer}[kind](args, w)
    if line is not None:
        print(json.dumps(line), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
