#!/usr/bin/env python
"""bench.py -- Msamples/s through Convert -> Shift -> FFT-convolution -> Decimate on B200.

    python bench.py --gpus N --steps K --warmup W            # our arm (libhzsdrcuda.so)
    python bench.py --impl reference --gpus N --steps K ...   # CPU arm: the reference's algorithm
                                                              #   on the box's host cores

Headline workload (BASELINE.json configs[1], "C2"): HackRF-format i8 IQ at 20 Msps in 2^22-sample
buffers -> sdr.ConvertBuffer(C64) -> stream.ShiftReader(-2.5 MHz) -> stream.ConvolutionReader(255-tap
lowpass as a 1024-bin frequency-domain filter; block-circular, the reference's semantics) ->
stream.DecimateReader(x10).  One *step* = one pass of that chain over `--buffers` consecutive
2^22-sample buffers of one stream (default 64 = 2^28 samples, 512 MiB of raw input: larger than the
126 MB L2, so every step streams from HBM).  With N GPUs every rank runs its own independent stream
(weak scaling, no data-path collective): that is `value`.

The same JSON line also carries
  e2e        the same metric through the pipelined host API (pinned H2D + kernel + pinned D2H inside
             the timed region), with the bare-copy ceiling of the box measured beside it;
  roofline   algorithmic bytes / CUDA-event launch time / measured HBM peak, plus the executed-FP32 figure
             (the chain is FP32-pipe-bound);
  sustained  the device-resident number again over a >= 2 s timed region, with clocks;
  cpu_baseline (N = 1)  the reference's algorithm on the host cores;
  extra      (N = 1) the other single-GPU BASELINE configs, compact: c3 (block-circular, the reference's
             semantics), c3os (overlap-save, BASELINE's wording), c1, c4, c5;
  sharded    (N > 1) BASELINE's sharded configs measured in the same run: c5 (512-stream channelizer, a
             FIXED job sharded s mod N, no collective) and c4 (64-channel Beamform sharded by channel)
             with the reduce-scatter fused into the kernel over NVLink peer memory and with ncclReduce --
             each with `parity_rel_l2` against the oracle, computed on rank 0 outside the timed region.

--workload c1|c3|c3os|c4|c5 prints one full-size line for that config instead.
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
    "c2": dict(kind="chain", fmt=4, fs=20_000_000, n=1 << 22, f0=2.5e6, taps=255, nfft=1024, D=10, raw=2,
               desc="i8 20 Msps, 2^22-sample buffers -> Convert -> Shift(-2.5 MHz) -> 255-tap FFT convolution "
                    "(N=1024, block-circular = reference semantics) -> Decimate x10"),
    "c3": dict(kind="chain", fmt=3, fs=61_440_000, n=1 << 24, f0=7.68e6, taps=4095, nfft=16384, D=16, raw=4,
               desc="i16 61.44 Msps, 2^24-sample buffers -> Convert -> Shift(-7.68 MHz) -> 4095-tap FFT convolution "
                    "(N=16384, block-circular = reference semantics) -> Decimate x16"),
    "c3os": dict(kind="chain", fmt=3, fs=61_440_000, n=1 << 24, f0=7.68e6, taps=4095, nfft=16384, D=16, raw=4, overlap_save=True,
                 desc="i16 61.44 Msps, 2^24-sample buffers -> Convert -> Shift(-7.68 MHz) -> 4095-tap OVERLAP-SAVE FIR "
                      "(windows of 16384 at hop 12288, history carried; true linear convolution) -> Decimate x16"),
    "c1": dict(kind="convert_shift", fmt=2, fs=2_400_000, n=1 << 20, f0=300e3, raw=2, buffers=256,
               desc="rtl u8 2.4 Msps, 2^20-sample buffers -> fused Convert + Shift(-300 kHz) (hzsdr_convert_shift_batch: 64 buffers per launch)"),
    "c5": dict(kind="channelizer", fmt=3, fs=61_440_000, n=1 << 20, f0=1e6, taps=255, nfft=1024, D=16, raw=4, streams=512, buffers=1,
               desc="channelizer fan-out: 512 independent i16 streams x 2^20 samples, each Convert -> Shift(own f) -> 255-tap FFT "
                    "convolution (N=1024) -> Decimate x16; streams sharded across the GPUs, no collective; one launch per 64 streams"),
    "c4": dict(kind="beamform", fmt=2, fs=2_400_000, n=1 << 20, f0=100e3, raw=2, channels=64, buffers=8,
               desc="64 coherent u8 channels x 2^20 samples -> Convert -> steering Multiply -> Beamform sum; channels "
                    "sharded across the GPUs, partial beams summed over NVLink"),
}

FP32_PEAK_TFLOPS = 74.0  # 148 SMs x 128 lanes x 2 flop x 1.965 GHz


def peaks() -> dict:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return {"hbm_gbs": float(p["hbm_gbs"]), "source": "MEASURED_PEAKS.json (measured)"}
    return {"hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def _profile_json(name: str) -> dict:
    try:
        with open(os.path.join(ROOT, "profiles", name)) as fh:
            return json.load(fh)
    except Exception:
        return {}


def ncu_traffic(workload: str):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    return _profile_json("ncu_traffic.json").get(workload)


def executed_flop_per_sample(workload: str):
    """fp32 flop the dominant kernel EXECUTES per input sample, from the opcode mix of the committed
    ncu capture (profiles/ncu_opmix.json: packed FFMA2 = 4 flop per lane, FADD2 / FMUL2 = 2, scalar
    FFMA = 2, FADD / FMUL = 1)."""
    return _profile_json("ncu_opmix.json").get(workload, {}).get("flop_per_sample_executed")


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


def taps_for(w: dict):
    import hzsdr_synth as Y
    return Y.lowpass_taps(w["taps"], 1.0 / (2 * w["D"]))


def filter_for(w: dict):
    import hzsdr_synth as Y
    return Y.filter_freq(taps_for(w), w["nfft"])


def chain_config(args, w: dict, world: int, nbuf: int) -> dict:
    """The `config` object: identical, key for key, in our arm and in the reference arm."""
    return {"workload": args.workload + ": " + w["desc"], "buffers_per_step": nbuf, "samples_per_buffer": w["n"],
            "parallelism": f"{world} independent stream(s), one per GPU, no collective",
            "outputs": "two sets of output buffers, used by alternate steps",
            "l2": f"each step streams {nbuf} distinct buffers per stream = {nbuf * w['n'] * w['raw'] >> 20} MiB of raw input "
                  "(> 126 MB L2); no explicit flush"}


# ---- CPU arm: the reference's algorithm on host cores -------------------------------------------
def host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


def host_env_note(threads: int) -> dict:
    return {"threads": threads, "cpu_count": os.cpu_count(), "OMP_NUM_THREADS": os.environ.get("OMP_NUM_THREADS"),
            "GOMAXPROCS": "n/a (C restatement; no Go toolchain in the image)",
            "fft": "scipy.fft (pocketfft, complex64), one transform batch per buffer per thread"}


class CpuChain:
    """The reference chain over one buffer on one host thread: the oracle's C twin for
    Convert/Shift/Decimate (oracle/cpu_ref.c, the reference's loop structure and SSE width) and
    scipy's pocketfft (complex64) as the user-supplied fft.Planner -- 3x faster than the C twin's
    plain radix-2 stand-in, so the faster (fairer) of the two is what gets timed."""

    def __init__(self, w: dict, filt: np.ndarray):
        import scipy.fft as sf

        import cpu_ref as CR
        self.w, self.filt, self.sf, self.CR = w, filt, sf, CR

    def run(self, raw: np.ndarray, ts0: float) -> int:
        w, CR, sf = self.w, self.CR, self.sf
        x = CR.convert_to_c64(raw, w["fmt"])
        y, _ = CR.shift_buffer(x, -w["f0"], w["fs"], ts0)
        nblk = y.size // w["nfft"]
        F = sf.fft(y[: nblk * w["nfft"]].reshape(nblk, w["nfft"]), axis=-1)
        F *= self.filt
        z = sf.ifft(F, axis=-1, norm="forward").reshape(-1)
        lz = (z.size // 32768) * 32768
        out = z[:lz].reshape(-1, 32768)[:, : (32768 // w["D"]) * w["D"] : w["D"]]
        return int(np.ascontiguousarray(out).size)


def buffer_start_times(w: dict, nbuf: int) -> list:
    """The carried NCO time at the start of each of `nbuf` consecutive buffers of one stream (the serial
    fp64 accumulator of stream/shifter.go:73-79, run once outside any timed region) -- lets the host
    threads work on different buffers of the SAME stream at once, which the reference's one goroutine
    per stage cannot: a generous baseline."""
    import cpu_ref as CR
    ts, out = 0.0, []
    for _ in range(nbuf):
        out.append(ts)
        _, ts = CR.shift_ts(w["fs"], w["n"], ts, want_array=False)
    return out


def cpu_stage_rates(w: dict, raw: np.ndarray, reps: int = 3) -> dict:
    """the reference's stages one by one on ONE host thread (Msamples/s of stage input,
    best of `reps`), and what one stream gets with a thread per stage as the reference's
    ReadTransformers run them (the slowest stage sets the pace)."""
    import scipy.fft as sf

    import cpu_ref as CR
    filt = filter_for(w)
    n = raw.shape[0] // 2

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


class CpuStep:
    """One bench step on the host: `nbuf` consecutive buffers (of each of `streams` streams) through the
    reference chain, the buffers dealt round-robin to all host threads."""

    def __init__(self, w: dict, nbuf: int, streams: int, threads: int, distinct: int = 8):
        import cpu_ref as CR
        CR.build()
        self.w, self.nbuf, self.streams, self.threads = w, nbuf, streams, threads
        self.bufs = make_buffers(w, min(distinct, nbuf), seed=2, distinct=min(distinct, nbuf))
        self.ts0 = buffer_start_times(w, nbuf)
        filt = filter_for(w)
        self.chains = [CpuChain(w, filt) for _ in range(threads)]
        self.pool = cf.ThreadPoolExecutor(max_workers=threads)
        self.jobs = [(s, b) for s in range(streams) for b in range(nbuf)]

    def _work(self, t: int):
        ch = self.chains[t]
        for s, b in self.jobs[t::self.threads]:
            ch.run(self.bufs[(s + b) % len(self.bufs)], self.ts0[b])

    def __call__(self):
        list(self.pool.map(self._work, range(self.threads)))

    @property
    def samples(self) -> int:
        return self.streams * self.nbuf * self.w["n"]


def run_reference(args, w: dict) -> dict:
    """The reference arm: the SAME step as our arm (same config object) on all host threads."""
    threads = host_threads()
    world = max(1, args.gpus)
    step = CpuStep(w, args.buffers, world, threads)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = step.samples * args.steps / dt / 1e6
    sample = (f"{world} stream(s) x {args.buffers} consecutive buffers of {w['n']} samples per step, dealt round-robin to "
              f"{threads} host threads")
    return {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": chain_config(args, w, world, args.buffers),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "host": host_env_note(threads),
                         "note": "C restatement of the reference loops (oracle/cpu_ref.c) + scipy pocketfft as the "
                                 "Planner; the Go reference cannot be built here (no Go toolchain; GOMAXPROCS n/a)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


# ---- our arm: environment -----------------------------------------------------------------------
class Env:
    """One process per GPU: device, torch.distributed (NCCL) for barriers / max-over-ranks, the library
    context and its stream as a torch ExternalStream (CUDA events are recorded on THAT stream)."""

    def __init__(self):
        import torch

        import hzsdr as H
        self.torch, self.H = torch, H
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist
        self.affinity = self._bind_to_gpu_cpus()  # before any pinned allocation or thread is made
        self.ctx = H.Context(self.local)  # raises without a B200: there is no CPU fallback
        self.stream = torch.cuda.ExternalStream(self.ctx.stream, device=torch.device("cuda", self.local))

    def _bind_to_gpu_cpus(self) -> dict:
        """The rank's threads (submit loop, ring producer) and its pinned host buffers (first touch) onto the CPUs NVML
        reports as local to the rank's GPU.  On a single-NUMA box this changes nothing; it is reported either way."""
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.local)
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
            cpus = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (int(wd) >> b) & 1}
            allowed = os.sched_getaffinity(0)
            self._allowed = set(allowed)
            use = (cpus & allowed) or allowed
            os.sched_setaffinity(0, use)
            return {"gpu_local_cpus": len(cpus), "bound_to": len(use), "allowed": len(allowed)}
        except Exception as e:
            return {"error": repr(e)[:120]}

    def unbind_cpus(self):
        try:
            os.sched_setaffinity(0, getattr(self, "_allowed", os.sched_getaffinity(0)))
        except Exception:
            pass

    def barrier(self):
        self.ctx.sync()
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()

    def max_over_ranks(self, x: float) -> float:
        if self.dist is None:
            return x
        t = self.torch.tensor([x], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def time_region(self, steps: int, step_fn, finish_fn=None):
        """Barrier + sync on both sides, CUDA events on the library's stream, max over ranks.
        `finish_fn` (still inside the timed region) makes the library's stream wait for work the steps
        left on side streams.  Returns (ms total, clocks summary).  Warm-ups are the caller's."""
        torch = self.torch
        self.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(self.local) as clocks:
            ev0.record(self.stream)
            for _ in range(steps):
                step_fn()
            if finish_fn is not None:
                finish_fn()
            ev1.record(self.stream)
            self.ctx.sync()
            torch.cuda.synchronize()
        ms = self.max_over_ranks(ev0.elapsed_time(ev1))
        self.barrier()
        return ms, clocks.summary()

    def wall_region(self, steps: int, step_fn, finish=None) -> float:
        """Host wall clock around `steps` end-to-end steps (+ `finish`), max over ranks."""
        self.barrier()
        t0 = time.perf_counter()
        for i in range(steps):
            step_fn(i)
        if finish is not None:
            finish()
        s = self.max_over_ranks(time.perf_counter() - t0)
        self.barrier()
        return s

    def copy_ceiling(self, h2d_bytes: int, d2h_bytes: int, pieces: int, reps: int = 8) -> dict:
        """The box's bare pinned-copy ceiling for an end-to-end step of this shape: the same bytes H2D
        and D2H per step, in the same number of pieces, as plain cudaMemcpyAsync on two streams, all
        ranks at once, nothing computed."""
        torch = self.torch
        dev = torch.device("cuda", self.local)
        hin = torch.empty(h2d_bytes, dtype=torch.uint8, pin_memory=True)
        hout = torch.empty(max(d2h_bytes, 1), dtype=torch.uint8, pin_memory=True)
        din = torch.empty(h2d_bytes, dtype=torch.uint8, device=dev)
        dout = torch.empty(max(d2h_bytes, 1), dtype=torch.uint8, device=dev)
        s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        pi, po = h2d_bytes // pieces, d2h_bytes // pieces
        ci = [(din[k * pi:(k + 1) * pi], hin[k * pi:(k + 1) * pi]) for k in range(pieces)]
        co = [(hout[k * po:(k + 1) * po], dout[k * po:(k + 1) * po]) for k in range(pieces)]

        def once(_i=0):
            for (d, h), (h2, d2) in zip(ci, co):
                with torch.cuda.stream(s_in):
                    d.copy_(h, non_blocking=True)
                with torch.cuda.stream(s_out):
                    h2.copy_(d2, non_blocking=True)

        def finish():
            s_in.synchronize()
            s_out.synchronize()
        once()
        finish()
        s = self.wall_region(reps, once, finish)
        return {"seconds_per_step": s / reps, "gbs_per_gpu": (h2d_bytes + d2h_bytes) * reps / s / 1e9,
                "gbs_all_gpus": (h2d_bytes + d2h_bytes) * reps * self.world / s / 1e9,
                "how": "pinned cudaMemcpyAsync H2D and D2H of the step's byte counts on two streams, all ranks "
                       "concurrently, no kernel; wall clock, max over ranks"}

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


def hbm_roofline(alg_bytes: int, launch_s: float, workload: str, kernel: str, **more) -> dict:
    pk = peaks()
    achieved = alg_bytes / launch_s / 1e9
    r = {"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"],
         "traffic": ncu_traffic(workload), "kernel": kernel, "algorithmic_bytes_per_launch": alg_bytes,
         "launch_us": launch_s * 1e6, "peak_source": pk["source"]}
    r.update(more)
    return r


# ---- chain workloads (c2, c3, c3os) -------------------------------------------------------------
def measure_chain(env: Env, args, w: dict, name: str, steps: int, warmup: int, nbuf: int, full: bool) -> dict:
    """Device-resident and end-to-end throughput of one stream per rank through the fused chain.
    `full`: also the sustained run and the copy ceiling (the headline line); extras skip them."""
    H, ctx = env.H, env.ctx
    n = w["n"]
    filt = filter_for(w)
    distinct = 8 if full else 3
    host_bufs = make_buffers(w, nbuf, seed=2 + env.rank, distinct=distinct)

    def new_chain():
        if w.get("overlap_save"):
            return H.Chain(ctx, w["fmt"], w["fs"], -w["f0"], filt, w["D"], overlap_save_taps=w["taps"])
        return H.Chain(ctx, w["fmt"], w["fs"], -w["f0"], filt, w["D"])

    chain = new_chain()
    per_out = chain.out_len(n)
    src = [ctx.to_device(b) for b in host_bufs[: min(nbuf, distinct)]]
    pool = []  # nbuf distinct device buffers (copies of the distinct host buffers)
    for i in range(nbuf):
        if i < len(src):
            pool.append(src[i])
        else:
            d = ctx.alloc(n * w["raw"])
            H._check(H.load().hzsdr_copy(ctx.h, d.ptr, src[i % len(src)].ptr, n * w["raw"]))
            pool.append(d)
    # two output sets, used alternately: consecutive steps then touch disjoint memory and their launches may overlap
    # (a step that rewrote the buffers the previous launch is still writing is serialised by the library's hazard check)
    outs = [ctx.alloc(per_out * 8) for _ in range(nbuf)]
    outs_b = [ctx.alloc(per_out * 8) for _ in range(nbuf)]
    ctx.sync()
    packed = H.Chain.pack_batch([p.ptr for p in pool], [o.ptr for o in outs])
    packed_b = H.Chain.pack_batch([p.ptr for p in pool], [o.ptr for o in outs_b])
    flip = [0]

    def step_device():
        flip[0] ^= 1
        chain.exec_batch(packed if flip[0] else packed_b, n, per_out)  # nbuf consecutive buffers of the stream in one call

    # N = 1024 chains with an even decimation factor: hzsdr_chain_exec_batch is ONE launch per <= 64 buffers (descriptors
    # in the kernel parameters); the other kernels launch per buffer (overlapped)
    bufs_per_launch = min(nbuf, 64) if (w["nfft"] == 1024 and w["D"] % 2 == 0 and nbuf >= 8 and not w.get("overlap_save")) else 1
    for _ in range(warmup):
        step_device()
    ms, clocks = env.time_region(steps, step_device)
    launches = steps * ((nbuf + bufs_per_launch - 1) // bufs_per_launch)
    samples_per_step = nbuf * n * env.world
    value = samples_per_step * steps / (ms / 1e3) / 1e6
    out = {"value": value, "unit": UNIT, "ms_per_step": ms / steps, "steps": steps, "warmup": warmup, "clocks": clocks,
           "gpu_launches": launches, "timed_region_ms": ms}

    if full:  # the same thing over a >= 2 s timed region (clocks, power and the NCO wrap all in steady state)
        sus_steps = max(steps, int(2200.0 / max(ms / steps, 1e-3)))
        ms_s, clocks_s = env.time_region(sus_steps, step_device)
        out["sustained"] = {"value": samples_per_step * sus_steps / (ms_s / 1e3) / 1e6, "unit": UNIT, "steps": sus_steps,
                            "seconds": ms_s / 1e3, "clocks": clocks_s, "gpu_launches": sus_steps * nbuf}

    # ---- end to end: pinned host buffers, H2D + kernel + D2H per buffer, pipelined ----
    e2e_chain = new_chain()
    pin_in = H.PinnedBuffer(nbuf * n * w["raw"])
    pin_out = H.PinnedBuffer(nbuf * per_out * 8)
    view = pin_in.view(H.NP_DTYPE[w["fmt"]])
    for i in range(nbuf):
        view[i * 2 * n:(i + 1) * 2 * n] = host_bufs[i]

    def step_e2e(_i=0):
        for i in range(nbuf):
            e2e_chain.submit_host(pin_in.ptr + i * n * w["raw"], n, pin_out.ptr + i * per_out * 8, per_out)
        e2e_chain.wait_host()  # the step's result is in host memory

    for _ in range(max(1, warmup // 2)):
        step_e2e()
    e2e_steps = max(2, min(steps, 20 if full else 5))
    e2e_s = env.wall_region(e2e_steps, step_e2e)
    h2d, d2h = nbuf * n * w["raw"], nbuf * per_out * 8
    out["e2e"] = {"value": samples_per_step * e2e_steps / e2e_s / 1e6, "unit": UNIT, "h2d_bytes_per_step": h2d,
                  "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                  "api": "hzsdr_chain_submit_host/wait_host: pinned H2D -> fused kernel -> pinned D2H, 3-deep pipeline",
                  "pcie_gbs_per_gpu": (h2d + d2h) * e2e_steps / e2e_s / 1e9,
                  "pcie_gbs_all_gpus": (h2d + d2h) * e2e_steps * env.world / e2e_s / 1e9}
    if full:
        ceil = env.copy_ceiling(h2d, d2h, nbuf)
        out["e2e"]["copy_ceiling"] = ceil
        out["e2e"]["frac_of_copy_ceiling"] = ceil["seconds_per_step"] / (e2e_s / e2e_steps)
        out["e2e"]["pcie_gen5_x16_nominal_gbs_per_direction"] = 64.0
        if not w.get("overlap_save"):
            out["e2e"]["ring"] = measure_ring_path(env, w, new_chain, host_bufs, nbuf, per_out, e2e_steps)

    alg_bytes = (n * w["raw"] + per_out * 8) * bufs_per_launch
    launch_s = (ms / 1e3) / launches
    kernel = {1024: "hz::k_chain1024", 16384: "hz::k_chain16k"}.get(w["nfft"], f"hz::k_chain<{w['nfft']}>") + f"<fmt {w['fmt']}>"
    if w.get("overlap_save"):
        kernel = "hz::k_chain16k_os" + f"<fmt {w['fmt']}>"
    log2n = w["nfft"].bit_length() - 1
    nominal = 2 * 5 * log2n + 6 + 30  # unpruned textbook count: FFT pair + pointwise + convert/NCO
    per_gpu = value / env.world * 1e6
    roof = hbm_roofline(alg_bytes, launch_s, name, kernel, bytes_per_sample=alg_bytes / (n * bufs_per_launch),
                        buffers_per_launch=bufs_per_launch,
                        note="the fused chain is FP32-pipe-bound, not HBM-bound (SURVEY.md 8(d)): fp32 figures alongside",
                        fp32_peak_tflops=FP32_PEAK_TFLOPS, fp32_flop_per_sample_nominal=nominal,
                        fp32_frac_nominal=nominal * per_gpu / 1e12 / FP32_PEAK_TFLOPS)
    ex = executed_flop_per_sample(name)
    if ex:
        roof["fp32_flop_per_sample_executed"] = ex
        roof["fp32_tflops_executed"] = ex * per_gpu / 1e12
        roof["fp32_frac_executed"] = ex * per_gpu / 1e12 / FP32_PEAK_TFLOPS
        roof["fp32_source"] = "opcode mix of the committed ncu capture (profiles/ncu_opmix.json)"
    out["roofline"] = roof
    out["host_bufs"] = host_bufs
    chain.close()
    e2e_chain.close()
    return out


def measure_ring_path(env: Env, w: dict, new_chain, host_bufs, nbuf: int, per_out: int, steps: int) -> dict:
    """The driver hand-off end to end: a PRODUCER THREAD hands pinned ring slots over with
    hzsdr_ring_write_peek / write_poke (the synthetic "driver" has filled the slot's pinned memory, as an SDR driver's
    callback would -- UnsafeRingBuffer.WritePeekUnsafePointer, stream/ring.go:344-392), this thread drains them with
    hzsdr_chain_submit_ring: H2D on the ring's copy stream, fused kernel, D2H to pinned host memory, no host wait
    in between.  The producer is paced by a semaphore of free slots (an overrun would drop the oldest slot)."""
    import ctypes as C
    H, ctx = env.H, env.ctx
    n, slots = w["n"], 8
    ring = H.Ring(ctx, w["fmt"], slots, n)
    chain = new_chain()
    pin_out = H.PinnedBuffer(nbuf * per_out * 8)
    free_slots = threading.Semaphore(slots)
    filled = [0]

    def producer(count):
        for _ in range(count):
            free_slots.acquire()
            p = ring.write_peek()
            if filled[0] < slots:  # first lap: the driver's samples land in the slot; later laps reuse what is there
                b = host_bufs[filled[0] % len(host_bufs)]
                C.memmove(p, b.ctypes.data, b.nbytes)
                filled[0] += 1
            ring.write_poke(n)

    def run(count):
        t = threading.Thread(target=producer, args=(count,))
        t.start()
        done = 0
        while done < count:
            try:
                chain.submit_ring(ring, pin_out.ptr + (done % nbuf) * per_out * 8, per_out)
                done += 1
                free_slots.release()
            except H.HzsdrError as e:
                if e.status != H.ERR_RING_UNDERRUN:
                    raise
        t.join()
        chain.wait_host()

    run(2 * slots)  # warm-up: slots filled, staging allocated
    env.barrier()
    t0 = time.perf_counter()
    run(steps * nbuf)
    dt = env.max_over_ranks(time.perf_counter() - t0)
    h2d, d2h = nbuf * n * w["raw"], nbuf * per_out * 8
    rec = {"value": nbuf * n * env.world * steps / dt / 1e6, "unit": UNIT, "steps": steps, "slots": slots,
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "pcie_gbs_per_gpu": (h2d + d2h) * steps / dt / 1e9,
           "api": "producer thread: hzsdr_ring_write_peek/write_poke (pinned slots) -> hzsdr_chain_submit_ring -> pinned D2H; "
                  "hzsdr_chain_wait_host at the end"}
    ring.close()
    chain.close()
    return rec


def run_chain_line(env: Env, args, w: dict) -> dict:
    m = measure_chain(env, args, w, args.workload, args.steps, args.warmup, args.buffers, full=True)
    host_bufs = m.pop("host_bufs")
    line = {"metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": env.world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": chain_config(args, w, env.world, args.buffers),
            "timing": "CUDA events on the library's stream, barrier + synchronize on both sides, max over ranks",
            "clocks": m["clocks"], "e2e": m["e2e"], "gpu_launches": m["gpu_launches"], "roofline": m["roofline"],
            "cpu_affinity": env.affinity,
            "sustained": m["sustained"], "value_sustained": m["sustained"]["value"], "timed_region_ms": m["timed_region_ms"]}
    if env.world == 1 and not args.no_cpu_baseline and env.rank == 0:
        env.unbind_cpus()  # the CPU leg gets every core the process was allowed, like `--impl reference`
        threads = host_threads()
        nb = max(threads, min(args.buffers, 2 * threads))
        step = CpuStep(w, nb, 1, threads)
        step()
        reps = max(1, args.cpu_reps // max(1, nb // threads))
        t0 = time.perf_counter()
        for _ in range(reps):
            step()
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {
            "value": step.samples * reps / dt / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{reps} x {nb} consecutive buffers of {w['n']} samples of one stream, dealt round-robin to {threads} host threads, {dt:.1f} s",
            "host": host_env_note(threads),
            "stages_one_thread": cpu_stage_rates(w, host_bufs[0]),
            "note": "oracle/cpu_ref.c loops + scipy pocketfft as the Planner; Go reference not buildable here (no Go); "
                    "stages_one_thread: Msamples/s of each stage alone on one thread, and of one stream with a thread per "
                    "stage (the reference's goroutine-per-ReadTransformer layout) / with everything on one thread"}
    else:
        line["cpu_baseline"] = None
    return line


# ---- C1: fused convert + shift ------------------------------------------------------------------
def measure_convert_shift(env: Env, w: dict, steps: int, warmup: int, nbuf: int) -> dict:
    """C1 on the GPU: fused u8 -> complex64 -> NCO mix, HBM-bound (2 + 8 B per sample)."""
    import hzsdr_synth as Y
    H, ctx = env.H, env.ctx
    n = w["n"]
    raw = [Y.synth_raw(w["fmt"], n, w["fs"], w["f0"], seed=env.rank * 100 + i) for i in range(4)]
    src = [ctx.to_device(raw[i % 4]) for i in range(nbuf)]
    dst = [ctx.alloc(n * 8) for _ in range(nbuf)]
    st = H.NcoState(w["fs"], 0.0)

    packed = H.Chain.pack_batch([s_.ptr for s_ in src], [d.ptr for d in dst])

    def step():  # nbuf consecutive buffers of the stream in one call: one launch per 64 buffers
        ctx.convert_shift_batch(w["fmt"], packed, n, n, -w["f0"], st)
    for _ in range(warmup):
        step()
    ms, clocks = env.time_region(steps, step)
    per_launch = min(nbuf, 64)
    launches = steps * ((nbuf + per_launch - 1) // per_launch)
    alg = n * (w["raw"] + 8) * per_launch
    return {"metric": "Msamples/s through fused Convert->Shift", "value": nbuf * n * env.world * steps / (ms / 1e3) / 1e6, "unit": UNIT,
            "ms_per_step": ms / steps, "steps": steps, "warmup": warmup, "scaling": "weak", "clocks": clocks, "gpu_launches": launches,
            "config": {"workload": "c1: " + w["desc"], "buffers_per_step": nbuf,
                       "l2": f"{nbuf} distinct buffer pairs = {nbuf * alg >> 20} MiB per step (> 126 MB L2)"},
            "roofline": hbm_roofline(alg, (ms / 1e3) / launches, "c1", "hz::k_shift_batch<U8, 4>", buffers_per_launch=per_launch)}


def measure_polyphase(env: Env, w: dict, steps: int, warmup: int) -> dict:
    """The fused polyphase decimator (hzsdr_polyphase_*: Convert -> Shift -> real-tap FIR -> keep every D-th sample, true
    linear convolution) on C2's format, rate, filter length and decimation -- the measurement behind "the FFT chain is the
    C2 path" (DESIGN.md 4.7).  Per 2^22-sample buffer (C2's buffer size) and per 2^24-sample call."""
    import hzsdr_synth as Y
    H, ctx = env.H, env.ctx
    taps = np.asarray(taps_for(w), dtype=np.float32)
    out = {"metric": "Msamples/s through fused Convert->Shift->polyphase FIR->decimate", "unit": UNIT, "taps": int(taps.size), "decimate": w["D"],
           "kernel": "hz::k_polyphase_chain"}
    for key, n, nbuf in (("value", w["n"], 32), ("value_2p24_calls", 1 << 24, 8)):
        pp = H.Polyphase(ctx, w["fmt"], w["fs"], -w["f0"], taps, w["D"])
        raw = [ctx.to_device(Y.synth_raw(w["fmt"], n, w["fs"], w["f0"], seed=env.rank * 10 + i)) for i in range(2)]
        pool = []
        for i in range(nbuf):
            d = ctx.alloc(n * w["raw"])
            H._check(H.load().hzsdr_copy(ctx.h, d.ptr, raw[i & 1].ptr, n * w["raw"]))
            pool.append(d)
        per = n // w["D"] + 1
        outs = [ctx.alloc(per * 8) for _ in range(nbuf)]

        def step():
            for i in range(nbuf):
                pp.exec(pool[i].ptr, n, outs[i].ptr, per)
        for _ in range(warmup):
            step()
        ms, clocks = env.time_region(steps, step)
        out[key] = nbuf * n * env.world * steps / (ms / 1e3) / 1e6
        if key == "value":
            out.update({"ms_per_step": ms / steps, "steps": steps, "gpu_launches": steps * nbuf, "clocks": clocks,
                        "config": {"workload": f"c2 shape through the polyphase decimator: {nbuf} x 2^22-sample buffers per step"}})
        pp.close()
    return out


# ---- C5: channelizer ----------------------------------------------------------------------------
def measure_channelizer(env: Env, w: dict, steps: int, warmup: int, with_e2e: bool) -> dict:
    """C5: the rank's share of 512 independent streams through hzsdr_channelizer_exec (the total is fixed
    at 512 streams, sharded s mod N: strong scaling of a fixed job, no collective).  Parity: rank 0's
    first stream, last timed buffer, against the oracle started from the carried NCO time."""
    import hzsdr_shard as S
    import hzsdr_synth as Y
    H, ctx = env.H, env.ctx
    n = w["n"]
    mine = S.stream_shard(w["streams"], env.world, env.rank)
    shift_of = lambda s: -(w["f0"] + 10e3 * s)  # noqa: E731
    shifts = [shift_of(s) for s in mine]
    filt = filter_for(w)
    raw_host = [Y.synth_raw(w["fmt"], n, w["fs"], w["f0"] + 10e3 * i, seed=i) for i in range(4)]
    base = [ctx.to_device(r) for r in raw_host]
    srcs = []
    for i, _ in enumerate(mine):
        d = ctx.alloc(n * w["raw"])
        H._check(H.load().hzsdr_copy(ctx.h, d.ptr, base[i % 4].ptr, n * w["raw"]))
        srcs.append(d)
    chz = H.Channelizer(ctx, w["fmt"], w["fs"], shifts, filt, w["D"])
    per = n // 32768 * (32768 // w["D"])
    # two output sets, used alternately (what a pipeline that hands a step's output on does): a step that wrote into the
    # buffers the previous launch is still writing would be serialised behind it by the library's hazard check
    dsts = [ctx.alloc(per * 8) for _ in mine]
    dsts_b = [ctx.alloc(per * 8) for _ in mine]
    sp, dp, dp_b = [x.ptr for x in srcs], [x.ptr for x in dsts], [x.ptr for x in dsts_b]
    flip = [0]

    def step():
        flip[0] ^= 1
        chz.exec(sp, n, dp if flip[0] else dp_b, per)
    for _ in range(warmup + 1):  # the first buffer of a stream takes the long segment tables
        step()
    ms, clocks = env.time_region(steps, step)

    # parity of what was just timed: one more buffer from the carried state, checked on rank 0
    parity = None
    ts_before = chz.ts.copy()
    flip[0] = 0  # the checked step writes `dsts`
    step()
    ctx.sync()
    if env.rank == 0:
        import go_sdr_oracle as O  # the checker, outside every timed region
        got = dsts[0].download(np.complex64, per)
        want, ts_want = O.chain(raw_host[0], w["fmt"], w["fs"], shifts[0], filt, w["D"], ts0=float(ts_before[0]))
        parity = {"parity_rel_l2": float(O.rel_l2(got, want)), "ts_bit_equal": bool(chz.ts[0] == ts_want),
                  "checked": f"stream {mine[0]}, the buffer after the timed region (carried ts {ts_before[0]:.6f} s), all {per} outputs"}

    alg = len(mine) * (n * w["raw"] + per * 8)
    nl = (len(mine) + 63) // 64  # hzsdr_channelizer_exec: launches of <= 64 streams (descriptors in the kernel parameters), overlapped
    out = {"metric": METRIC + " (512-stream channelizer)", "value": w["streams"] * n * steps / (ms / 1e3) / 1e6, "unit": UNIT,
           "ms_per_step": ms / steps, "steps": steps, "warmup": warmup, "scaling": "strong", "clocks": clocks, "gpu_launches": steps * nl,
           "config": {"workload": "c5: " + w["desc"], "streams_per_gpu": len(mine), "parallelism": "streams s mod G, no collective",
                      "l2": f"{alg >> 20} MiB touched per GPU per step (> 126 MB L2 up to 8 GPUs)"},
           "roofline": hbm_roofline(alg // nl, (ms / 1e3) / steps / nl, "c5", "hz::k_chain1024<I16, batch>", launches_per_step=nl,
                                    note="FP32-pipe-bound like C2; traffic = the ncu capture of one 512-stream launch")}
    if parity:
        out.update(parity)
    if with_e2e:
        # every stream's raw samples start in pinned host memory and its decimated output ends there
        e2e_steps = max(3, min(10, steps))
        pin_in = [H.PinnedBuffer(n * w["raw"]) for _ in mine]
        pin_out = [H.PinnedBuffer(per * 8) for _ in mine]
        for i, b in enumerate(pin_in):
            b.view(np.uint8)[:] = raw_host[i % 4].view(np.uint8).reshape(-1)
        hp, op = [b.ptr for b in pin_in], [b.ptr for b in pin_out]
        chz.submit_host(hp, n, op, per)
        ctx.wait_host()
        e2e_s = env.wall_region(e2e_steps, lambda i: chz.submit_host(hp, n, op, per), ctx.wait_host)
        h2d, d2h = len(mine) * n * w["raw"], len(mine) * per * 8
        out["e2e"] = {"value": w["streams"] * n * e2e_steps / e2e_s / 1e6, "unit": UNIT, "h2d_bytes_per_step": h2d * env.world,
                      "d2h_bytes_per_step": d2h * env.world, "steps": e2e_steps,
                      "api": "hzsdr_channelizer_submit_host + hzsdr_ctx_wait_host: pinned H2D in stream groups -> batched kernel -> pinned D2H",
                      "pcie_gbs_per_gpu": (h2d + d2h) * e2e_steps / e2e_s / 1e9}
    chz.close()
    return out


# ---- C4: beamform -------------------------------------------------------------------------------
def measure_beamform(env: Env, w: dict, steps: int, warmup: int, nbuf: int, mode: str, with_e2e: bool) -> dict:
    """C4: channels sharded across ranks by contiguous blocks.  mode "fused": ONE exchange per step -- the
    reduce-scatter fused into the beamform kernel over NVLink peer memory (hzsdr_beam_group_exec_batch over
    the step's `nbuf` buffers), the beam stays sliced across the GPUs.  mode "nccl": per buffer, the partial
    beam + ncclReduce onto rank 0.  One GPU: hzsdr_beamform per buffer.  Parity: the first buffer's beam
    (rank 0's slice when fused) against the oracle over all 64 channels."""
    import hzsdr_shard as S
    import hzsdr_synth as Y
    H, ctx, dist = env.H, env.ctx, env.dist
    n, nchan, world, rank = w["n"], w["channels"], env.world, env.rank
    mine = S.channel_shard(nchan, world, rank)
    weights = H.beamform_angles(433e6, 30.0, [0.15 * c for c in range(nchan)])
    NB = 4  # distinct raw buffers; channel c of buffer b is base[(b + c) % NB]
    base_host = [Y.synth_raw(w["fmt"], n, w["fs"], w["f0"], seed=c, phase=0.37 * c) for c in range(NB)]
    chans = [[ctx.to_device(base_host[(b + c) % NB]) for c in mine] for b in range(nbuf)]
    fused = world > 1 and mode == "fused"
    comm = grp = None
    if fused:
        grp = H.BeamGroup(ctx, world, rank, n, max_batch=nbuf)
        handles = [None] * world
        dist.all_gather_object(handles, grp.handle)
        grp.connect(handles)
        slices = [ctx.alloc(n // world * 8) for _ in range(nbuf)]
        packed = H.BeamGroup.pack_batch([[c.ptr for c in chans[b]] for b in range(nbuf)], weights[mine.start:mine.stop],
                                        [s.ptr for s in slices])
    else:
        outs = [ctx.alloc(n * 8) for _ in range(nbuf)]
        if world > 1:
            uid = [H.Comm.unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            comm = H.Comm(ctx, world, rank, uid[0])

    def step():
        if fused:
            grp.exec_batch_packed(w["fmt"], packed)
            return
        for b in range(nbuf):
            if len(mine):
                ctx.beamform(w["fmt"], [c.ptr for c in chans[b]], weights[mine.start:mine.stop], n, outs[b].ptr)
            else:
                H._check(H.load().hzsdr_dev_memset(ctx.h, outs[b].ptr, 0, n * 8))
            if comm is not None:
                comm.reduce_c64(outs[b].ptr, n, 0)

    # fused: the finishing sums run on a side stream (step s's overlaps step s+1's compute); join makes the
    # library's stream wait for all of them before the closing event, so the timed region contains every one
    join = grp.join if fused else None
    for _ in range(warmup):
        step()
    if fused:
        grp.join()
    ms, clocks = env.time_region(steps, step, join)

    parity = None
    ctx.sync()
    if rank == 0:
        import go_sdr_oracle as O  # the checker, outside every timed region
        lo, hi = (0, n // world) if fused else (0, n)
        all_ch = [base_host[(0 + c) % NB][2 * lo:2 * hi] for c in range(nchan)]
        want = O.beamform(all_ch, w["fmt"], weights)
        got = (slices[0] if fused else outs[0]).download(np.complex64, hi - lo)
        parity = {"parity_rel_l2": float(O.rel_l2(got, want)),
                  "checked": f"buffer 0 of the last step, samples [{lo}, {hi}) on rank 0, all {nchan} channels in the oracle"}

    launches = steps * (1 if fused else nbuf)
    alg = nbuf * (len(mine) * n * w["raw"] + (n // world if fused else n) * 8)  # HBM bytes per step on one GPU
    step_s = (ms / 1e3) / steps
    out = {"metric": "Msamples/s (channel-samples) through Convert->Multiply->Beamform", "unit": UNIT,
           "value": nchan * n * nbuf * steps / (ms / 1e3) / 1e6, "ms_per_step": ms / steps, "steps": steps, "warmup": warmup,
           "scaling": "strong", "clocks": clocks, "gpu_launches": launches,
           "config": {"workload": "c4: " + w["desc"], "buffers_per_step": nbuf, "channels_per_gpu": len(mine),
                      "collective": "none (1 GPU)" if world == 1 else (
                          "reduce-scatter fused into the beamform kernel: peer stores over NVLink into the owner rank's staging "
                          f"slot + flag + ack, then a local ordered sum; ONE exchange per step of {nbuf} buffers "
                          "(hzsdr_beam_group_exec_batch); the beam stays sliced across the GPUs"
                          if fused else "ncclReduce(sum, fp32, 2*2^20 floats) per buffer onto rank 0, in the timed region"),
                      "l2": f"{nbuf} distinct buffer sets per step = {alg >> 20} MiB per GPU"},
           "roofline": hbm_roofline(alg, step_s, "c4", "hz::k_beamform_rs<U8>" if fused else "hz::k_beamform<U8>",
                                    note="per STEP on one GPU; includes the exchange when n_gpus > 1")}
    if world > 1:
        nv = nbuf * n * 8 * (world - 1) / world  # bytes this rank sends (and receives) per step, reduce-scatter
        out["nvlink"] = {"bytes_out_per_gpu_per_step": nv, "achieved_gbs_per_direction": nv / step_s / 1e9,
                         "nominal_gbs_per_direction": 900.0,
                         "note": "reduce-scatter minimum for channel sharding: (G-1)/G x 8 B per output sample each way; "
                                 "with ncclReduce the root alone receives (G-1) x 8 B per sample"}
    if parity:
        out.update(parity)
    if with_e2e and world == 1:
        # 64 raw channels in one pinned block -> hzsdr_beamform_submit_host -> the beam in pinned host memory
        e2e_steps = max(3, min(10, steps))
        block = H.PinnedBuffer(nchan * n * w["raw"])
        rows = block.view(np.uint8).reshape(nchan, n * w["raw"])
        for c in range(nchan):
            rows[c] = base_host[c % NB].view(np.uint8).reshape(-1)
        beam = [H.PinnedBuffer(n * 8) for _ in range(2)]
        cp = [block.ptr + c * n * w["raw"] for c in range(nchan)]
        ctx.beamform_submit_host(w["fmt"], cp, weights, n, beam[0].ptr)
        ctx.wait_host()
        e2e_s = env.wall_region(e2e_steps * nbuf, lambda i: ctx.beamform_submit_host(w["fmt"], cp, weights, n, beam[i & 1].ptr),
                                ctx.wait_host)
        h2d, d2h = nchan * n * w["raw"] * nbuf, n * 8 * nbuf
        out["e2e"] = {"value": nchan * n * nbuf * e2e_steps / e2e_s / 1e6, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                      "steps": e2e_steps, "api": "hzsdr_beamform_submit_host + hzsdr_ctx_wait_host: pinned H2D in time slices -> kernel -> pinned D2H",
                      "pcie_gbs_per_gpu": (h2d + d2h) * e2e_steps / e2e_s / 1e9}
    if comm is not None:
        ctx.sync()
        comm.close()
    if grp is not None:
        env.barrier()
        grp.close()
    return out


def compact(m: dict) -> dict:
    """An extra / sharded record: the measured part without the bulky bookkeeping."""
    keep = ("metric", "value", "unit", "ms_per_step", "steps", "scaling", "gpu_launches", "parity_rel_l2", "ts_bit_equal", "checked",
            "nvlink", "e2e", "roofline", "clocks", "config", "value_2p24_calls", "taps", "decimate", "kernel")
    out = {k: m[k] for k in keep if k in m}
    if "config" in out:
        out["config"] = {k: v for k, v in out["config"].items() if k in ("workload", "buffers_per_step", "streams_per_gpu",
                                                                         "channels_per_gpu", "collective")}
    return out


def as_line(env: Env, args, m: dict) -> dict:
    """A stand-alone line (--workload c1|c4|c5) in the bench contract's shape."""
    line = {"metric": m["metric"], "value": m["value"], "unit": m["unit"], "n_gpus": env.world, "steps": m["steps"], "warmup": m["warmup"],
            "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": m["scaling"], "vs_baseline": None, "dtype": "f32",
            "data": "synthetic"}
    line.update({k: v for k, v in m.items() if k not in line})
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--buffers", type=int, default=0, help="buffers per step (default: 2^28 samples' worth: 64 for c2, 16 for c3)")
    ap.add_argument("--cpu-reps", type=int, default=24, help="buffers per host thread in the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra (N=1) / sharded (N>1) records")
    ap.add_argument("--beam-mode", default="fused", choices=["fused", "nccl"], help="c4 at N>1: fused peer-memory reduce-scatter or ncclReduce")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    w = WORKLOADS[args.workload]
    if args.buffers <= 0:
        args.buffers = w.get("buffers", max(1, (1 << 28) // w["n"]))

    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return 0
        if w["kind"] != "chain" or w.get("overlap_save"):
            print(json.dumps({"impl": "reference", "unavailable": "the CPU arm times the chain workloads c2 / c3"}), flush=True)
            return 0
        print(json.dumps(run_reference(args, w)), flush=True)
        return 0

    env = Env()
    kind = w["kind"]
    if kind == "chain":
        line = run_chain_line(env, args, w)
        if not args.no_extras and args.workload == "c2":
            short = max(3, min(args.steps, 10))
            if env.world == 1:
                ex = {}
                for name in ("c3", "c3os"):
                    try:
                        m = measure_chain(env, args, WORKLOADS[name], name, short, 3, 8, full=False)
                        m.pop("host_bufs", None)
                        m["config"] = {"workload": name + ": " + WORKLOADS[name]["desc"], "buffers_per_step": 8}
                        ex[name] = compact(m)
                    except Exception as e:  # an extra must never take the headline down with it
                        ex[name] = {"error": repr(e)}
                for name, fn in (("c1", lambda: measure_convert_shift(env, WORKLOADS["c1"], short, 3, 256)),
                                 ("c4", lambda: measure_beamform(env, WORKLOADS["c4"], 4 * short, 3, 8, "fused", True)),  # (a step is 0.2 ms)
                                 ("c5", lambda: measure_channelizer(env, WORKLOADS["c5"], short, 3, True)),
                                 ("c2_polyphase", lambda: measure_polyphase(env, WORKLOADS["c2"], short, 3))):
                    try:
                        ex[name] = compact(fn())
                    except Exception as e:
                        ex[name] = {"error": repr(e)}
                line["extra"] = ex
            else:
                sh = {}
                for name, fn in (("c5", lambda: measure_channelizer(env, WORKLOADS["c5"], short, 3, True)),
                                 # (32 buffers per exchange: at 8 GPUs an exchange costs ~30 us of flag / ack / launch latency
                                 # on top of ~14 us per 2^20-sample buffer)
                                 ("c4_fused", lambda: measure_beamform(env, WORKLOADS["c4"], 3 * short, 3, 32, "fused", False)),
                                 ("c4_nccl", lambda: measure_beamform(env, WORKLOADS["c4"], short, 3, 32, "nccl", False))):
                    try:
                        sh[name] = compact(fn())
                    except Exception as e:
                        sh[name] = {"error": repr(e)}
                    env.barrier()
                line["sharded"] = sh
    elif kind == "convert_shift":
        line = as_line(env, args, measure_convert_shift(env, w, args.steps, args.warmup, args.buffers))
    elif kind == "channelizer":
        line = as_line(env, args, measure_channelizer(env, w, args.steps, args.warmup, True))
    else:
        line = as_line(env, args, measure_beamform(env, w, args.steps, args.warmup, args.buffers, args.beam_mode, True))
    rank = env.rank
    env.close()
    if rank == 0:
        print(json.dumps(line), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
