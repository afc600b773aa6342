"""Synthetic inputs for the benchmarks and tests: CW + complex Gaussian noise
quantised as an ADC would present it, windowed-sinc lowpass taps and their frequency-domain form.
Plain numpy, independent of the oracle and of the CUDA library."""
from __future__ import annotations

import math

import numpy as np

FORMAT_C64, FORMAT_U8, FORMAT_I16, FORMAT_I8 = 1, 2, 3, 4
TAU = math.pi * 2


def synth_raw(fmt: int, n: int, sample_rate: int, f0: float, seed: int, amp: float = 0.5, sigma: float = 0.05,
              phase: float = 0.0) -> np.ndarray:
    """Interleaved raw integer vector (2n,) of a CW at f0 plus noise, clipped to [-1, 1)."""
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64) / float(sample_rate)
    a = TAU * f0 * t + phase
    x = np.empty(2 * n, dtype=np.float64)
    x[0::2] = amp * np.cos(a) + sigma * rng.standard_normal(n)
    x[1::2] = amp * np.sin(a) + sigma * rng.standard_normal(n)
    np.clip(x, -1.0, np.nextafter(1.0, 0.0), out=x)
    if fmt == FORMAT_U8:
        return np.clip(np.rint(127.5 * x + 127.5), 0, 255).astype(np.uint8)
    if fmt == FORMAT_I8:
        return np.clip(np.rint(128.0 * x), -128, 127).astype(np.int8)
    if fmt == FORMAT_I16:
        return np.clip(np.rint(32767.0 * x), -32768, 32767).astype(np.int16)
    raise ValueError(f"unknown raw format {fmt}")


def lowpass_taps(ntaps: int, cutoff: float) -> np.ndarray:
    """Hamming-windowed sinc, `cutoff` in cycles/sample (one-sided), unity DC gain."""
    k = np.arange(ntaps, dtype=np.float64) - (ntaps - 1) / 2.0
    h = 2 * cutoff * np.sinc(2 * cutoff * k) * np.hamming(ntaps)
    return (h / h.sum()).astype(np.float32)


def filter_freq(taps: np.ndarray, nfft: int) -> np.ndarray:
    """The frequency-domain `filter` argument of stream.ConvolutionReader: FFT_N of the zero-padded
    taps, pre-scaled by 1/N because both transforms are unnormalised."""
    h = np.zeros(nfft, dtype=np.complex128)
    h[: len(taps)] = taps
    return (np.fft.fft(h) / nfft).astype(np.complex64)
