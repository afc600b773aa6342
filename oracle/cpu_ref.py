"""ctypes loader for oracle/_build/libcpu_ref.so (the oracle's C twin).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libcpu_ref.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "cpu_ref.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, sz, u, d, f, i = C.c_void_p, C.c_size_t, C.c_uint, C.c_double, C.c_float, C.c_int
        pd = C.POINTER(C.c_double)
        L.ref_convert_u8_c64.argtypes = [vp, vp, sz]
        L.ref_convert_i8_c64.argtypes = [vp, vp, sz]
        L.ref_convert_i16_c64.argtypes = [vp, vp, sz]
        L.ref_shift_buffer.argtypes = [vp, sz, d, u, pd]
        L.ref_shift_ts.argtypes = [vp, sz, u, pd]
        L.ref_rotate.argtypes = [vp, sz, f, f]
        L.ref_scale.argtypes = [vp, sz, f]
        L.ref_add.argtypes = [vp, vp, vp, sz]
        L.ref_decimate_c64.argtypes = [vp, sz, vp, u]
        L.ref_decimate_c64.restype = sz
        L.ref_downsample_c64.argtypes = [vp, sz, vp, u]
        L.ref_downsample_c64.restype = sz
        L.ref_plan_create.argtypes = [sz]
        L.ref_plan_create.restype = vp
        L.ref_plan_destroy.argtypes = [vp]
        L.ref_fft.argtypes = [vp, vp, i]
        L.ref_convolution_blocks.argtypes = [vp, vp, sz, vp]
        L.ref_chain.argtypes = [vp, i, sz, u, d, vp, vp, u, pd, vp, vp]
        L.ref_chain.restype = sz
        L.ref_beamform_u8.argtypes = [vp, sz, vp, sz, vp, vp]
        for name in ("ref_convert_u8_c64", "ref_convert_i8_c64", "ref_convert_i16_c64", "ref_shift_buffer",
                     "ref_shift_ts", "ref_rotate", "ref_scale", "ref_add", "ref_plan_destroy", "ref_fft",
                     "ref_convolution_blocks", "ref_beamform_u8"):
            getattr(L, name).restype = None
        _lib = L
    return _lib


def _p(a: np.ndarray) -> int:
    return a.ctypes.data


def convert_to_c64(raw: np.ndarray, fmt: int) -> np.ndarray:
    raw = np.ascontiguousarray(raw)
    n = raw.size // 2
    out = np.empty(n, dtype=np.complex64)
    fn = {2: lib().ref_convert_u8_c64, 4: lib().ref_convert_i8_c64, 3: lib().ref_convert_i16_c64}[fmt]
    fn(_p(raw), _p(out), n)
    return out


def shift_buffer(buf: np.ndarray, freq: float, sample_rate: int, ts0: float = 0.0):
    out = np.array(buf, dtype=np.complex64, copy=True)
    ts = C.c_double(ts0)
    lib().ref_shift_buffer(_p(out), out.size, float(freq), int(sample_rate), C.byref(ts))
    return out, ts.value


def shift_ts(sample_rate: int, n: int, ts0: float = 0.0, want_array: bool = True):
    out = np.empty(n, dtype=np.float64) if want_array else None
    ts = C.c_double(ts0)
    lib().ref_shift_ts(_p(out) if want_array else None, n, int(sample_rate), C.byref(ts))
    return out, ts.value


def rotate(buf, m: complex):
    out = np.array(buf, dtype=np.complex64, copy=True)
    m = np.complex64(m)
    lib().ref_rotate(_p(out), out.size, float(m.real), float(m.imag))
    return out


def scale(buf, r: float):
    out = np.array(buf, dtype=np.complex64, copy=True)
    lib().ref_scale(_p(out), out.size, float(np.float32(r)))
    return out


def add2(a, b):
    a = np.ascontiguousarray(a, dtype=np.complex64)
    b = np.ascontiguousarray(b, dtype=np.complex64)
    c = np.empty_like(a)
    lib().ref_add(_p(a), _p(b), _p(c), a.size)
    return c


class Plan:
    def __init__(self, n: int):
        self.n = n
        self.h = lib().ref_plan_create(n)
        if not self.h:
            raise ValueError("ref_plan_create: n must be a power of two")

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_plan_destroy(self.h)
            self.h = None

    def fft(self, x, direction: int):
        out = np.array(x, dtype=np.complex64, copy=True)
        lib().ref_fft(self.h, _p(out), direction)
        return out


def chain(raw, fmt, sample_rate, shift_hz, filt, decim, ts0=0.0, plan: Plan | None = None,
          scratch=None, out=None):
    raw = np.ascontiguousarray(raw)
    n = raw.size // 2
    filt = np.ascontiguousarray(filt, dtype=np.complex64)
    plan = plan or Plan(filt.size)
    scratch = np.empty(n, dtype=np.complex64) if scratch is None else scratch
    out = np.empty(n // decim + 1, dtype=np.complex64) if out is None else out
    ts = C.c_double(ts0)
    w = lib().ref_chain(_p(raw), fmt, n, int(sample_rate), float(shift_hz), plan.h, _p(filt), int(decim),
                        C.byref(ts), _p(scratch), _p(out))
    return out[:w], ts.value


def beamform_u8(chans, weights):
    chans = [np.ascontiguousarray(c, dtype=np.uint8) for c in chans]
    n = chans[0].size // 2
    ptrs = (C.c_void_p * len(chans))(*[_p(c) for c in chans])
    w = np.ascontiguousarray(weights, dtype=np.complex64)
    tmp = np.empty(n, dtype=np.complex64)
    out = np.empty(n, dtype=np.complex64)
    lib().ref_beamform_u8(ptrs, len(chans), _p(w), n, _p(tmp), _p(out))
    return out
