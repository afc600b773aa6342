"""ctypes binding of libhzsdrcuda.so -- the harness the tests and bench.py drive the C ABI with.

This is NOT the product's host language (that is Go over cgo, go-sdr_b200/go, with a C++ mirror
in go-sdr_b200/host because no Go toolchain exists in this image); it exists because pytest is
the test runner.  It calls exactly the symbols include/hzsdr_cuda.h declares and nothing else:
there is no CPU fallback, and a missing library or GPU raises immediately.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HZSDR_LIB", os.path.join(_HERE, "..", "lib", "libhzsdrcuda.so"))

FORMAT_C64, FORMAT_U8, FORMAT_I16, FORMAT_I8 = 1, 2, 3, 4
FFT_FORWARD, FFT_BACKWARD = 1, 0

OK = 0
ERR_NO_DEVICE, ERR_CUDA, ERR_INVALID, ERR_DST_TOO_SMALL, ERR_FORMAT_MISMATCH = 1, 2, 3, 4, 5
ERR_FORMAT_UNKNOWN, ERR_CONVERSION_NOT_IMPLEMENTED, ERR_NOMEM, ERR_NCCL, ERR_UNSUPPORTED = 6, 7, 8, 9, 10
ERR_RING_UNDERRUN = 11

NP_DTYPE = {FORMAT_C64: np.complex64, FORMAT_U8: np.uint8, FORMAT_I16: np.int16, FORMAT_I8: np.int8}


class HzsdrError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"hzsdr status {status}: {msg}")
        self.status = status


class NcoState(C.Structure):
    _fields_ = [("sample_rate", C.c_uint32), ("ts", C.c_double)]


class ChainConfig(C.Structure):
    _fields_ = [("src_format", C.c_int), ("sample_rate", C.c_uint32), ("shift_hz", C.c_double),
                ("n_fft", C.c_size_t), ("filter_host", C.c_void_p), ("decimate", C.c_uint32),
                ("decimate_block", C.c_uint32), ("i16_lsb_bits", C.c_int), ("overlap_save_taps", C.c_uint32)]


# every symbol include/hzsdr_cuda.h declares: name -> (restype, argtypes)
_vp, _sz, _i, _u, _d, _f = C.c_void_p, C.c_size_t, C.c_int, C.c_uint, C.c_double, C.c_float
_pvp, _psz, _pi = C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_int)
SYMBOLS = {
    "hzsdr_last_error": (C.c_char_p, []),
    "hzsdr_version": (C.c_char_p, []),
    "hzsdr_format_size": (_i, [_i]),
    "hzsdr_device_count": (_i, [_pi]),
    "hzsdr_ctx_create": (_i, [_i, _pvp]),
    "hzsdr_ctx_destroy": (_i, [_vp]),
    "hzsdr_ctx_sync": (_i, [_vp]),
    "hzsdr_ctx_stream": (_i, [_vp, _pvp]),
    "hzsdr_ctx_info": (_i, [_vp, C.c_char_p, _sz, _pi, _pi, _pi, _psz]),
    "hzsdr_dev_alloc": (_i, [_vp, _sz, _pvp]),
    "hzsdr_dev_free": (_i, [_vp, _vp]),
    "hzsdr_dev_memset": (_i, [_vp, _vp, _i, _sz]),
    "hzsdr_pinned_alloc": (_i, [_sz, _pvp]),
    "hzsdr_pinned_free": (_i, [_vp]),
    "hzsdr_upload": (_i, [_vp, _vp, _vp, _sz]),
    "hzsdr_download": (_i, [_vp, _vp, _vp, _sz]),
    "hzsdr_copy": (_i, [_vp, _vp, _vp, _sz]),
    "hzsdr_convert_to_c64": (_i, [_vp, _i, _vp, _sz, _vp, _sz, _psz]),
    "hzsdr_convert": (_i, [_vp, _i, _vp, _sz, _i, _vp, _sz, _psz]),
    "hzsdr_add_int": (_i, [_vp, _i, _vp, _pvp, _i, _sz]),
    "hzsdr_i16_shift_lsb_to_msb": (_i, [_vp, _vp, _sz, _i]),
    "hzsdr_lookup": (_i, [_vp, _i, _vp, _sz, _i, _vp, _vp, _sz]),
    "hzsdr_shift": (_i, [_vp, _vp, _sz, _d, C.POINTER(NcoState)]),
    "hzsdr_convert_shift": (_i, [_vp, _i, _vp, _sz, _vp, _sz, _d, C.POINTER(NcoState)]),
    "hzsdr_convert_shift_batch": (_i, [_vp, _i, _pvp, _sz, _pvp, _sz, _sz, _d, C.POINTER(NcoState)]),
    "hzsdr_rotate": (_i, [_vp, _vp, _sz, _f, _f]),
    "hzsdr_scale": (_i, [_vp, _vp, _sz, _f]),
    "hzsdr_add": (_i, [_vp, _vp, _pvp, _i, _sz]),
    "hzsdr_decimate": (_i, [_vp, _i, _vp, _sz, _vp, _sz, _u, _sz, _psz]),
    "hzsdr_downsample": (_i, [_vp, _i, _vp, _sz, _vp, _sz, _u, _sz, _psz]),
    "hzsdr_fft_plan_create": (_i, [_vp, _sz, _sz, _i, _pvp]),
    "hzsdr_fft_exec": (_i, [_vp, _vp, _vp, _sz]),
    "hzsdr_fft_plan_destroy": (_i, [_vp]),
    "hzsdr_convolve_freq": (_i, [_vp, _vp, _vp, _vp, _sz, _sz]),
    "hzsdr_fft_convolve": (_i, [_vp, _vp, _vp, _vp, _sz, _sz, _i, _vp]),
    "hzsdr_fftshift_scale": (_i, [_vp, _vp, _sz, _sz, C.c_float]),
    "hzsdr_graft": (_i, [_vp, _vp, _sz, _sz, _vp, _vp]),
    "hzsdr_correlate_peak": (_i, [_vp, _vp, _sz, _sz, C.POINTER(C.c_int32)]),
    "hzsdr_phase_offsets": (_i, [_vp, _vp, _sz, _sz, C.POINTER(C.c_float)]),
    "hzsdr_beamform": (_i, [_vp, _i, _pvp, _i, C.POINTER(C.c_float), _sz, _vp]),
    "hzsdr_beamform_angles_2d": (_i, [_d, _d, C.POINTER(C.c_double), C.POINTER(C.c_double), _i, C.POINTER(C.c_float)]),
    "hzsdr_chain_create": (_i, [_vp, C.POINTER(ChainConfig), _pvp]),
    "hzsdr_chain_destroy": (_i, [_vp]),
    "hzsdr_chain_out_len": (_i, [_vp, _sz, _psz]),
    "hzsdr_chain_exec": (_i, [_vp, _vp, _sz, _vp, _sz, _psz]),
    "hzsdr_chain_exec_batch": (_i, [_vp, _pvp, _sz, _pvp, _sz, _sz, _psz]),
    "hzsdr_chain_exec_host": (_i, [_vp, _vp, _sz, _vp, _sz, _psz]),
    "hzsdr_chain_submit_host": (_i, [_vp, _vp, _sz, _vp, _sz, _psz]),
    "hzsdr_chain_wait_host": (_i, [_vp]),
    "hzsdr_chain_get_ts": (_i, [_vp, C.POINTER(C.c_double)]),
    "hzsdr_chain_set_ts": (_i, [_vp, _d]),
    "hzsdr_chain_submit_ring": (_i, [_vp, _vp, _vp, _sz, _psz]),
    "hzsdr_fir_create": (_i, [_vp, C.POINTER(C.c_float), _sz, _u, _i, _pvp]),
    "hzsdr_fir_destroy": (_i, [_vp]),
    "hzsdr_fir_reset": (_i, [_vp]),
    "hzsdr_fir_exec": (_i, [_vp, _vp, _sz, _vp, _sz, _psz]),
    "hzsdr_polyphase_create": (_i, [_vp, _i, C.c_uint32, _d, C.POINTER(C.c_float), _sz, _u, _i, _pvp]),
    "hzsdr_polyphase_destroy": (_i, [_vp]),
    "hzsdr_polyphase_out_len": (_i, [_vp, _sz, _psz]),
    "hzsdr_polyphase_exec": (_i, [_vp, _vp, _sz, _vp, _sz, _psz]),
    "hzsdr_polyphase_get_ts": (_i, [_vp, C.POINTER(C.c_double)]),
    "hzsdr_polyphase_set_ts": (_i, [_vp, _d]),
    "hzsdr_channelizer_create": (_i, [_vp, C.POINTER(ChainConfig), C.POINTER(C.c_double), _sz, _pvp]),
    "hzsdr_channelizer_destroy": (_i, [_vp]),
    "hzsdr_channelizer_exec": (_i, [_vp, _pvp, _sz, _pvp, _sz, _psz]),
    "hzsdr_channelizer_submit_host": (_i, [_vp, _pvp, _sz, _pvp, _sz, _psz]),
    "hzsdr_beamform_submit_host": (_i, [_vp, _i, _pvp, _i, C.POINTER(C.c_float), _sz, _vp]),
    "hzsdr_ctx_wait_host": (_i, [_vp]),
    "hzsdr_channelizer_get_ts": (_i, [_vp, C.POINTER(C.c_double)]),
    "hzsdr_channelizer_set_ts": (_i, [_vp, C.POINTER(C.c_double)]),
    "hzsdr_ring_create": (_i, [_vp, _i, _sz, _sz, _pvp]),
    "hzsdr_ring_destroy": (_i, [_vp]),
    "hzsdr_ring_write_peek": (_i, [_vp, _pvp]),
    "hzsdr_ring_write_poke": (_i, [_vp, _sz]),
    "hzsdr_ring_read": (_i, [_vp, _pvp, _psz]),
    "hzsdr_ring_read_done": (_i, [_vp]),
    "hzsdr_comm_unique_id": (_i, [_vp]),
    "hzsdr_comm_create": (_i, [_vp, _i, _i, _vp, _pvp]),
    "hzsdr_comm_destroy": (_i, [_vp]),
    "hzsdr_comm_reduce_c64": (_i, [_vp, _vp, _sz, _i]),
    "hzsdr_comm_allreduce_c64": (_i, [_vp, _vp, _sz]),
    "hzsdr_beam_group_create": (_i, [_vp, _i, _i, _sz, _sz, _vp, _pvp]),
    "hzsdr_beam_group_connect": (_i, [_vp, _vp]),
    "hzsdr_beam_group_exec": (_i, [_vp, _i, _pvp, _i, C.POINTER(C.c_float), _vp]),
    "hzsdr_beam_group_exec_batch": (_i, [_vp, _i, _pvp, _i, C.POINTER(C.c_float), _sz, _pvp]),
    "hzsdr_beam_group_join": (_i, [_vp]),
    "hzsdr_beam_group_destroy": (_i, [_vp]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library and bind every declared symbol.  Raises if it is missing: the
    product has no other implementation to fall back to."""
    global _lib
    if _lib is None:
        path = os.path.abspath(LIB_PATH)
        if not os.path.exists(path):
            raise HzsdrError(ERR_NO_DEVICE, f"{path} not built (run go-sdr_b200/build.sh); there is no CPU fallback")
        L = C.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the export is missing
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _check(rc: int):
    if rc != OK:
        raise HzsdrError(rc, load().hzsdr_last_error().decode(errors="replace"))


class DeviceBuffer:
    """A device allocation (the `cuda.SamplesC64` / raw device buffer of the Go package)."""

    def __init__(self, ctx: "Context", nbytes: int):
        self.ctx = ctx
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        _check(load().hzsdr_dev_alloc(ctx.h, self.nbytes, C.byref(p)))
        self.ptr = p.value

    def free(self):
        if getattr(self, "ptr", None):
            load().hzsdr_dev_free(self.ctx.h, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def upload(self, arr: np.ndarray, offset_bytes: int = 0):
        arr = np.ascontiguousarray(arr)
        assert offset_bytes + arr.nbytes <= self.nbytes
        _check(load().hzsdr_upload(self.ctx.h, self.ptr + offset_bytes, arr.ctypes.data, arr.nbytes))
        self.ctx.sync()  # pageable source: make the call safe for temporaries
        return self

    def download(self, dtype, count: int, offset_bytes: int = 0) -> np.ndarray:
        out = np.empty(count, dtype=dtype)
        assert offset_bytes + out.nbytes <= self.nbytes
        _check(load().hzsdr_download(self.ctx.h, out.ctypes.data, self.ptr + offset_bytes, out.nbytes))
        return out


class Context:
    def __init__(self, device: int = 0):
        self.h = None
        p = C.c_void_p()
        _check(load().hzsdr_ctx_create(device, C.byref(p)))
        self.h = p.value
        self.device = device

    def close(self):
        if self.h:
            load().hzsdr_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        _check(load().hzsdr_ctx_sync(self.h))

    @property
    def stream(self) -> int:
        p = C.c_void_p()
        _check(load().hzsdr_ctx_stream(self.h, C.byref(p)))
        return p.value or 0

    def info(self) -> dict:
        name = C.create_string_buffer(256)
        maj, mnr, sms, mem = C.c_int(), C.c_int(), C.c_int(), C.c_size_t()
        _check(load().hzsdr_ctx_info(self.h, name, 256, C.byref(maj), C.byref(mnr), C.byref(sms), C.byref(mem)))
        return {"name": name.value.decode(), "sm": (maj.value, mnr.value), "sm_count": sms.value, "hbm_bytes": mem.value}

    def alloc(self, nbytes: int) -> DeviceBuffer:
        return DeviceBuffer(self, nbytes)

    def to_device(self, arr: np.ndarray) -> DeviceBuffer:
        arr = np.ascontiguousarray(arr)
        return DeviceBuffer(self, max(arr.nbytes, 16)).upload(arr)

    # ---- thin wrappers, device pointers in / out --------------------------------------------
    def convert_to_c64(self, fmt: int, src_ptr: int, src_len: int, dst_ptr: int, dst_len: int) -> int:
        n = C.c_size_t()
        _check(load().hzsdr_convert_to_c64(self.h, fmt, src_ptr, src_len, dst_ptr, dst_len, C.byref(n)))
        return n.value

    def convert(self, src_fmt: int, src_ptr: int, src_len: int, dst_fmt: int, dst_ptr: int, dst_len: int) -> int:
        n = C.c_size_t()
        _check(load().hzsdr_convert(self.h, src_fmt, src_ptr, src_len, dst_fmt, dst_ptr, dst_len, C.byref(n)))
        return n.value

    def add_int(self, fmt: int, dst_ptr: int, src_ptrs, n: int):
        arr = (C.c_void_p * len(src_ptrs))(*src_ptrs)
        _check(load().hzsdr_add_int(self.h, fmt, dst_ptr, arr, len(src_ptrs), n))

    def shift(self, buf_ptr: int, n: int, freq: float, state: NcoState):
        _check(load().hzsdr_shift(self.h, buf_ptr, n, float(freq), C.byref(state)))

    def convert_shift(self, fmt: int, src_ptr: int, n: int, dst_ptr: int, dst_len: int, freq: float, state: NcoState):
        _check(load().hzsdr_convert_shift(self.h, fmt, src_ptr, n, dst_ptr, dst_len, float(freq), C.byref(state)))

    def convert_shift_batch(self, fmt: int, packed, n_each: int, dst_len_each: int, freq: float, state: NcoState):
        """`count` consecutive buffers of the stream in one call (packed = Chain.pack_batch(srcs, dsts))."""
        s, d, count = packed
        _check(load().hzsdr_convert_shift_batch(self.h, fmt, s, n_each, d, dst_len_each, count, float(freq), C.byref(state)))

    def rotate(self, buf_ptr: int, n: int, m: complex):
        m = np.complex64(m)
        _check(load().hzsdr_rotate(self.h, buf_ptr, n, float(m.real), float(m.imag)))

    def scale(self, buf_ptr: int, n: int, r: float):
        _check(load().hzsdr_scale(self.h, buf_ptr, n, float(np.float32(r))))

    def add(self, dst_ptr: int, src_ptrs, n: int):
        arr = (C.c_void_p * len(src_ptrs))(*src_ptrs)
        _check(load().hzsdr_add(self.h, dst_ptr, arr, len(src_ptrs), n))

    def decimate(self, fmt, src_ptr, n, dst_ptr, dst_len, factor, block=0) -> int:
        out = C.c_size_t()
        _check(load().hzsdr_decimate(self.h, fmt, src_ptr, n, dst_ptr, dst_len, factor, block, C.byref(out)))
        return out.value

    def downsample(self, fmt, src_ptr, n, dst_ptr, dst_len, factor, block=0) -> int:
        out = C.c_size_t()
        _check(load().hzsdr_downsample(self.h, fmt, src_ptr, n, dst_ptr, dst_len, factor, block, C.byref(out)))
        return out.value

    def lookup(self, src_fmt, src_ptr, n, table_fmt, table_ptr, dst_ptr, dst_len):
        _check(load().hzsdr_lookup(self.h, src_fmt, src_ptr, n, table_fmt, table_ptr, dst_ptr, dst_len))

    def convolve_freq(self, src_ptr, dst_ptr, filter_ptr, n_fft, n_blocks):
        _check(load().hzsdr_convolve_freq(self.h, src_ptr, dst_ptr, filter_ptr, n_fft, n_blocks))

    # ---- coherent-receiver helpers (rtl/kerberos/internal) ----
    def fftshift_scale(self, data_ptr: int, n: int, batch: int, scale: float):
        _check(load().hzsdr_fftshift_scale(self.h, data_ptr, n, batch, scale))

    def graft(self, iq_ptr: int, n_readers: int, fft_size: int, dst_ptr: int, freq_ptr: int):
        """One pass of GraftReaders' loop (graft.go:96-125)."""
        _check(load().hzsdr_graft(self.h, iq_ptr, n_readers, fft_size, dst_ptr, freq_ptr))

    def cross_correlate(self, dst_ptr: int, a_ptr: int, b_ptr: int, n: int, batch: int, scratch_ptr: int):
        """CrossCorrelater.run (align.go:44-55): IFFT(FFT(a) * conj(FFT(b)))."""
        _check(load().hzsdr_fft_convolve(self.h, dst_ptr, a_ptr, b_ptr, n, batch, 1, scratch_ptr))

    def correlate_peak(self, cc_ptr: int, n: int, batch: int = 1) -> np.ndarray:
        out = np.zeros(batch, dtype=np.int32)
        _check(load().hzsdr_correlate_peak(self.h, cc_ptr, n, batch, out.ctypes.data_as(C.POINTER(C.c_int32))))
        return out

    def phase_offsets(self, bufs_ptr: int, n_chan: int, n: int) -> np.ndarray:
        out = np.zeros(n_chan, dtype=np.complex64)
        _check(load().hzsdr_phase_offsets(self.h, bufs_ptr, n_chan, n, out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def beamform(self, fmt, chan_ptrs, weights: np.ndarray, n: int, dst_ptr: int):
        w = np.ascontiguousarray(weights, dtype=np.complex64)
        arr = (C.c_void_p * len(chan_ptrs))(*chan_ptrs)
        _check(load().hzsdr_beamform(self.h, fmt, arr, len(chan_ptrs), w.ctypes.data_as(C.POINTER(C.c_float)), n, dst_ptr))

    def beamform_submit_host(self, fmt, chan_host_ptrs, weights: np.ndarray, n: int, dst_host_ptr: int):
        """End to end from (pinned) host buffers; wait_host() completes it."""
        w = np.ascontiguousarray(weights, dtype=np.complex64)
        arr = (C.c_void_p * len(chan_host_ptrs))(*chan_host_ptrs)
        _check(load().hzsdr_beamform_submit_host(self.h, fmt, arr, len(chan_host_ptrs), w.ctypes.data_as(C.POINTER(C.c_float)),
                                                 n, dst_host_ptr))

    def wait_host(self):
        _check(load().hzsdr_ctx_wait_host(self.h))


class FftPlan:
    """fft.Plan (fft/fft.go:52-59)."""

    def __init__(self, ctx: Context, iq_len: int, freq_len: int, direction: int):
        self.ctx = ctx
        self.n = iq_len
        self.h = None
        p = C.c_void_p()
        _check(load().hzsdr_fft_plan_create(ctx.h, iq_len, freq_len, direction, C.byref(p)))
        self.h = p.value

    def transform(self, src_ptr: int, dst_ptr: int, batch: int = 1):
        _check(load().hzsdr_fft_exec(self.h, src_ptr, dst_ptr, batch))

    def close(self):
        if self.h:
            load().hzsdr_fft_plan_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Chain:
    """The fused ConvertReader -> ShiftReader -> ConvolutionReader -> DecimateReader."""

    def __init__(self, ctx: Context, src_format: int, sample_rate: int, shift_hz: float, filt: np.ndarray,
                 decimate: int, decimate_block: int = 0, i16_lsb_bits: int = 0, overlap_save_taps: int = 0):
        self.ctx = ctx
        self.h = None
        filt = np.ascontiguousarray(filt, dtype=np.complex64)
        cfg = ChainConfig(src_format, sample_rate, float(shift_hz), filt.size, filt.ctypes.data, decimate,
                          decimate_block, i16_lsb_bits, overlap_save_taps)
        p = C.c_void_p()
        _check(load().hzsdr_chain_create(ctx.h, C.byref(cfg), C.byref(p)))
        self.h = p.value
        self.src_format = src_format

    def out_len(self, n: int) -> int:
        out = C.c_size_t()
        _check(load().hzsdr_chain_out_len(self.h, n, C.byref(out)))
        return out.value

    def exec(self, src_ptr: int, n: int, dst_ptr: int, dst_len: int) -> int:
        out = C.c_size_t()
        _check(load().hzsdr_chain_exec(self.h, src_ptr, n, dst_ptr, dst_len, C.byref(out)))
        return out.value

    @staticmethod
    def pack_batch(src_ptrs, dst_ptrs):
        return (C.c_void_p * len(src_ptrs))(*src_ptrs), (C.c_void_p * len(dst_ptrs))(*dst_ptrs), len(src_ptrs)

    def exec_batch(self, packed, n_each: int, dst_len_each: int) -> int:
        """`count` consecutive buffers of the stream in one call (packed = Chain.pack_batch(srcs, dsts))."""
        s, d, count = packed
        out = C.c_size_t()
        _check(load().hzsdr_chain_exec_batch(self.h, s, n_each, d, dst_len_each, count, C.byref(out)))
        return out.value

    def exec_host(self, src_host_ptr: int, n: int, dst_host_ptr: int, dst_len: int) -> int:
        out = C.c_size_t()
        _check(load().hzsdr_chain_exec_host(self.h, src_host_ptr, n, dst_host_ptr, dst_len, C.byref(out)))
        return out.value

    def submit_host(self, src_host_ptr: int, n: int, dst_host_ptr: int, dst_len: int) -> int:
        out = C.c_size_t()
        _check(load().hzsdr_chain_submit_host(self.h, src_host_ptr, n, dst_host_ptr, dst_len, C.byref(out)))
        return out.value

    def submit_ring(self, ring: "Ring", dst_host_ptr: int, dst_len: int) -> int:
        """Next unread ring slot -> chain -> dst_host (pinned), enqueued; raises HzsdrError(RING_UNDERRUN) when none is pending."""
        out = C.c_size_t()
        _check(load().hzsdr_chain_submit_ring(self.h, ring.h, dst_host_ptr, dst_len, C.byref(out)))
        return out.value

    def wait_host(self):
        _check(load().hzsdr_chain_wait_host(self.h))

    @property
    def ts(self) -> float:
        v = C.c_double()
        _check(load().hzsdr_chain_get_ts(self.h, C.byref(v)))
        return v.value

    @ts.setter
    def ts(self, v: float):
        _check(load().hzsdr_chain_set_ts(self.h, float(v)))

    def close(self):
        if self.h:
            load().hzsdr_chain_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


FIR_AUTO, FIR_OVERLAP_SAVE, FIR_POLYPHASE = 0, 1, 2


class Polyphase:
    """Fused Convert -> Shift -> real-tap FIR -> keep every D-th sample on a raw device stream (extension, see the header)."""

    def __init__(self, ctx: Context, src_format: int, sample_rate: int, shift_hz: float, taps: np.ndarray, decimate: int,
                 i16_lsb_bits: int = 0):
        self.ctx = ctx
        self.h = None
        t = np.ascontiguousarray(taps, dtype=np.float32)
        p = C.c_void_p()
        _check(load().hzsdr_polyphase_create(ctx.h, src_format, sample_rate, float(shift_hz), t.ctypes.data_as(C.POINTER(C.c_float)),
                                             t.size, decimate, i16_lsb_bits, C.byref(p)))
        self.h = p.value

    def out_len(self, n: int) -> int:
        out = C.c_size_t()
        _check(load().hzsdr_polyphase_out_len(self.h, n, C.byref(out)))
        return out.value

    def exec(self, src_ptr: int, n: int, dst_ptr: int, dst_len: int) -> int:
        out = C.c_size_t()
        _check(load().hzsdr_polyphase_exec(self.h, src_ptr, n, dst_ptr, dst_len, C.byref(out)))
        return out.value

    @property
    def ts(self) -> float:
        v = C.c_double()
        _check(load().hzsdr_polyphase_get_ts(self.h, C.byref(v)))
        return v.value

    @ts.setter
    def ts(self, v: float):
        _check(load().hzsdr_polyphase_set_ts(self.h, float(v)))

    def close(self):
        if self.h:
            load().hzsdr_polyphase_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Fir:
    """True linear FIR convolution + decimation of a device c64 stream (extension, see the header)."""

    def __init__(self, ctx: Context, taps: np.ndarray, decimate: int, method: int = FIR_AUTO):
        self.ctx = ctx
        self.h = None
        t = np.ascontiguousarray(taps, dtype=np.complex64)
        p = C.c_void_p()
        _check(load().hzsdr_fir_create(ctx.h, t.ctypes.data_as(C.POINTER(C.c_float)), t.size, decimate, method, C.byref(p)))
        self.h = p.value

    def exec(self, src_ptr: int, n: int, dst_ptr: int, dst_len: int) -> int:
        out = C.c_size_t()
        _check(load().hzsdr_fir_exec(self.h, src_ptr, n, dst_ptr, dst_len, C.byref(out)))
        return out.value

    def reset(self):
        _check(load().hzsdr_fir_reset(self.h))

    def close(self):
        if self.h:
            load().hzsdr_fir_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Channelizer:
    """n independent streams through the fused chain, one launch per buffer set (config 5)."""

    def __init__(self, ctx: Context, src_format: int, sample_rate: int, shifts_hz, filt: np.ndarray, decimate: int,
                 decimate_block: int = 0, i16_lsb_bits: int = 0):
        self.ctx = ctx
        self.h = None
        self.n = len(shifts_hz)
        filt = np.ascontiguousarray(filt, dtype=np.complex64)
        cfg = ChainConfig(src_format, sample_rate, 0.0, filt.size, filt.ctypes.data, decimate, decimate_block, i16_lsb_bits, 0)
        sh = (C.c_double * self.n)(*[float(x) for x in shifts_hz])
        p = C.c_void_p()
        _check(load().hzsdr_channelizer_create(ctx.h, C.byref(cfg), sh, self.n, C.byref(p)))
        self.h = p.value

    def exec(self, src_ptrs, n: int, dst_ptrs, dst_len: int) -> int:
        out = C.c_size_t()
        s = (C.c_void_p * self.n)(*src_ptrs)
        d = (C.c_void_p * self.n)(*dst_ptrs)
        _check(load().hzsdr_channelizer_exec(self.h, s, n, d, dst_len, C.byref(out)))
        return out.value

    def submit_host(self, src_host_ptrs, n: int, dst_host_ptrs, dst_len: int) -> int:
        """End to end from (pinned) host buffers; ctx.wait_host() completes it."""
        out = C.c_size_t()
        s = (C.c_void_p * self.n)(*src_host_ptrs)
        d = (C.c_void_p * self.n)(*dst_host_ptrs)
        _check(load().hzsdr_channelizer_submit_host(self.h, s, n, d, dst_len, C.byref(out)))
        return out.value

    @property
    def ts(self) -> np.ndarray:
        out = np.empty(self.n, dtype=np.float64)
        _check(load().hzsdr_channelizer_get_ts(self.h, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def set_ts(self, ts):
        """Resume every stream at its own carried NCO time (checkpoint / resume, shifter.go:68)."""
        v = np.ascontiguousarray(ts, dtype=np.float64)
        assert v.size == self.n
        _check(load().hzsdr_channelizer_set_ts(self.h, v.ctypes.data_as(C.POINTER(C.c_double))))

    def close(self):
        if self.h:
            load().hzsdr_channelizer_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PinnedBuffer:
    """cudaHostAlloc'd memory viewed as a numpy array (what yikes.Samples does for Go)."""

    def __init__(self, nbytes: int):
        self.ptr = None
        p = C.c_void_p()
        _check(load().hzsdr_pinned_alloc(nbytes, C.byref(p)))
        self.ptr = p.value
        self.nbytes = nbytes

    def view(self, dtype) -> np.ndarray:
        buf = (C.c_uint8 * self.nbytes).from_address(self.ptr)
        return np.frombuffer(buf, dtype=dtype)

    def free(self):
        if self.ptr:
            load().hzsdr_pinned_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Ring:
    def __init__(self, ctx: Context, fmt: int, slots: int, slot_len: int):
        self.ctx, self.fmt, self.slot_len = ctx, fmt, slot_len
        self.h = None
        p = C.c_void_p()
        _check(load().hzsdr_ring_create(ctx.h, fmt, slots, slot_len, C.byref(p)))
        self.h = p.value

    def write(self, arr: np.ndarray):
        arr = np.ascontiguousarray(arr)
        p = C.c_void_p()
        _check(load().hzsdr_ring_write_peek(self.h, C.byref(p)))
        C.memmove(p.value, arr.ctypes.data, arr.nbytes)
        per = 1 if self.fmt == FORMAT_C64 else 2
        _check(load().hzsdr_ring_write_poke(self.h, arr.size // per))

    def write_peek(self) -> int:
        """Host address of the next pinned slot (blocks while a lapped copy out of it is still pending)."""
        p = C.c_void_p()
        _check(load().hzsdr_ring_write_peek(self.h, C.byref(p)))
        return p.value

    def write_poke(self, n_samples: int):
        _check(load().hzsdr_ring_write_poke(self.h, n_samples))

    def read(self):
        p, n = C.c_void_p(), C.c_size_t()
        _check(load().hzsdr_ring_read(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def read_done(self):
        _check(load().hzsdr_ring_read_done(self.h))

    def close(self):
        if self.h:
            load().hzsdr_ring_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Comm:
    """The library's NCCL communicator (multi-GPU Beamform reduce)."""

    ID_BYTES = 128

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(Comm.ID_BYTES)
        _check(load().hzsdr_comm_unique_id(buf))
        return bytes(buf.raw)

    def __init__(self, ctx: Context, nranks: int, rank: int, uid: bytes):
        self.ctx = ctx
        self.h = None
        p = C.c_void_p()
        buf = C.create_string_buffer(uid, Comm.ID_BYTES)
        _check(load().hzsdr_comm_create(ctx.h, nranks, rank, buf, C.byref(p)))
        self.h = p.value

    def reduce_c64(self, buf_ptr: int, n: int, root: int = 0):
        _check(load().hzsdr_comm_reduce_c64(self.h, buf_ptr, n, root))

    def allreduce_c64(self, buf_ptr: int, n: int):
        _check(load().hzsdr_comm_allreduce_c64(self.h, buf_ptr, n))

    def close(self):
        if self.h:
            load().hzsdr_comm_destroy(self.h)
            self.h = None


class BeamGroup:
    """Multi-GPU Beamform with the reduce-scatter fused into the kernel over NVLink peer memory."""

    HANDLE_BYTES = 64

    def __init__(self, ctx: Context, nranks: int, rank: int, n: int, max_batch: int = 1):
        self.ctx = ctx
        self.h = None
        self.nranks, self.rank, self.n, self.max_batch = nranks, rank, n, max_batch
        buf = C.create_string_buffer(BeamGroup.HANDLE_BYTES)
        p = C.c_void_p()
        _check(load().hzsdr_beam_group_create(ctx.h, nranks, rank, n, max_batch, buf, C.byref(p)))
        self.h = p.value
        self.handle = bytes(buf.raw)

    def connect(self, handles):
        """handles: the `handle` of every rank, in rank order."""
        blob = b"".join(handles)
        assert len(blob) == self.nranks * BeamGroup.HANDLE_BYTES
        buf = C.create_string_buffer(blob, len(blob))
        _check(load().hzsdr_beam_group_connect(self.h, buf))

    @staticmethod
    def pack(chan_ptrs, weights: np.ndarray):
        """Marshal the per-call arguments once (hot loops reuse the result with exec_packed)."""
        w = np.ascontiguousarray(weights, dtype=np.complex64)
        arr = (C.c_void_p * max(len(chan_ptrs), 1))(*chan_ptrs)
        return arr, len(chan_ptrs), w, w.ctypes.data_as(C.POINTER(C.c_float))

    def exec_packed(self, fmt: int, packed, dst_slice_ptr: int):
        arr, n, _keep, wp = packed
        rc = load().hzsdr_beam_group_exec(self.h, fmt, arr, n, wp, dst_slice_ptr)
        if rc:
            _check(rc)

    def exec(self, fmt: int, chan_ptrs, weights: np.ndarray, dst_slice_ptr: int):
        self.exec_packed(fmt, self.pack(chan_ptrs, weights), dst_slice_ptr)

    @staticmethod
    def pack_batch(chan_ptrs_per_buffer, weights: np.ndarray, dst_slice_ptrs):
        """Marshal one exchange over len(chan_ptrs_per_buffer) buffers (buffer-major pointer table)."""
        nbuf, nchan = len(chan_ptrs_per_buffer), len(chan_ptrs_per_buffer[0])
        flat = [p for row in chan_ptrs_per_buffer for p in row]
        w = np.ascontiguousarray(weights, dtype=np.complex64)
        arr = (C.c_void_p * max(len(flat), 1))(*flat)
        dst = (C.c_void_p * nbuf)(*dst_slice_ptrs)
        return arr, nchan, w, w.ctypes.data_as(C.POINTER(C.c_float)), nbuf, dst

    def exec_batch_packed(self, fmt: int, packed):
        arr, nchan, _keep, wp, nbuf, dst = packed
        rc = load().hzsdr_beam_group_exec_batch(self.h, fmt, arr, nchan, wp, nbuf, dst)
        if rc:
            _check(rc)

    def exec_batch(self, fmt: int, chan_ptrs_per_buffer, weights: np.ndarray, dst_slice_ptrs):
        self.exec_batch_packed(fmt, self.pack_batch(chan_ptrs_per_buffer, weights, dst_slice_ptrs))

    def join(self):
        _check(load().hzsdr_beam_group_join(self.h))

    def close(self):
        if self.h:
            load().hzsdr_beam_group_destroy(self.h)
            self.h = None


def beamform_angles_2d(frequency_hz: float, angle_deg: float, center, antennas):
    """stream.BeamformAngles2D (beamform.go:57-107); host math inside the library, no GPU needed."""
    if len(antennas) == 0:
        return None
    ants = np.ascontiguousarray(antennas, dtype=np.float64).reshape(-1)
    ctr = (C.c_double * 2)(float(center[0]), float(center[1]))
    out = np.empty(len(antennas), dtype=np.complex64)
    _check(load().hzsdr_beamform_angles_2d(float(frequency_hz), float(angle_deg), ctr,
                                           ants.ctypes.data_as(C.POINTER(C.c_double)), len(antennas),
                                           out.ctypes.data_as(C.POINTER(C.c_float))))
    return out


def beamform_angles(frequency_hz: float, angle_deg: float, distances):
    """stream.BeamformAngles (beamform.go:115-128)."""
    if len(distances) == 0:
        return None
    ants = [(float(d), 0.0) for d in distances]
    return beamform_angles_2d(frequency_hz, angle_deg, ants[0], ants)
