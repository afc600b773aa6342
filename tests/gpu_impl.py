"""Adapter: the oracle's function names on top of libhzsdrcuda.so (host numpy in, host numpy out),
so the reference's KATs and the parity tests run unchanged against the CUDA path."""
from __future__ import annotations

import numpy as np

import hzsdr as H

RAW_DTYPE = {H.FORMAT_U8: np.uint8, H.FORMAT_I8: np.int8, H.FORMAT_I16: np.int16}


class GpuImpl:
    def __init__(self, ctx: H.Context | None = None):
        self.ctx = ctx or H.Context(0)

    # ---- helpers ----
    def _up(self, arr):
        return self.ctx.to_device(np.ascontiguousarray(arr))

    # ---- K1 ----
    def convert_to_c64(self, raw, fmt, dst_len=None):
        raw = np.ascontiguousarray(raw)
        n = raw.size if fmt == H.FORMAT_C64 else raw.size // 2
        dst_len = n if dst_len is None else dst_len
        src = self._up(raw)
        dst = self.ctx.alloc(max(dst_len, 1) * 8)
        got = self.ctx.convert_to_c64(fmt, src.ptr, n, dst.ptr, dst_len)
        return dst.download(np.complex64, got)

    def convert(self, data, src_fmt, dst_fmt, dst_len=None):
        data = np.ascontiguousarray(data)
        n = data.size if src_fmt == H.FORMAT_C64 else data.size // 2
        dst_len = n if dst_len is None else dst_len
        src = self._up(data)
        size = {H.FORMAT_C64: 8, H.FORMAT_U8: 2, H.FORMAT_I8: 2, H.FORMAT_I16: 4}.get(dst_fmt, 8)
        dst = self.ctx.alloc(max(dst_len, 1) * size)
        got = self.ctx.convert(src_fmt, src.ptr, n, dst_fmt, dst.ptr, dst_len)
        if dst_fmt == H.FORMAT_C64:
            return dst.download(np.complex64, got)
        return dst.download(H.NP_DTYPE[dst_fmt], 2 * got).reshape(-1, 2)

    def convert_from_c64(self, buf, fmt):
        return self.convert(buf, H.FORMAT_C64, fmt)

    def convert_int(self, raw, src, dst):
        return self.convert(raw, src, dst).reshape(-1)

    def add_int(self, *bufs):
        fmt = {np.dtype(np.int8): H.FORMAT_I8, np.dtype(np.int16): H.FORMAT_I16}[np.asarray(bufs[0]).dtype]
        n = np.asarray(bufs[0]).size // 2
        ds = [self._up(np.ascontiguousarray(b)) for b in bufs]
        out = self.ctx.alloc(max(n, 1) * (2 if fmt == H.FORMAT_I8 else 4))
        self.ctx.add_int(fmt, out.ptr, [d.ptr for d in ds], n)
        return out.download(np.asarray(bufs[0]).dtype, 2 * n)

    def multiply_lut(self, raw, m, fmt):
        """stream.Multiply on a raw u8 / i8 stream (stream/multiply.go:91-251): the table is built on
        the GPU by Convert -> Multiply -> Convert exactly as the reference builds it, reads are
        hzsdr_lookup.  The u8 reader's x0*255+x1 index (and its collisions) is reproduced by
        re-indexing the reference's 65535-entry table into the library's little-endian one."""
        if fmt == H.FORMAT_I8:
            ident = np.arange(65536, dtype=np.uint16).view(np.int8)  # LookupTableIdentityI8
            c = self.rotate(self.convert_to_c64(ident, H.FORMAT_I8), m)
            tab = self.convert_from_c64(c, H.FORMAT_I8)
        else:
            ubuf = np.zeros((65535, 2), dtype=np.uint8)
            rv, iv = np.meshgrid(np.arange(256), np.arange(257), indexing="ij")
            for r_, i_ in zip(rv.reshape(-1), iv.reshape(-1)):  # the reference's loop order decides collisions
                ubuf[(r_ & 0xff) * 255 + (i_ & 0xff)] = (r_ & 0xff, i_ & 0xff)
            c = self.rotate(self.convert_to_c64(ubuf.reshape(-1), H.FORMAT_U8), m)
            t65535 = self.convert_from_c64(c, H.FORMAT_U8)
            x0 = np.arange(65536) & 0xff
            x1 = np.arange(65536) >> 8
            tab = t65535[x0 * 255 + x1]  # little-endian pair index -> the reference's index
        return self.lookup(tab, raw, src_fmt=fmt, table_fmt=fmt)

    def lookup(self, table, raw, src_fmt=H.FORMAT_U8, table_fmt=H.FORMAT_C64):
        raw = np.ascontiguousarray(raw)
        n = raw.size // 2
        t = self._up(table)
        s = self._up(raw)
        size = {H.FORMAT_C64: 8, H.FORMAT_U8: 2, H.FORMAT_I8: 2, H.FORMAT_I16: 4}[table_fmt]
        d = self.ctx.alloc(max(n, 1) * size)
        self.ctx.lookup(src_fmt, s.ptr, n, table_fmt, t.ptr, d.ptr, n)
        if table_fmt == H.FORMAT_C64:
            return d.download(np.complex64, n)
        return d.download(H.NP_DTYPE[table_fmt], 2 * n).reshape(-1, 2)

    # ---- K2 ----
    def shift_buffer(self, buf, freq, sample_rate, ts0=0.0):
        buf = np.ascontiguousarray(buf, dtype=np.complex64)
        d = self._up(buf)
        st = H.NcoState(int(sample_rate), float(ts0))
        self.ctx.shift(d.ptr, buf.size, freq, st)
        return d.download(np.complex64, buf.size), st.ts

    def convert_shift(self, raw, fmt, freq, sample_rate, ts0=0.0):
        raw = np.ascontiguousarray(raw)
        n = raw.size // 2
        s = self._up(raw)
        d = self.ctx.alloc(max(n, 1) * 8)
        st = H.NcoState(int(sample_rate), float(ts0))
        self.ctx.convert_shift(fmt, s.ptr, n, d.ptr, n, freq, st)
        return d.download(np.complex64, n), st.ts

    # ---- K3/K4/K5 ----
    def rotate(self, buf, m):
        buf = np.ascontiguousarray(buf, dtype=np.complex64)
        if np.complex64(m) == np.complex64(1):  # stream/multiply.go:59-62: the reader skips m == 1
            return buf.copy()
        d = self._up(buf)
        self.ctx.rotate(d.ptr, buf.size, m)
        return d.download(np.complex64, buf.size)

    def scale(self, buf, r):
        buf = np.ascontiguousarray(buf, dtype=np.complex64)
        d = self._up(buf)
        self.ctx.scale(d.ptr, buf.size, r)
        return d.download(np.complex64, buf.size)

    def add(self, *bufs):
        n = bufs[0].size
        ds = [self._up(np.ascontiguousarray(b, dtype=np.complex64)) for b in bufs]
        out = self.ctx.alloc(max(n, 1) * 8)
        self.ctx.add(out.ptr, [d.ptr for d in ds], n)
        return out.download(np.complex64, n)

    # ---- K7 ----
    def decimate_reader(self, stream, factor, block=32768, fmt=H.FORMAT_C64, dst_len=None):
        stream = np.ascontiguousarray(stream)
        n = stream.size if fmt == H.FORMAT_C64 else stream.size // 2
        s = self._up(stream)
        cap = n if dst_len is None else dst_len
        size = 8 if fmt == H.FORMAT_C64 else (2 if fmt == H.FORMAT_U8 else 4)
        d = self.ctx.alloc(max(cap, 1) * size)
        got = self.ctx.decimate(fmt, s.ptr, n, d.ptr, cap, factor, block)
        if fmt == H.FORMAT_C64:
            return d.download(np.complex64, got)
        return d.download(H.NP_DTYPE[fmt], 2 * got).reshape(-1, 2)

    def decimate_buffer(self, frm, factor, to_len=None, fmt=H.FORMAT_C64):
        return self.decimate_reader(frm, factor, block=0, fmt=fmt, dst_len=to_len)

    def downsample_reader(self, stream, factor, fmt=H.FORMAT_C64, block=32768):
        stream = np.ascontiguousarray(stream)
        n = stream.size if fmt == H.FORMAT_C64 else stream.size // 2
        s = self._up(stream)
        d = self.ctx.alloc(max(n, 1) * 8)
        got = self.ctx.downsample(fmt, s.ptr, n, d.ptr, n, factor, block)
        return d.download(np.complex64, got)

    def downsample_buffer(self, frm, factor, fmt=H.FORMAT_C64):
        return self.downsample_reader(frm, factor, fmt=fmt, block=0)

    # ---- K6 ----
    def _fft(self, x, direction):
        x = np.ascontiguousarray(x, dtype=np.complex64)
        n = x.shape[-1]
        batch = x.size // n
        plan = H.FftPlan(self.ctx, n, n, direction)
        s = self._up(x)
        d = self.ctx.alloc(x.nbytes)
        plan.transform(s.ptr, d.ptr, batch)
        out = d.download(np.complex64, x.size).reshape(x.shape)
        plan.close()
        return out

    def fft_forward(self, x):
        return self._fft(x, H.FFT_FORWARD)

    def fft_backward(self, x):
        return self._fft(x, H.FFT_BACKWARD)

    def convolution_reader(self, stream, filt):
        stream = np.ascontiguousarray(stream, dtype=np.complex64)
        filt = np.ascontiguousarray(filt, dtype=np.complex64)
        n = filt.size
        nblk = stream.size // n
        s = self._up(stream)
        f = self._up(filt)
        d = self.ctx.alloc(max(nblk * n, 1) * 8)
        self.ctx.convolve_freq(s.ptr, d.ptr, f.ptr, n, nblk)
        return d.download(np.complex64, nblk * n)

    # ---- fused chain ----
    def chain(self, raw, fmt, sample_rate, shift_hz, filt, decim, ts0=0.0, host_path=False, lsb_bits=0):
        raw = np.ascontiguousarray(raw)
        n = raw.size // 2
        ch = H.Chain(self.ctx, fmt, sample_rate, shift_hz, filt, decim, i16_lsb_bits=lsb_bits)
        ch.ts = ts0
        total = ch.out_len(n)
        if host_path:
            out = np.empty(max(total, 1), dtype=np.complex64)
            got = ch.exec_host(raw.ctypes.data, n, out.ctypes.data, total)
            res = out[:got].copy()
        else:
            s = self._up(raw)
            d = self.ctx.alloc(max(total, 1) * 8)
            got = ch.exec(s.ptr, n, d.ptr, total)
            res = d.download(np.complex64, got)
        ts = ch.ts
        ch.close()
        return res, ts

    # ---- K8 ----
    def beamform(self, channels, fmt, weights):
        chans = [self._up(np.ascontiguousarray(c)) for c in channels]
        n = np.asarray(channels[0]).size // 2
        out = self.ctx.alloc(max(n, 1) * 8)
        self.ctx.beamform(fmt, [c.ptr for c in chans], weights, n, out.ptr)
        return out.download(np.complex64, n)

    beamform_angles = staticmethod(H.beamform_angles)
    beamform_angles_2d = staticmethod(H.beamform_angles_2d)
