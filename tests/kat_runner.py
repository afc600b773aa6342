"""Replay of the reference's known-answer tests (tests/golden/reference_kats.json) against any
implementation exposing the oracle's function names.  Used with the numpy oracle on CPU
(tests/test_oracle.py) and with the CUDA library on the GPU (tests/test_gpu_kats.py), so both
are pinned by the same reference-held vectors."""
from __future__ import annotations

import json
import math
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "reference_kats.json")) as fh:
    KATS = json.load(fh)

FMT = {"u8": (2, np.uint8), "i8": (4, np.int8), "i16": (3, np.int16)}


def in_epsilon(expected, actual, eps):
    """testify's assert.InEpsilon: |expected-actual| / |expected| <= eps."""
    expected = np.asarray(expected, dtype=np.float64)
    actual = np.asarray(actual, dtype=np.float64)
    assert np.all(np.abs(expected - actual) <= eps * np.abs(expected)), (expected, actual, eps)


def const_c64(n, pair):
    return np.full(n, complex(pair[0], pair[1]), dtype=np.complex64)


def run_convert(impl):
    for k in KATS["convert"]:
        fmt, dt = FMT[k["fmt"]]
        raw = np.tile(np.array(k["in"], dtype=dt), k["repeat"])
        out = impl.convert_to_c64(raw, fmt)
        assert out.dtype == np.complex64 and out.shape == (k["repeat"],), k["cite"]
        off = 1.0 if k.get("plus_one") else 0.0
        in_epsilon(off + k["out"][0], off + out.real, k["tol"])
        in_epsilon(off + k["out"][1], off + out.imag, k["tol"])
    k = KATS["convert_u8_midscale"]
    out = impl.convert_to_c64(np.array(k["in"], dtype=np.uint8), 2)
    in_epsilon(1.0, float(out.real[0]) + float(out.real[1]) + 1.0, k["tol"])
    in_epsilon(1.0, float(out.imag[0]) + float(out.imag[1]) + 1.0, k["tol"])


def run_scale(impl):
    for k in KATS["scale"]:
        buf = const_c64(k["n"], k["in"])
        m = k.get("scaled_n", k["n"])
        out = buf.copy()
        out[:m] = impl.scale(buf[:m], k["r"])
        assert np.array_equal(out[:m], const_c64(m, k["out"])), k["cite"]
        assert np.array_equal(out[m:], buf[m:]), k["cite"]


def run_rotate(impl):
    for k in KATS["rotate"]:
        out = impl.rotate(const_c64(k["n"], k["in"]), complex(*k["m"]))
        assert np.array_equal(out, const_c64(k["n"], k["out"])), k["cite"]
    k = KATS["rotate_cw"]
    import go_sdr_oracle as O  # input generator only (testutils/cw.go)
    cw0 = O.cw(k["n"], k["freq"], k["sample_rate"], 0.0)
    cw90 = O.cw(k["n"], k["freq"], k["sample_rate"], k["phase"])
    out = impl.rotate(cw90, complex(*k["m"]))
    in_epsilon(1.0 + cw0.real.astype(np.float64), 1.0 + out.real.astype(np.float64), k["tol"])
    in_epsilon(1.0 + cw0.imag.astype(np.float64), 1.0 + out.imag.astype(np.float64), k["tol"])


def run_add(impl):
    for k in KATS["add"]:
        bufs = [const_c64(k["n"], p) for p in k["in"]]
        out = impl.add(*bufs)
        assert np.array_equal(out, const_c64(k["n"], k["out"])), k["cite"]


def run_shift_roundtrip(impl):
    k = KATS["shift_roundtrip"]
    import go_sdr_oracle as O
    cw = O.cw(k["n"], k["cw_freq"], k["sample_rate"], 0.0)
    up, _ = impl.shift_buffer(cw, k["shift"], k["sample_rate"], 0.0)
    down, _ = impl.shift_buffer(up, -k["shift"], k["sample_rate"], 0.0)
    in_epsilon(1.0 + cw.real.astype(np.float64), 1.0 + down.real.astype(np.float64), k["tol"])
    in_epsilon(1.0 + cw.imag.astype(np.float64), 1.0 + down.imag.astype(np.float64), k["tol"])


def run_decimate(impl):
    for k in KATS["decimate"]:
        n = k["n"]
        if "pattern_mod" in k:
            z = (np.arange(n) % k["pattern_mod"]).astype(np.float32)
        else:
            z = np.zeros(n, dtype=np.float32)
        x = (z + 1j * z).astype(np.complex64)
        out = impl.decimate_reader(x, k["factor"])
        assert out.shape[0] == k["out_n"], k["cite"]
        if "out_value" in k:
            assert np.all(out == k["out_value"]), k["cite"]


def run_downsample(impl):
    for k in KATS["downsample"]:
        n = k["n"]
        z = (np.arange(n) % k.get("pattern_mod", 1)).astype(np.float32)
        x = (z + 1j * z).astype(np.complex64)
        out = impl.downsample_reader(x, k["factor"])
        assert out.shape[0] == k["out_n"], k["cite"]
        if "out" in k:
            assert np.array_equal(out, const_c64(k["out_n"], k["out"])), k["cite"]


def run_beamform_angles(impl):
    for k in KATS["beamform_angles"]:
        if k["kind"] == "1d":
            w = impl.beamform_angles(k["freq"], k["angle"], k["distances"])
        else:
            w = impl.beamform_angles_2d(k["freq"], k["angle"], k["center"], k["antennas"])
        w = np.asarray(w, dtype=np.complex64)
        tol = k["tol"]
        for i in k.get("expect_unity", []):
            in_epsilon(1.0, float(w[i].real), tol)
            in_epsilon(1.0, 1.0 + float(w[i].imag), tol)
        for i in k.get("expect_real_unity", []):
            in_epsilon(1.0, float(w[i].real), tol)
        if "expect_phase_plus_2pi" in k:
            e = k["expect_phase_plus_2pi"]
            ph = np.angle(np.conj(np.complex128(w[e["index"]])))
            in_epsilon(e["deg"] * math.pi / 180, 2 * math.pi + ph, tol)
        for i, ph in k.get("expect_conj_phase", []):
            got = float(np.angle(np.conj(np.complex128(w[i]))))
            if abs(abs(ph) - math.pi) < 1e-12:  # +pi and -pi are the same angle; cmplx.Phase picks by sign of imag
                in_epsilon(math.pi, abs(got), tol)
            else:
                in_epsilon(ph, got, tol)
        for i, v in k.get("expect_one_plus_conj_phase", []):
            in_epsilon(v, 1.0 + float(np.angle(np.conj(np.complex128(w[i])))), tol)


def run_fft_contract(impl):
    """testutils/fft.go:54-125: peak-bin location of a forward FFT, and backward-then-forward
    of a single bin returns to that bin."""
    k = KATS["fft_contract"]
    import go_sdr_oracle as O
    for f, idx in k["peaks"]:
        x = O.cw(k["n"], f, k["sample_rate"], 0.0)
        X = impl.fft_forward(x)
        assert int(np.argmax(np.abs(X.astype(np.complex128)))) == idx
    for b in (5, 10, 127, 522, 242, 415, 825):
        F = np.zeros(k["n"], dtype=np.complex64)
        F[b] = 1 + 1j
        x = impl.fft_backward(F)
        F2 = impl.fft_forward(x)
        assert int(np.argmax(np.abs(F2.astype(np.complex128)))) == b


def run_convert_matrix(impl):
    """iq_c64_test.go:38-108 and the integer <-> integer tests: exact values."""
    for kat in KATS["convert_from_c64"]:
        fmt, dt = FMT[kat["fmt"]]
        x = np.array([complex(*kat["in"])], dtype=np.complex64)
        out = np.asarray(impl.convert_from_c64(x, fmt)).reshape(-1)
        assert out.dtype == dt and list(out) == kat["out"], (kat["cite"], out)
    for kat in KATS["convert_int"]:
        sf, sdt = FMT[kat["src"]]
        df, ddt = FMT[kat["dst"]]
        out = np.asarray(impl.convert_int(np.array(kat["in"], dtype=sdt), sf, df)).reshape(-1)
        assert out.dtype == ddt and list(out) == kat["out"], (kat["cite"], out)
