/*
 * cpu_ref.c -- C restatement of hz.tools/sdr's IQ sample chain.  TEST INFRASTRUCTURE ONLY.
 *
 * This is the oracle's compiled twin: it exists (a) to pin the numpy oracle and the NCO
 * segment table against the reference's literal serial loops at sizes a Python loop cannot
 * reach, and (b) to be timed as `cpu_baseline` / `--impl reference` in bench.py (kind
 * "port": the reference is Go + Go-assembler and no Go toolchain exists in this image, so
 * the reference itself cannot be compiled -- see DESIGN.md).  Only tests/, smoke() and
 * bench.py's CPU legs may load it.  The product (libhzsdrcuda.so) never links or calls it.
 *
 * Loop structure and vector width follow the reference: 128-bit SSE, two samples per
 * iteration, for u8->c64 / scale / add (iq_u8_amd64.s:71-89, internal/simd/
 * mult_simd_amd64.s:47-54, add_simd_amd64.s:58-70); scalar loops everywhere else.
 * Build: gcc -O2 -msse4.1 -fno-fast-math (oracle/Makefile).  All paths cited are relative
 * to the hztools/go-sdr tree.
 *
 * FFT: the reference has NO in-tree FFT (fft/fft.go:45-59 is an interface; go.mod names no
 * planner).  ref_fft below is a plain radix-2 stand-in for the user-supplied Planner with
 * the convention the project adopts (forward e^{-2 pi i kn/N}, backward e^{+...}, both
 * unnormalised).  PARITY UNPINNED for FFT/convolution values.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <smmintrin.h>

#define REF_TAU (M_PI * 2.0)

/* ---- a2: SamplesU8.ToC64, iq_u8.go:111-121 == iq_u8_amd64.s:27-90 ------------------- */
void ref_convert_u8_c64(const uint8_t *src, float *dst, size_t n) {
    /* the asm builds 127.5 as 255/2 (iq_u8_amd64.s:30-40) and does SUBPS then DIVPS on
     * four floats = two samples per iteration; an odd head sample is done by the scalar
     * Go formula (iq_u8_amd64.go:34-37). */
    const __m128 half = _mm_set1_ps(127.5f);
    size_t i = 0;
    if (n & 1) {
        dst[0] = ((float)src[0] - 127.5f) / 127.5f;
        dst[1] = ((float)src[1] - 127.5f) / 127.5f;
        i = 1;
    }
    for (; i < n; i += 2) {
        int32_t w;
        memcpy(&w, src + 2 * i, 4);
        __m128i b = _mm_cvtepu8_epi32(_mm_cvtsi32_si128(w)); /* PMOVZXBD */
        __m128 f = _mm_cvtepi32_ps(b);                       /* VCVTDQ2PS */
        f = _mm_div_ps(_mm_sub_ps(f, half), half);           /* SUBPS, DIVPS */
        _mm_storeu_ps(dst + 2 * i, f);
    }
}

/* ---- a3: SamplesI8.ToC64, iq_i8.go:107-119 ----------------------------------------- */
void ref_convert_i8_c64(const int8_t *src, float *dst, size_t n) {
    for (size_t i = 0; i < n; i++) {
        dst[2 * i] = ((float)src[2 * i]) / 128;
        dst[2 * i + 1] = ((float)src[2 * i + 1]) / 128;
    }
}

/* ---- a4: SamplesI16.ToC64, iq_i16.go:141-145 --------------------------------------- */
void ref_convert_i16_c64(const int16_t *src, float *dst, size_t n) {
    for (size_t i = 0; i < n; i++) {
        dst[2 * i] = (float)src[2 * i] / 32767;
        dst[2 * i + 1] = (float)src[2 * i + 1] / 32767;
    }
}

/* Go complex64 multiply: operands widened to float64, narrowed per component. */
static inline void go_c64_mul(float ar, float ai, float br, float bi, float *re, float *im) {
    *re = (float)((double)ar * (double)br - (double)ai * (double)bi);
    *im = (float)((double)ar * (double)bi + (double)ai * (double)br);
}

/* ---- a6: ShiftBuffer, stream/shifter.go:66-85 -------------------------------------- */
void ref_shift_buffer(float *buf, size_t n, double freq, unsigned sample_rate, double *ts_io) {
    double ts = *ts_io;
    const double inc = 1.0 / (double)sample_rate;
    const double tau = REF_TAU;
    for (size_t j = 0; j < n; j++) {
        ts += inc;
        if (ts > tau) ts -= tau;
        double s, c;
        sincos(tau * freq * ts, &s, &c);
        go_c64_mul(buf[2 * j], buf[2 * j + 1], (float)c, (float)s, &buf[2 * j], &buf[2 * j + 1]);
    }
    *ts_io = ts;
}

/* the accumulator alone (for pinning the segment table at full buffer sizes) */
void ref_shift_ts(double *out, size_t n, unsigned sample_rate, double *ts_io) {
    double ts = *ts_io;
    const double inc = 1.0 / (double)sample_rate;
    const double tau = REF_TAU;
    for (size_t j = 0; j < n; j++) {
        ts += inc;
        if (ts > tau) ts -= tau;
        if (out) out[j] = ts;
    }
    *ts_io = ts;
}

/* ---- a10: RotateComplex, internal/simd/mult.go:29-33 ------------------------------- */
void ref_rotate(float *buf, size_t n, float mr, float mi) {
    for (size_t i = 0; i < n; i++)
        go_c64_mul(buf[2 * i], buf[2 * i + 1], mr, mi, &buf[2 * i], &buf[2 * i + 1]);
}

/* ---- a11: ScaleComplex, internal/simd/mult_simd_amd64.s:27-55 (MULPS, 2 c64/iter) --- */
void ref_scale(float *buf, size_t n, float r) {
    const __m128 rr = _mm_set1_ps(r);
    size_t i = 0;
    if (n & 1) {
        buf[0] *= r;
        buf[1] *= r;
        i = 1;
    }
    for (; i < n; i += 2) _mm_storeu_ps(buf + 2 * i, _mm_mul_ps(_mm_loadu_ps(buf + 2 * i), rr));
}

/* ---- a12: AddComplex, internal/simd/add_simd_amd64.s:27-71 (ADDPS): c = a + b -------- */
void ref_add(const float *a, const float *b, float *c, size_t n) {
    size_t i = 0;
    if (n & 1) {
        c[0] = a[0] + b[0];
        c[1] = a[1] + b[1];
        i = 1;
    }
    for (; i < n; i += 2)
        _mm_storeu_ps(c + 2 * i, _mm_add_ps(_mm_loadu_ps(a + 2 * i), _mm_loadu_ps(b + 2 * i)));
}

/* ---- a8: DecimateBuffer (c64), stream/decimate.go:84-98; returns samples written ---- */
size_t ref_decimate_c64(const float *from, size_t n, float *to, unsigned factor) {
    size_t m = n / factor;
    for (size_t i = 0; i < m; i++) {
        to[2 * i] = from[2 * (size_t)factor * i];
        to[2 * i + 1] = from[2 * (size_t)factor * i + 1];
    }
    return m;
}

/* ---- a9: DownsampleBuffer (c64), stream/downsample.go:97-124 ----------------------- */
size_t ref_downsample_c64(const float *from, size_t n, float *to, unsigned factor) {
    size_t m = n / factor;
    for (size_t i = 0; i < m; i++) {
        float sr = 0, si = 0;
        for (unsigned j = 0; j < factor; j++) {
            sr += from[2 * (i * factor + j)];
            si += from[2 * (i * factor + j) + 1];
        }
        to[2 * i] = sr / (float)factor;
        to[2 * i + 1] = si / (float)factor;
    }
    return m;
}

/* ---- Planner stand-in: in-place radix-2, unnormalised.  dir<0 forward, >0 backward --- */
typedef struct {
    size_t n;
    float *tw; /* n/2 complex twiddles e^{-2 pi i k / n} */
    uint32_t *rev;
} ref_plan;

ref_plan *ref_plan_create(size_t n) {
    if (n == 0 || (n & (n - 1))) return NULL;
    ref_plan *p = (ref_plan *)malloc(sizeof(ref_plan));
    p->n = n;
    p->tw = (float *)malloc(sizeof(float) * (n > 1 ? n : 2));
    p->rev = (uint32_t *)malloc(sizeof(uint32_t) * n);
    for (size_t k = 0; k < n / 2; k++) {
        double a = -REF_TAU * (double)k / (double)n;
        p->tw[2 * k] = (float)cos(a);
        p->tw[2 * k + 1] = (float)sin(a);
    }
    unsigned bits = 0;
    while (((size_t)1 << bits) < n) bits++;
    for (size_t i = 0; i < n; i++) {
        uint32_t r = 0;
        for (unsigned b = 0; b < bits; b++)
            if (i & ((size_t)1 << b)) r |= 1u << (bits - 1 - b);
        p->rev[i] = r;
    }
    return p;
}

void ref_plan_destroy(ref_plan *p) {
    if (!p) return;
    free(p->tw);
    free(p->rev);
    free(p);
}

void ref_fft(const ref_plan *p, float *x, int dir) {
    const size_t n = p->n;
    for (size_t i = 0; i < n; i++) {
        size_t r = p->rev[i];
        if (r > i) {
            float tr = x[2 * i], ti = x[2 * i + 1];
            x[2 * i] = x[2 * r];
            x[2 * i + 1] = x[2 * r + 1];
            x[2 * r] = tr;
            x[2 * r + 1] = ti;
        }
    }
    const float sgn = dir > 0 ? -1.0f : 1.0f; /* backward conjugates the twiddle */
    for (size_t len = 2; len <= n; len <<= 1) {
        const size_t half = len >> 1, stride = n / len;
        for (size_t s = 0; s < n; s += len) {
            for (size_t k = 0; k < half; k++) {
                const float wr = p->tw[2 * k * stride], wi = sgn * p->tw[2 * k * stride + 1];
                float *a = x + 2 * (s + k), *b = x + 2 * (s + k + half);
                const float tr = b[0] * wr - b[1] * wi, ti = b[0] * wi + b[1] * wr;
                b[0] = a[0] - tr;
                b[1] = a[1] - ti;
                a[0] += tr;
                a[1] += ti;
            }
        }
    }
}

/* ---- a7: ConvolveFreq over consecutive N-blocks, fft/convolution.go:183-191 +
 *          stream/convolution.go:62-80 (block-circular, in place) ------------------- */
void ref_convolution_blocks(const ref_plan *p, float *iq, size_t nblocks, const float *filter) {
    const size_t n = p->n;
    for (size_t b = 0; b < nblocks; b++) {
        float *x = iq + 2 * b * n;
        ref_fft(p, x, -1);
        for (size_t i = 0; i < n; i++)
            go_c64_mul(x[2 * i], x[2 * i + 1], filter[2 * i], filter[2 * i + 1], &x[2 * i], &x[2 * i + 1]);
        ref_fft(p, x, +1);
    }
}

/* ---- the composed chain of SURVEY.md 2.3b on one raw buffer ------------------------
 * fmt: 2 = u8, 3 = i16, 4 = i8 (sdr.SampleFormat ids, iq.go:113-129).
 * scratch: n c64 samples.  Returns samples written to `out`. */
size_t ref_chain(const void *raw, int fmt, size_t n, unsigned sample_rate, double shift_hz,
                 const ref_plan *plan, const float *filter, unsigned decim, double *ts_io,
                 float *scratch, float *out) {
    if (fmt == 2)
        ref_convert_u8_c64((const uint8_t *)raw, scratch, n);
    else if (fmt == 4)
        ref_convert_i8_c64((const int8_t *)raw, scratch, n);
    else
        ref_convert_i16_c64((const int16_t *)raw, scratch, n);
    ref_shift_buffer(scratch, n, shift_hz, sample_rate, ts_io);
    const size_t nfft = plan->n;
    const size_t nblk = n / nfft;
    ref_convolution_blocks(plan, scratch, nblk, filter);
    const size_t lz = nblk * nfft, dblk = 32768;
    size_t w = 0;
    for (size_t q = 0; q + dblk <= lz; q += dblk) /* stream/decimate.go:40-50 */
        w += ref_decimate_c64(scratch + 2 * q, dblk, out + 2 * w, decim);
    return w;
}

/* ---- a13: ReadBeamform data path, stream/beamform.go:148-171 ----------------------- *
 * chans: nchan pointers to raw u8 IQ; weights: nchan complex; tmp: n c64 scratch. */
void ref_beamform_u8(const uint8_t *const *chans, size_t nchan, const float *weights, size_t n,
                     float *tmp, float *out) {
    memset(out, 0, sizeof(float) * 2 * n); /* stream/add.go:165-167 */
    for (size_t c = 0; c < nchan; c++) {
        ref_convert_u8_c64(chans[c], tmp, n);
        if (!(weights[2 * c] == 1.0f && weights[2 * c + 1] == 0.0f)) /* multiply.go:59-62 */
            ref_rotate(tmp, n, weights[2 * c], weights[2 * c + 1]);
        ref_add(out, tmp, out, n); /* stream/add.go:115-119 */
    }
}
