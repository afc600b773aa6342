// nco.cuh -- index-addressed NCO for stream.ShiftBuffer (stream/shifter.go:66-85).
//
// The reference's phase is a serially-rounded fp64 *time* accumulator:
//     ts += 1/fs;  if ts > 2*pi { ts -= 2*pi };  rot = sincos(2*pi*shift*ts)
// A GPU cannot run that loop, and an ideal (j+1)/fs phase misses it by > 1e-5 relative L2 within
// one 2^22-sample buffer.  But inside one fp64 binade every `ts += inc` adds the
// same grid-rounded step, so the accumulator is piecewise linear in the sample index:
//     ts[j0 + k] = base + (k+1)*step      (exact in fp64)
// build_segments() emits that table on the host (a few dozen entries at stream start, 1-3 for a
// steady-state buffer); binade crossings, the 2*pi wrap and round-half-even tie binades are
// emitted as single real steps.  It is bit-equal to the serial loop (tests/test_oracle.py pins
// the same construction; tests/test_gpu_parity.py pins this C++ copy through `ts`).
//
// The device never touches fp64: per segment the host converts base/step into 64-bit fixed-point
// *turns* of the mixer, P0 = frac(shift*base)*2^64 and dP = frac(shift*step)*2^64 (error-free
// two-product, so the only error is the 2^-65 rounding of dP, <= 2^-41 turns after 2^24 steps).
// phase(k) = P0 + (k+1)*dP wraps mod 2^64 for free, and the top bits give quadrant + a small
// residual angle for a short fp32 minimax sincos.
#pragma once
#include <math.h>
#include <stdint.h>

#include <vector>

#include "common.cuh"

namespace hz {

struct NcoSegment {
    uint32_t j0;     // first sample index (relative to the launch) this segment covers
    uint32_t count;  // samples in the segment
    uint64_t p0;     // phase before the segment's first step, turns * 2^64
    uint64_t dp;     // phase step per sample, turns * 2^64 (0 for single-step segments)
};

constexpr int kMaxSegsPerLaunch = 112;  // 112 * 24 B = 2688 B of kernel parameters

struct NcoTable {
    int count;
    NcoSegment seg[kMaxSegsPerLaunch];
};

// Per-stream descriptor of a batched launch (channelizer streams, or consecutive buffers of one stream).  Its
// accumulator segments sit in a pool next to the descriptors (a steady-state buffer needs 1-3, a stream-start or
// post-wrap buffer up to ~90).  Two homes for both: device memory, copied in front of the launch (the
// channelizer's hundreds of streams), or the kernel's own parameters (BatchTable: up to 64 buffers of one
// stream per launch -- no copy in the stream, so consecutive batched launches can overlap like single ones).
struct StreamDesc {
    const uint8_t *src;
    void *dst;
    uint32_t seg_off;  // first segment in the pool
    int count;         // segments, launch-relative sample indices
    uint64_t dp_nom;   // split launches (chain1024.cu): phase step of the stream's dominant segment
};
constexpr int kParamStreams = 64, kParamSegs = 416;
struct BatchTable {  // 64 x 32 B + 416 x 24 B = 12 KB of kernel parameters
    StreamDesc desc[kParamStreams];
    NcoSegment seg[kParamSegs];
};
struct SegView {  // what nco_find / NcoCursor::seek need of a table
    const NcoSegment *seg;
    int count;
};

// frac(a*b) * 2^64 (mod 2^64), exact product via FMA
inline uint64_t turns_fix(double a, double b) {
    double p = a * b;
    double e = fma(a, b, -p);
    double fp = p - floor(p);                  // exact, in [0,1)
    uint64_t hi = (uint64_t)ldexp(fp, 64);     // exact scaling; fp < 1
    if (fp >= 1.0) hi = 0;                     // cannot happen, belt and braces
    int64_t lo = (int64_t)llrint(ldexp(e, 64));  // |e| <= ulp(p)/2 << 2^-1
    return hi + (uint64_t)lo;
}

struct HostSeg {
    uint64_t j0, count;
    double base, step;  // ts[j0+k] = base + (k+1)*step; step == 0: ts[j0] = base (already stepped)
};

// Mirrors oracle.shift_segments.  Advances *ts exactly as n iterations of the reference loop.
inline void build_segments(uint32_t sample_rate, uint64_t n, double *ts_io, std::vector<HostSeg> &out) {
    const double inc = 1.0 / (double)sample_rate;
    const double tau = M_PI * 2;
    double ts = *ts_io;
    uint64_t j = 0;
    out.clear();
    while (j < n) {
        uint64_t m = 0;
        double step = 0.0;
        if (ts > 0.0) {
            int e;
            frexp(ts, &e);                                   // ts = f * 2^e, f in [0.5, 1)
            const double u = ldexp(1.0, e - 53);             // ulp(ts)
            const double q = inc / u;
            if (q < 4503599627370496.0 && (q - floor(q)) != 0.5) {
                step = (ts + inc) - ts;                      // inc rounded to ts's grid
                double top = ldexp(1.0, e);
                if (top > tau) top = tau;
                if (step > 0.0) {
                    double room = floor((top - ts) / step) - 2.0;
                    if (room > 0.0) {
                        m = room > (double)(n - j) ? (n - j) : (uint64_t)room;
                    }
                }
            }
        }
        if (m > 0) {
            out.push_back({j, m, ts, step});
            ts = ts + (double)m * step;
            j += m;
        } else {
            double nxt = ts + inc;
            if (nxt > tau) nxt -= tau;
            out.push_back({j, 1, nxt, 0.0});
            ts = nxt;
            j += 1;
        }
    }
    *ts_io = ts;
}

inline NcoSegment to_device_segment(const HostSeg &s, uint64_t launch_origin, double shift_hz) {
    NcoSegment d;
    d.j0 = (uint32_t)(s.j0 - launch_origin);
    d.count = (uint32_t)s.count;
    if (s.step == 0.0) {
        d.p0 = turns_fix(shift_hz, s.base);
        d.dp = 0;
    } else {
        d.p0 = turns_fix(shift_hz, s.base);
        d.dp = turns_fix(shift_hz, s.step);
    }
    return d;
}

#ifdef __CUDACC__
// segment lookup: the table is tiny and almost always uniform across a warp
template <class Table>
__device__ __forceinline__ int nco_find(const Table &t, uint32_t j) {
    int lo = 0, hi = t.count - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (t.seg[mid].j0 <= j)
            lo = mid;
        else
            hi = mid - 1;
    }
    return lo;
}

__device__ __forceinline__ uint64_t nco_phase(const NcoSegment &s, uint32_t j) {
    // single-step segments carry dp == 0 and p0 = the phase of their (only) sample
    return s.p0 + (uint64_t)(j - s.j0 + 1) * s.dp;
}

// a lane's cached segment: re-looked-up only when the walk leaves it
struct NcoCursor {
    uint32_t j0 = 1, end = 0;
    uint64_t p0 = 0, dp = 0;
    template <class Table>
    __device__ __forceinline__ void seek(const Table &t, uint32_t j) {
        if (j >= j0 && j < end) return;
        const int s = nco_find(t, j);
        j0 = t.seg[s].j0;
        end = j0 + t.seg[s].count;
        p0 = t.seg[s].p0;
        dp = t.seg[s].dp;
    }
    __device__ __forceinline__ uint64_t phase(uint32_t j) const { return p0 + (uint64_t)(j - j0 + 1) * dp; }
};

// e^{i * 2*pi * ph / 2^64} in fp32: quadrant from the top bits, minimax polynomials on
// [-pi/4, pi/4] (max error ~1 ulp of fp32; total phase error <= ~6e-8 rad).
__device__ __forceinline__ float2 nco_rot(uint64_t ph) {
    const uint32_t hi = (uint32_t)(ph >> 32);
    const uint32_t q = (hi + 0x20000000u) >> 30;          // nearest quarter turn
    const int32_t r = (int32_t)(hi - (q << 30));           // residual, units of 2^-32 turn, |r| <= 2^29
    const float x = (float)r * 1.4629180792671596e-9f;     // * 2*pi / 2^32
    const float x2 = x * x;
    float s = fmaf(x2, -1.9515295891e-4f, 8.3321608736e-3f);
    s = fmaf(s, x2, -1.6666654611e-1f);
    s = fmaf(s * x2, x, x);
    float c = fmaf(x2, 2.443315711809948e-5f, -1.388731625493765e-3f);
    c = fmaf(c, x2, 4.166664568298827e-2f);
    c = fmaf(c * x2, x2, fmaf(x2, -0.5f, 1.0f));
    // quadrant: (c, s), (-s, c), (-c, -s), (s, -c) -- two selects and two sign flips, no branches
    const bool odd = (q & 1u) != 0u;
    const float a = odd ? s : c, b = odd ? c : s;
    float2 o;
    o.x = __uint_as_float(__float_as_uint(a) ^ (((q + 1u) & 2u) << 30));
    o.y = __uint_as_float(__float_as_uint(b) ^ ((q & 2u) << 30));
    return o;
}
#endif

}  // namespace hz
