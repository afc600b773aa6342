// fft.cuh -- register-resident Stockham FFT building blocks (complex64, unnormalised).
//
// A transform of N = R1*R2*R3 points is done by T = N/P threads holding P points each (P = 8, 16
// or 32).  Every pass is the Stockham step
//     v[r]  = in[j + r*N/R] * W_{Ns*R}^{r*(j mod Ns)}           (gather + twiddle)
//     V     = DFT_R(v)                                           (in registers, fully unrolled)
//     out[(j div Ns)*Ns*R + (j mod Ns) + q*Ns] = V[q]            (scatter)
// with Ns the product of the radices already done; after the last pass the data is in natural
// order, no bit-reversal pass.  Passes exchange data through shared memory (one padded buffer per
// transform, 2 barriers per exchange); the first pass reads its input and the last pass delivers
// its output straight from/to registers, so a caller can fuse a producer (convert + NCO mix) and a
// consumer (pointwise multiply, decimating store) without touching memory.
//
// Useful identity for FFT -> multiply -> IFFT fusion: the last forward pass (radix R, Ns = N/R)
// leaves item j holding X[j + q*N/R], q < R -- exactly the elements the FIRST pass of the next
// transform (radix R, Ns = 1) gathers for item j.  So the spectrum never leaves the registers.
#pragma once
#include "common.cuh"

namespace hz {

// cos(pi*k/16), k = 0..8
constexpr double kCosPi16[9] = {1.0,
                                0.98078528040323044912618223613424,
                                0.92387953251128675612818318939679,
                                0.83146961230254523707878837761791,
                                0.70710678118654752440084436210485,
                                0.55557023301960222474283081394853,
                                0.38268343236508977172845998403040,
                                0.19509032201612826784828486847702,
                                0.0};
constexpr double cos_pi16(int k) {
    k &= 31;
    if (k > 16) k = 32 - k;
    return k <= 8 ? kCosPi16[k] : -kCosPi16[16 - k];
}
constexpr double sin_pi16(int k) { return cos_pi16(k - 8); }

constexpr int bitrev(int x, int bits) {
    int r = 0;
    for (int b = 0; b < bits; b++)
        if (x & (1 << b)) r |= 1 << (bits - 1 - b);
    return r;
}
constexpr int ilog2(int x) {
    int l = 0;
    while ((1 << l) < x) l++;
    return l;
}

constexpr int FFT_FWD = -1;  // e^{-2 pi i kn/N}   fft.Forward
constexpr int FFT_BWD = +1;  // e^{+2 pi i kn/N}   fft.Backward

// d * (c + i*DIR*s): (fma(+-d.y, s, d.x c), fma(-+d.x, s, d.y c)), two packed instructions
template <int DIR>
__device__ __forceinline__ float2 tw_mul(float2 d, float c, float s) {
    const float2 m = mul2(d, make_float2(c, c));
    if constexpr (DIR < 0)
        return fma2(make_float2(d.y, -d.x), make_float2(s, s), m);
    else
        return fma2(make_float2(-d.y, d.x), make_float2(s, s), m);
}

// In-register DFT of R points on v[BASE .. BASE+R): natural-order in, bit-reversed out (DFT value q
// ends up in v[BASE + bitrev(q)]).  R in {1,2,4,8,16,32}; all twiddles are immediates.
//
// Decimation in time on the compile-time-permuted view u[m] = v[BASE + bitrev(m)] (so no data
// movement), with the FMA form of the butterfly:
//     u' = a + w b          2 FFMA2 (the twiddle multiply rides inside the add)
//     u''= a - w b = 2a - u'  1 FFMA2
// on packed (re, im) pairs (common.cuh): 3 instructions per non-trivial butterfly, 2 FADD2 for the
// trivial twiddles (1, -+i).  For R = 32: 46 trivial + 34 non-trivial butterflies = 194
// instructions (388 in scalar FMA form, 456 as multiply-then-add).
// stages S0 .. log2(R)-1 of the transform below (S0 = 0: all of it)
template <int R, int DIR, int BASE, int P, int S0>
__device__ __forceinline__ void fft_reg_from(float2 (&v)[P]) {
    constexpr int LOG2R = ilog2(R);
    static_for<LOG2R - S0>([&](auto SS) {
        constexpr int span = 1 << (decltype(SS)::value + S0);
        static_for<R / 2>([&](auto TT) {
            constexpr int tt = decltype(TT)::value;
            constexpr int i = tt % span, start = (tt / span) * 2 * span;
            constexpr int ia = BASE + bitrev(start + i, LOG2R), ib = BASE + bitrev(start + i + span, LOG2R);
            constexpr int k = i * (32 / (2 * span));  // w = W_{2 span}^i = e^{DIR * i*pi*k/16}
            const float2 a = v[ia], b = v[ib];
            if constexpr (k == 0) {
                v[ia] = add2(a, b);
                v[ib] = sub2(a, b);
            } else if constexpr (k == 8) {  // w = -i (forward) / +i (backward)
                const float2 p = make_float2(b.y, -b.x), m = make_float2(-b.y, b.x);
                v[ia] = add2(a, DIR < 0 ? p : m);
                v[ib] = add2(a, DIR < 0 ? m : p);
            } else {
                constexpr float c = (float)cos_pi16(k);
                constexpr float sn = DIR < 0 ? -(float)sin_pi16(k) : (float)sin_pi16(k);  // w = c + i*sn
                // re = fma(-sn, b.y, fma(c, b.x, a.x)),  im = fma(sn, b.x, fma(c, b.y, a.y))
                const float2 t = fma2(b, make_float2(c, c), a);
                const float2 u = fma2(make_float2(-b.y, b.x), make_float2(sn, sn), t);
                v[ia] = u;
                v[ib] = fma2(a, make_float2(2.0f, 2.0f), make_float2(-u.x, -u.y));
            }
        });
    });
}

template <int R, int DIR, int BASE, int P>
__device__ __forceinline__ void fft_reg(float2 (&v)[P]) {
    fft_reg_from<R, DIR, BASE, P, 0>(v);
}

// The same transform of v[n] * m_n, the multipliers m_n = pre(n) = (re, im) folded into the first
// butterfly stage.  That stage pairs (n, n + R/2) with a unit twiddle, so
//     a' = v[n] m_n                         2 instructions
//     u' = a' + m_{n+R/2} v[n+R/2]          2 FFMA2
//     u''= 2 a' - u'                        1 FFMA2
// 5 per pair against 6 for multiplying both first and adding after: R/2 instructions saved (and
// 2 more when UNIT0 says m_0 = 1).
template <int R, int DIR, int BASE, int P, bool UNIT0, class Pre>
__device__ __forceinline__ void fft_reg_pre(float2 (&v)[P], Pre pre) {
    constexpr int LOG2R = ilog2(R);
    static_for<R / 2>([&](auto TT) {
        constexpr int tt = decltype(TT)::value;
        constexpr int ia = bitrev(2 * tt, LOG2R), ib = bitrev(2 * tt + 1, LOG2R);  // ib = ia + R/2
        float2 a = v[BASE + ia];
        if constexpr (!(UNIT0 && ia == 0)) {
            const float2 ma = pre(std::integral_constant<int, ia>{});
            a = fma2(ma, make_float2(a.x, a.x), mul2(make_float2(-ma.y, ma.x), make_float2(a.y, a.y)));
        }
        const float2 b = v[BASE + ib], m = pre(std::integral_constant<int, ib>{});
        const float2 t = fma2(b, make_float2(m.x, m.x), a);
        const float2 u = fma2(make_float2(-b.y, b.x), make_float2(m.y, m.y), t);
        v[BASE + ia] = u;
        v[BASE + ib] = fma2(a, make_float2(2.0f, 2.0f), make_float2(-u.x, -u.y));
    });
    fft_reg_from<R, DIR, BASE, P, 1>(v);
}

// shared-memory index padding (units of float2): one pad slot every 32 elements makes the
// stride-R scatter of the first pass conflict-free for 64-bit accesses.
__device__ __forceinline__ int smem_pad(int a) { return a + (a >> 5); }
constexpr int smem_elems(int n) { return n + (n >> 5) + 1; }

template <int N, int P>
struct FftShape {
    static constexpr int T = N / P;  // threads per transform
};

// butterflies of one pass: P/R independent radix-R transforms per thread
template <int P, int R, int DIR>
__device__ __forceinline__ void pass_butterflies(float2 (&v)[P]) {
    static_for<P / R>([&](auto I) { fft_reg<R, DIR, decltype(I)::value * R, P>(v); });
}

// twiddle before the butterflies of a pass with radix R, Ns done so far.
// tw: W_N table, tw[m] = (cos(2 pi m/N), sin(2 pi m/N)), m < N.
template <int N, int P, int R, int NS, int DIR>
__device__ __forceinline__ void pass_twiddle(float2 (&v)[P], const float2 *__restrict__ tw, int t) {
    if constexpr (NS > 1) {
        constexpr int T = N / P;
        static_for<P / R>([&](auto I) {
            constexpr int i = decltype(I)::value;
            const int j = t + T * i;
            const int k = j & (NS - 1);
            static_for<R - 1>([&](auto RR) {
                constexpr int r = decltype(RR)::value + 1;
                const float2 w = __ldg(tw + r * k * (N / (NS * R)));
                v[i * R + r] = tw_mul<DIR>(v[i * R + r], w.x, w.y);
            });
        });
    }
}

// gather the inputs of a radix-R pass from shared memory
template <int N, int P, int R>
__device__ __forceinline__ void pass_gather(float2 (&v)[P], const float2 *sm, int t) {
    constexpr int T = N / P;
    static_for<P / R>([&](auto I) {
        constexpr int i = decltype(I)::value;
        const int j = t + T * i;
        static_for<R>([&](auto RR) {
            constexpr int r = decltype(RR)::value;
            v[i * R + r] = sm[smem_pad(j + r * (N / R))];
        });
    });
}

// destination index (natural order within the transform) of DFT value q of item i
template <int N, int P, int R, int NS>
__device__ __forceinline__ int pass_out_index(int t, int i, int q) {
    constexpr int T = N / P;
    const int j = t + T * i;
    if constexpr (NS == 1)
        return j * R + q;
    else
        return (j / NS) * (NS * R) + (j & (NS - 1)) + q * NS;
}

// scatter the outputs of a radix-R pass to shared memory
template <int N, int P, int R, int NS>
__device__ __forceinline__ void pass_scatter(const float2 (&v)[P], float2 *sm, int t) {
    static_for<P / R>([&](auto I) {
        constexpr int i = decltype(I)::value;
        static_for<R>([&](auto QQ) {
            constexpr int q = decltype(QQ)::value;
            sm[smem_pad(pass_out_index<N, P, R, NS>(t, i, q))] = v[i * R + bitrev(q, ilog2(R))];
        });
    });
}

// barrier between passes: transforms with T <= 32 live inside one warp
template <int T>
__device__ __forceinline__ void fft_sync() {
    if constexpr (T <= 32)
        __syncwarp();
    else
        __syncthreads();
}

// Full transform from registers to registers.
// In:  v[i*R1 + r] = x[(t + T*i) + r*N/R1]            (the first pass's gather pattern)
// Out: DFT value at natural index pass_out_index<..., RL, N/RL>(t, i, q) is in
//      v[i*RL + bitrev(q)], RL = the last radix (R3 if > 1, else R2 if > 1, else R1).
template <int N, int P, int R1, int R2, int R3, int DIR>
__device__ __forceinline__ void fft_regs(float2 (&v)[P], float2 *sm, const float2 *__restrict__ tw, int t) {
    constexpr int T = N / P;
    static_assert(R1 * R2 * R3 == N, "radices must multiply to N");
    pass_butterflies<P, R1, DIR>(v);
    if constexpr (R2 > 1) {
        fft_sync<T>();  // everybody is done reading the buffer (previous exchange)
        pass_scatter<N, P, R1, 1>(v, sm, t);
        fft_sync<T>();
        pass_gather<N, P, R2>(v, sm, t);
        pass_twiddle<N, P, R2, R1, DIR>(v, tw, t);
        pass_butterflies<P, R2, DIR>(v);
        if constexpr (R3 > 1) {
            fft_sync<T>();
            pass_scatter<N, P, R2, R1>(v, sm, t);
            fft_sync<T>();
            pass_gather<N, P, R3>(v, sm, t);
            pass_twiddle<N, P, R3, R1 * R2, DIR>(v, tw, t);
            pass_butterflies<P, R3, DIR>(v);
        }
    }
}

template <int R1, int R2, int R3>
struct LastRadix {
    static constexpr int R = R3 > 1 ? R3 : (R2 > 1 ? R2 : R1);
};

// Per-length decomposition.  P = points per thread.
template <int N>
struct FftCfg;
#define HZ_FFT_CFG(n, p, r1, r2, r3)                                   \
    template <>                                                        \
    struct FftCfg<n> {                                                 \
        static constexpr int P = p, R1 = r1, R2 = r2, R3 = r3;         \
        static constexpr int T = n / p;                                \
        static constexpr int RL = LastRadix<r1, r2, r3>::R;            \
    };
HZ_FFT_CFG(2, 2, 2, 1, 1)
HZ_FFT_CFG(4, 4, 4, 1, 1)
HZ_FFT_CFG(8, 8, 8, 1, 1)
HZ_FFT_CFG(16, 16, 16, 1, 1)
HZ_FFT_CFG(32, 32, 32, 1, 1)
HZ_FFT_CFG(64, 8, 8, 8, 1)
HZ_FFT_CFG(128, 16, 16, 8, 1)
HZ_FFT_CFG(256, 16, 16, 16, 1)
HZ_FFT_CFG(512, 32, 32, 16, 1)
HZ_FFT_CFG(1024, 32, 32, 32, 1)
HZ_FFT_CFG(2048, 16, 16, 16, 8)
HZ_FFT_CFG(4096, 16, 16, 16, 16)
HZ_FFT_CFG(8192, 32, 32, 16, 16)
HZ_FFT_CFG(16384, 32, 32, 32, 16)
#undef HZ_FFT_CFG

// transforms per CTA: keep CTAs at >= 128 threads
template <int N>
struct FftCta {
    static constexpr int T = FftCfg<N>::T;
    static constexpr int F = T >= 128 ? 1 : 128 / T;
    static constexpr int threads = T * F;
    static constexpr size_t smem_bytes = (size_t)F * smem_elems(N) * sizeof(float2);
};

}  // namespace hz
