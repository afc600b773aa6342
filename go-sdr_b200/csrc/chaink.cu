// chaink.cu -- the fused chain for ConvolutionReader blocks of N = K * 1024, K = 2, 4, 8
// (2047-, 4095-, 8191-tap filters): one CTA of K warps per block.
//
// A length-N circular convolution splits, by one radix-K decimation-in-frequency step, into K
// independent length-1024 circular convolutions -- exactly what one warp does in registers
// (chain1024.cu):
//   phase 1  all 32K threads: raw samples -> float -> NCO mix; radix-K butterfly across the K
//            quarters of the block (n = m + 1024 c), times W_N^{m k1}  ->  x[k1][m]
//   phase 2  warp k1: 1024-point FFT of x[k1], times H[k1 + K k2], 1024-point inverse (four passes
//            of the same radix-32 code, the inverse on re/im-swapped data), in place
//   phase 3  all threads: times W_N^{-m k1}, inverse radix-K butterfly  ->  z[m + 1024 c] in shared
//            memory; then DecimateReader: the kept samples z[q*DB + D*i] are copied out, coalesced
// The generic kernel (fft_kernels.cuh, k_chain<N>) runs these lengths as 128-256 threads behind ten
// CTA barriers per block and reaches 75-95 Gsamples/s; here four barriers separate phases, the
// longest of which (phase 2) is barrier-free.
//
// Algorithmic HBM bytes: raw bytes in + 8 B per kept sample out, as for every chain kernel.
#include "common.cuh"
#include "fft.cuh"
#include "fft_kernels.cuh"
#include "nco.cuh"

namespace hz {

template <int K>
struct ChainKSmem {
    float2 tw[32][32];    // W_1024^{r lane}: twiddles between the radix-32 passes of the sub-transforms
    float2 x[K][1056];    // x[k1]: sub-sequence k1, index padded m + m/32; warp k1's exchange space in phase 2
    float2 e[32 / K + K]; // NCO step tables of the current segment: e^{i 32K i dP} (i < 32/K), then e^{i 1024 c dP} (c < K)
};

__device__ __forceinline__ int kpad(int a) { return a + (a >> 5); }

// d * (b.x + i DIR b.y) * e^{i DIR 2 pi j / 32}: the launch-constant factor from a register, the per-column
// one (j = i k1 mod 32 after unrolling) from a small constant-memory table
__constant__ float2 kW32[32];
template <int DIR>
__device__ __forceinline__ float2 ck_twiddle(float2 d, float2 b, int j) {
    const float2 c = kW32[j & 31];
    return tw_mul<DIR>(tw_mul<DIR>(d, b.x, b.y), c.x, c.y);
}

__device__ __forceinline__ uint32_t ck_udiv(uint32_t x, uint32_t d, float inv_d) {
    // floor(x / d) for x < 2^24: float estimate (never too large: inv_d is rounded down) + 1 fix-up
    uint32_t q = __float2uint_rz(__uint2float_rz(x) * inv_d);
    if (x - q * d >= d) q++;
    return q;
}

template <int FMT, bool LSB>
__device__ __forceinline__ uint32_t ck_load_raw(const uint8_t *__restrict__ src, uint32_t j, int lsb_shift) {
    if constexpr (FMT == HZSDR_FORMAT_I16) {
        uint32_t w = __ldg(reinterpret_cast<const uint32_t *>(src) + j);
        if constexpr (LSB) w = ((w & 0xffff0000u) << lsb_shift) | (((w & 0xffffu) << lsb_shift) & 0xffffu);
        return w;
    } else {
        return (uint32_t)__ldg(reinterpret_cast<const uint16_t *>(src) + j);
    }
}

// prm.tw   = [32][32] W_1024^{r lane}        (chain1024_twiddles' first table)
// prm.tw3  = [K-1][1024] W_N^{m k1}, k1 >= 1  (cos, sin)
// prm.tw1k = [K][1024] the filter by sub-convolution: Hp[k1][k2] = H[k1 + K k2]
template <int FMT, int K, bool LSB>
__global__ void __launch_bounds__(32 * K, 512 / (32 * K)) k_chaink(const __grid_constant__ ChainParams prm,
                                                                   const __grid_constant__ NcoTable nco) {
    constexpr int TH = 32 * K, N = 1024 * K, PER = 32 / K;  // threads, block length, phase-1/3 butterflies per thread
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ChainKSmem<K> &S = *reinterpret_cast<ChainKSmem<K> *>(smem_raw);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;

    {   // table prologue: every load in flight before the first store (a load -> store loop pays one L2
        // latency per iteration, and a CTA only lives for a few blocks)
        constexpr int kPer = 32 * 32 / 2 / TH;
        float4 tmp[kPer];
#pragma unroll
        for (int u = 0; u < kPer; u++) tmp[u] = __ldg(reinterpret_cast<const float4 *>(prm.tw) + t + u * TH);
#pragma unroll
        for (int u = 0; u < kPer; u++) reinterpret_cast<float4 *>(&S.tw[0][0])[t + u * TH] = tmp[u];
    }

    // W_N^{m k1} for the thread's columns m = t + 32K i factors into W_N^{t k1} (K-1 values the thread keeps
    // in registers for the whole launch) times W_32^{i k1} (compile-time constants): no table reads per block
    float2 base[K > 1 ? K - 1 : 1];
    static_for<K - 1>([&](auto KK) {
        constexpr int k1 = decltype(KK)::value + 1;
        base[k1 - 1] = __ldg(prm.tw3 + (k1 - 1) * 1024 + t);  // (cos, sin)(2 pi t k1 / N), t < 32K <= 1024
    });
    const float sc = RawTraits<FMT>::scale();
    const uint32_t db_mask = (1u << prm.db_log2) - 1u;
    const float2 *hp = prm.tw1k + warp * 1024 + lane;
    uint64_t e_dp = 0;
    bool e_ok = false;

    for (uint32_t b = blockIdx.x; b < prm.nblocks; b += gridDim.x) {
        const uint32_t s0 = b * (uint32_t)N;
        // ------------------------------------------------------------------ phase 1
        const int si = nco_find(nco, s0);
        const uint32_t seg_j0 = nco.seg[si].j0, seg_end = seg_j0 + nco.seg[si].count;
        const uint64_t seg_p0 = nco.seg[si].p0, seg_dp = nco.seg[si].dp;
        const bool fast = s0 + (uint32_t)N <= seg_end;
        if (fast && (!e_ok || e_dp != seg_dp)) {  // uniform over the CTA
            if (t < PER)
                S.e[t] = nco_rot((uint64_t)(TH * t) * seg_dp);
            else if (t < PER + K)
                S.e[t] = nco_rot((uint64_t)(1024u * (t - PER)) * seg_dp);
            e_dp = seg_dp;
            e_ok = true;
        }
        __syncthreads();  // step tables (and, first time, tw) visible; the previous block's phase 3 has read x
        float2 r0 = make_float2(0.f, 0.f);
        if (fast) r0 = mul2(nco_rot(seg_p0 + (uint64_t)(s0 + t - seg_j0 + 1) * seg_dp), make_float2(sc, sc));
        NcoCursor cur;
        // all 32 raw loads of the thread in flight before the first use (a loop that loads K samples per
        // butterfly keeps 2-8 in flight and waits on HBM latency every iteration)
        uint32_t raw[32];
        if (fast) {
#pragma unroll
            for (int i = 0; i < PER; ++i)
                static_for<K>([&](auto CC) {
                    constexpr int c = decltype(CC)::value;
                    raw[i * K + c] = ck_load_raw<FMT, LSB>(prm.src, s0 + (uint32_t)(t + TH * i) + 1024u * c, prm.lsb_shift);
                });
        }
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int m = t + TH * i;
            float2 w[K];
            if (fast) {
                const float2 ri = i == 0 ? r0 : cmul(r0, S.e[i]);
                static_for<K>([&](auto CC) {
                    constexpr int c = decltype(CC)::value;
                    const float2 rot = c == 0 ? ri : cmul(ri, S.e[PER + c]);
                    w[c] = cmul(RawTraits<FMT>::unscaled2(raw[i * K + c]), rot);
                });
            } else {  // the block straddles accumulator segments: every sample's phase on its own
#pragma unroll 1
                for (int c = 0; c < K; ++c) {
                    const uint32_t j = s0 + m + 1024u * c;
                    cur.seek(nco, j);
                    const float2 rot = mul2(nco_rot(cur.phase(j)), make_float2(sc, sc));
                    S.x[c][kpad(m)] = cmul(RawTraits<FMT>::unscaled2(ck_load_raw<FMT, LSB>(prm.src, j, prm.lsb_shift)), rot);
                }
                static_for<K>([&](auto CC) {
                    constexpr int c = decltype(CC)::value;
                    w[c] = S.x[c][kpad(m)];  // this thread's own slots
                });
            }
            fft_reg<K, FFT_FWD, 0, K>(w);
            static_for<K>([&](auto KK) {
                constexpr int k1 = decltype(KK)::value;
                float2 val = w[bitrev(k1, ilog2(K))];
                if constexpr (k1 > 0) val = ck_twiddle<FFT_FWD>(val, base[k1 - 1], i * k1);
                S.x[k1][kpad(m)] = val;
            });
        }
        __syncthreads();

        // the CTA's next block -> L2 while this one is transformed (one 128-byte line per thread and step)
        if (b + gridDim.x < prm.nblocks) {
            constexpr uint32_t kBytes = (uint32_t)N * (uint32_t)RawTraits<FMT>::bytes;
            const uint8_t *nx = prm.src + (size_t)(b + gridDim.x) * kBytes;
            for (uint32_t o = 128u * t; o < kBytes; o += 128u * TH) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + o));
        }

        // ------------------------------------------------------------------ phase 2: warp k1 = `warp`
        {
            float2 *buf = S.x[warp];
            float2 v[32];
            static_for<32>([&](auto RR) {
                constexpr int r = decltype(RR)::value;
                v[r] = buf[lane + 33 * r];  // element lane + 32 r
            });
#pragma unroll 1
            for (int pass = 0; pass < 4; ++pass) {
                fft_reg<32, FFT_FWD, 0, 32>(v);
                __syncwarp();
                if (pass & 1) {
                    static_for<32>([&](auto QQ) {
                        constexpr int q = decltype(QQ)::value;
                        buf[lane + 33 * q] = v[bitrev(q, 5)];  // Ns = 32: index lane + 32 q
                    });
                } else {
                    static_for<32>([&](auto QQ) {
                        constexpr int q = decltype(QQ)::value;
                        buf[lane * 33 + q] = v[bitrev(q, 5)];  // Ns = 1: index 32 lane + q
                    });
                }
                __syncwarp();
                if (pass == 3) break;
                static_for<32>([&](auto RR) {
                    constexpr int r = decltype(RR)::value;
                    v[r] = buf[lane + 33 * r];
                });
                if (pass == 1) {
                    // sub-spectrum x sub-filter (fft/convolution.go:187-189), re/im swapped for the inverse
                    static_for<32>([&](auto RR) {
                        constexpr int r = decltype(RR)::value;
                        v[r] = cmul_swapped(v[r], __ldg(hp + 32 * r));
                    });
                } else {
                    static_for<31>([&](auto RR) {
                        constexpr int r = decltype(RR)::value + 1;
                        const float2 w = S.tw[r][lane];
                        v[r] = tw_mul<FFT_FWD>(v[r], w.x, w.y);
                    });
                }
            }
        }
        __syncthreads();

        // ------------------------------------------------------------------ phase 3
        // x[k1] holds swap(u_k1) in natural order; z[m + 1024 c] = sum_k1 u_k1[m] W_N^{-m k1} W_K^{-c k1}
#pragma unroll 4
        for (int i = 0; i < PER; ++i) {
            const int m = t + TH * i;
            float2 w[K];
            static_for<K>([&](auto KK) {
                constexpr int k1 = decltype(KK)::value;
                const float2 u = S.x[k1][kpad(m)];
                float2 val = make_float2(u.y, u.x);
                if constexpr (k1 > 0) val = ck_twiddle<FFT_BWD>(val, base[k1 - 1], i * k1);
                w[k1] = val;
            });
            fft_reg<K, FFT_BWD, 0, K>(w);
            static_for<K>([&](auto CC) {
                constexpr int c = decltype(CC)::value;
                S.x[c][kpad(m)] = w[bitrev(c, ilog2(K))];  // z[m + 1024 c], in place: only this thread touches column m
            });
        }
        __syncthreads();
        // DecimateReader: the block lies inside one decimate block (both are aligned powers of two, DB >= N);
        // the threads copy its kept samples z[D*i - p0] out of shared memory, coalesced
        const uint32_t g0 = prm.z0 + s0;
        const uint32_t p0 = g0 & db_mask;
        const uint32_t o0 = ck_udiv(p0 + prm.D - 1u, prm.D, prm.inv_d);  // first kept output index at or after the block's start
        const uint32_t pos0 = o0 * prm.D - p0;                            // its position inside the block
        uint32_t cnt = 0;
        if (pos0 < (uint32_t)N && o0 < prm.M) {
            cnt = ck_udiv((uint32_t)N - 1u - pos0, prm.D, prm.inv_d) + 1u;
            if (cnt > prm.M - o0) cnt = prm.M - o0;
        }
        float2 *out = prm.dst + (size_t)(g0 >> prm.db_log2) * prm.M + o0;
        for (uint32_t k = t; k < cnt; k += TH) {
            const uint32_t pp = pos0 + k * prm.D;
            out[k] = S.x[pp >> 10][kpad((int)(pp & 1023u))];
        }
    }
}

template <int FMT, int K, bool LSB>
static int launchk(hzsdr_ctx *ctx, const ChainParams &prm, const NcoTable &nco) {
    static PerDevice attr_set;
    const size_t smem = sizeof(ChainKSmem<K>);
    int rc = attr_set.once(ctx->device, [&](int &) {
        float2 w32[32];
        for (int j = 0; j < 32; j++) w32[j] = make_float2((float)cos(2.0 * M_PI * j / 32.0), (float)sin(2.0 * M_PI * j / 32.0));
        HZ_CUDA(cudaMemcpyToSymbol(kW32, w32, sizeof(w32)));
        HZ_CUDA(cudaFuncSetAttribute((const void *)k_chaink<FMT, K, LSB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        return (int)HZSDR_OK;
    });
    if (rc) return rc;
    size_t per_sm = 512 / (32 * K);
    const size_t by_smem = ((size_t)220 * 1024) / (smem + 1024);
    if (by_smem < per_sm) per_sm = by_smem;
    const size_t cap = (size_t)ctx->sm_count * per_sm;
    // equal shares: with r = ceil(nblocks / cap) rounds, ceil(nblocks / r) CTAs do r blocks each (no tail round
    // with a third of the chip idle, and no CTA pays the table prologue for a single block)
    const size_t rounds = (prm.nblocks + cap - 1) / cap;
    const int grid = (int)((prm.nblocks + rounds - 1) / rounds);
    k_chaink<FMT, K, LSB><<<grid, 32 * K, smem, ctx->stream>>>(prm, nco);
    HZ_CHECK_LAUNCH();
    return HZSDR_OK;
}

template <int K>
static int launchk_fmt(hzsdr_ctx *ctx, int fmt, const ChainParams &prm, const NcoTable &nco) {
    switch (fmt) {
        case HZSDR_FORMAT_U8: return launchk<HZSDR_FORMAT_U8, K, false>(ctx, prm, nco);
        case HZSDR_FORMAT_I8: return launchk<HZSDR_FORMAT_I8, K, false>(ctx, prm, nco);
        default:
            return prm.lsb_shift ? launchk<HZSDR_FORMAT_I16, K, true>(ctx, prm, nco) : launchk<HZSDR_FORMAT_I16, K, false>(ctx, prm, nco);
    }
}

int launch_chaink(hzsdr_ctx *ctx, int fmt, int k, const ChainParams &prm, const NcoTable &nco) {
    switch (k) {
        case 2: return launchk_fmt<2>(ctx, fmt, prm, nco);
        case 4: return launchk_fmt<4>(ctx, fmt, prm, nco);
        case 8: return launchk_fmt<8>(ctx, fmt, prm, nco);
        default: return fail(HZSDR_ERR_UNSUPPORTED, "launch_chaink: K = %d", k);
    }
}

// host tables: twn[(k1-1)*1024 + m] = (cos, sin)(2 pi m k1 / N); hp[k1*1024 + k2] = H[k1 + K k2]
void chaink_tables(int k, const float2 *H, float2 *twn, float2 *hp) {
    const int n = 1024 * k;
    for (int k1 = 1; k1 < k; k1++)
        for (int m = 0; m < 1024; m++) {
            const double a = 2.0 * M_PI * (double)(m * k1) / (double)n;
            twn[(k1 - 1) * 1024 + m] = make_float2((float)cos(a), (float)sin(a));
        }
    for (int k1 = 0; k1 < k; k1++)
        for (int k2 = 0; k2 < 1024; k2++) hp[k1 * 1024 + k2] = H[k1 + k * k2];
}

}  // namespace hz
