// bigfft.cu -- fft.Plan.Transform for lengths beyond one CTA's shared memory: 2^15 .. 2^20 points.
//
// The reference's Kerberos helpers plan 65536-point transforms (cross-correlation alignment,
// rtl/kerberos/internal/align.go:92-100) and an nReaders x 65536-point inverse (frequency-domain
// grafting, graft.go:73-80), so the planner has to go past the 16384 points the single-kernel
// transforms cover.  N = N1 x N2 (both in 128 .. 1024), two kernels through a scratch buffer:
//
//   step 1  for every n2 < N2:  A[k1][n2] = W_N^{n2 k1} * sum_{n1} x[N2 n1 + n2] W_N1^{n1 k1}
//           (N2 transforms of length N1 over stride-N2 columns; a CTA takes F adjacent columns so
//           that every global access is an F*8-byte run)
//   step 2  for every k1 < N1:  X[k1 + N1 k2] = sum_{n2} A[k1][n2] W_N2^{n2 k2}
//           (N1 transforms of length N2 over contiguous rows, written out transposed, again F
//           adjacent k1 per CTA)
//
// Both steps are the same kernel: stage a tile in shared memory, run the register-resident
// transforms of fft.cuh on it (F transforms x T threads), stage the results, write them out.
// HBM traffic is 32 B per point (two reads, two writes) against the 16 B of a single pass.
#include <algorithm>

#include "common.cuh"
#include "fft.cuh"

namespace hz {

template <int NF>
struct TileCfg {
    using C = FftCfg<NF>;
    static constexpr int T = C::T;
    static constexpr int threads = 256;
    static constexpr int F = threads / T;   // transforms per CTA
    static constexpr int FP = F + 1;        // padded row of the staging tile
    static constexpr size_t stage_elems = (size_t)NF * FP;
    static constexpr size_t smem_bytes = (stage_elems + (size_t)F * smem_elems(NF)) * sizeof(float2);
};

struct BigFftParams {
    const float2 *src;
    float2 *dst;
    const float2 *tw;   // W_NF table of the sub-transform
    uint32_t n_total;   // N
    uint32_t n_other;   // the other factor (number of sub-transforms per big transform)
    uint32_t batch;     // big transforms
    int step;           // 1: columns in -> columns out + twiddle;  2: rows in -> transposed out
};

template <int NF, int DIR>
__global__ void __launch_bounds__(256) k_fft_tile(const BigFftParams p) {
    using TC = TileCfg<NF>;
    using C = FftCfg<NF>;
    constexpr int P = C::P, T = C::T, F = TC::F, FP = TC::FP, RL = C::RL;
    extern __shared__ float2 smem[];
    float2 *stage = smem;
    const int f = threadIdx.x / T, t = threadIdx.x % T;
    float2 *sm = smem + TC::stage_elems + (size_t)f * smem_elems(NF);

    const uint32_t tiles = p.n_other / F;
    for (uint32_t w = blockIdx.x; w < tiles * p.batch; w += gridDim.x) {
        const uint32_t b = w / tiles, o0 = (w - b * tiles) * F;  // first sub-transform of this tile
        const float2 *x = p.src + (size_t)b * p.n_total;
        float2 *y = p.dst + (size_t)b * p.n_total;

        // global -> stage[j][f]
        if (p.step == 1) {
            for (uint32_t idx = threadIdx.x; idx < (uint32_t)NF * F; idx += 256) {
                const uint32_t j = idx / F, ff = idx % F;
                stage[j * FP + ff] = x[(size_t)j * p.n_other + o0 + ff];
            }
        } else {
            for (uint32_t idx = threadIdx.x; idx < (uint32_t)NF * F; idx += 256) {
                const uint32_t ff = idx / NF, j = idx % NF;
                stage[j * FP + ff] = x[(size_t)(o0 + ff) * NF + j];
            }
        }
        __syncthreads();

        float2 v[P];
        static_for<P / C::R1>([&](auto I) {
            constexpr int i = decltype(I)::value;
            static_for<C::R1>([&](auto RR) {
                constexpr int r = decltype(RR)::value;
                v[i * C::R1 + r] = stage[((t + T * i) + r * (NF / C::R1)) * FP + f];
            });
        });
        __syncthreads();  // the tile is in registers: stage may be overwritten with results
        fft_regs<NF, P, C::R1, C::R2, C::R3, DIR>(v, sm, p.tw, t);

        constexpr int NS = NF / RL;
        const float inv_half_n = 2.0f / (float)p.n_total;  // exact: N is a power of two
        static_for<P / RL>([&](auto I) {
            constexpr int i = decltype(I)::value;
            static_for<RL>([&](auto QQ) {
                constexpr int q = decltype(QQ)::value;
                const int k = pass_out_index<NF, P, RL, NS>(t, i, q);
                float2 val = v[i * RL + bitrev(q, ilog2(RL))];
                if (p.step == 1) {  // W_N^{n2 k1}: n2 k1 < N, so 2 n2 k1 / N is exact in fp32
                    float s, c;
                    sincospif((float)((o0 + f) * (uint32_t)k) * inv_half_n, &s, &c);
                    val = tw_mul<DIR>(val, c, s);
                }
                stage[k * FP + f] = val;
            });
        });
        __syncthreads();

        // stage[k][f] -> global
        if (p.step == 1) {
            for (uint32_t idx = threadIdx.x; idx < (uint32_t)NF * F; idx += 256) {
                const uint32_t k = idx / F, ff = idx % F;
                y[(size_t)k * p.n_other + o0 + ff] = stage[k * FP + ff];
            }
        } else {
            for (uint32_t idx = threadIdx.x; idx < (uint32_t)NF * F; idx += 256) {
                const uint32_t k = idx / F, ff = idx % F;
                y[(size_t)k * p.n_other + o0 + ff] = stage[k * FP + ff];  // X[k1 + N1 k2], k1 = o0 + ff
            }
        }
        __syncthreads();
    }
}

template <int NF>
static int launch_tile(hzsdr_ctx *ctx, int dir, const BigFftParams &p) {
    using TC = TileCfg<NF>;
    static PerDevice attr_set[2];
    const void *fn = dir < 0 ? (const void *)k_fft_tile<NF, FFT_FWD> : (const void *)k_fft_tile<NF, FFT_BWD>;
    const int di = dir < 0 ? 0 : 1;
    int rc = attr_set[di].once(ctx->device, [&](int &) {
        HZ_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC::smem_bytes));
        return (int)HZSDR_OK;
    });
    if (rc) return rc;
    const size_t work = (size_t)(p.n_other / TC::F) * p.batch;
    const int grid = (int)std::min<size_t>(work, (size_t)ctx->sm_count * 2);
    if (dir < 0)
        k_fft_tile<NF, FFT_FWD><<<grid, 256, TC::smem_bytes, ctx->stream>>>(p);
    else
        k_fft_tile<NF, FFT_BWD><<<grid, 256, TC::smem_bytes, ctx->stream>>>(p);
    HZ_CHECK_LAUNCH();
    return HZSDR_OK;
}

static int dispatch_tile(hzsdr_ctx *ctx, int nf, int dir, const BigFftParams &p) {
    switch (nf) {
        case 128: return launch_tile<128>(ctx, dir, p);
        case 256: return launch_tile<256>(ctx, dir, p);
        case 512: return launch_tile<512>(ctx, dir, p);
        case 1024: return launch_tile<1024>(ctx, dir, p);
        default: return fail(HZSDR_ERR_UNSUPPORTED, "big FFT factor %d", nf);
    }
}

// N = N1 * N2 with N1 <= N2, both powers of two in [128, 1024]
bool bigfft_len_ok(size_t n) { return n >= (1u << 15) && n <= (1u << 20) && (n & (n - 1)) == 0; }
void bigfft_factors(size_t n, int *n1, int *n2) {
    int lg = 0;
    while (((size_t)1 << lg) < n) lg++;
    *n1 = 1 << (lg / 2);
    *n2 = 1 << (lg - lg / 2);
}

// dir: FFT_FWD / FFT_BWD.  tw1 / tw2: the W_N1 / W_N2 tables.  scratch: n * batch complex64, != src, != dst.
int launch_bigfft(hzsdr_ctx *ctx, size_t n, int dir, const float2 *src, float2 *dst, float2 *scratch, size_t batch,
                  const float2 *tw1, const float2 *tw2) {
    int n1, n2;
    bigfft_factors(n, &n1, &n2);
    BigFftParams p{};
    p.n_total = (uint32_t)n;
    p.batch = (uint32_t)batch;
    // step 1: N2 column transforms of length N1
    p.src = src;
    p.dst = scratch;
    p.tw = tw1;
    p.n_other = (uint32_t)n2;
    p.step = 1;
    int rc = dispatch_tile(ctx, n1, dir, p);
    if (rc) return rc;
    // step 2: N1 row transforms of length N2, transposed out
    p.src = scratch;
    p.dst = dst;
    p.tw = tw2;
    p.n_other = (uint32_t)n1;
    p.step = 2;
    return dispatch_tile(ctx, n2, dir, p);
}

}  // namespace hz
