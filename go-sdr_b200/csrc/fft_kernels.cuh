// fft_kernels.cuh -- kernel templates and per-length launchers for K6 and the fused chain.
// Instantiated once per FFT length by fft_inst.cu (compiled with -DHZ_FFT_N=<n>), so the lengths
// build in parallel; fft.cu only sees the declarations at the bottom of this file.
#pragma once
#include "batch_admit.h"
#include "common.cuh"
#include "fft.cuh"
#include "nco.cuh"

namespace hz {

struct ChainParams {
    const uint8_t *src;
    float2 *dst;
    const float2 *tw;
    const float2 *H;
    uint32_t nblocks;   // N-sample blocks in this launch
    uint32_t z0;        // z-stream position (samples since the exec call started) of the launch's first sample
    uint32_t D;         // decimation factor
    uint32_t M;         // outputs per decimate block = DB / D
    uint32_t db_log2;   // log2 of the decimate block
    float inv_d;        // 1/D rounded toward zero
    int lsb_shift;      // i16 only: ShiftLSBToMSBBits (iq_i16.go:103-111), 0 = none
    // batched (channelizer) launches of the N = 1024 kernel: nblocks = blocks per stream
    const StreamDesc *streams;    // device descriptors, or nullptr: they travel as kernel parameters (BatchTable)
    const NcoSegment *seg_pool;   // device segment pool of `streams`
    uint32_t nstreams;
    // extra twiddle tables of the N = 16384 kernel (chain16k.cu)
    const float2 *tw3;   // [15][1024]  W_16384^{r j}
    const float2 *tw1k;  // [16384]     the filter in the kernel's read order (chain16k_permute_filter)
    uint32_t *done;      // this launch's slot of hzsdr_ctx::overlap_done (set by the launcher)
    // N = 1024 kernel, SPLIT form (chain1024.cu): prm.tw was built by chain1024_split_twiddles for dp_nom
    uint64_t dp_nom;
    int split;
    const float2 *tw_bc;  // [twB | twC] when prm.tw holds only the 32 x 32 part (null: they follow prm.tw)
    uint32_t per_cta;     // batched SPLIT launches: blocks in a CTA's contiguous range (set by the launcher)
    // N = 16384 kernel, overlap-save form (chain16k.cu, OS): nblocks = windows; launch coordinates start
    // os_hist = 16384 - os_hop samples before the first new sample; src points at launch coordinate 0
    // (for the first launch of a call that is os_hist samples in front of the buffer: never dereferenced there)
    const uint8_t *hist;   // the carried raw history: launch coordinates [0, os_head)
    uint8_t *hist_out;     // where this launch saves the buffer's tail for the next call (nullptr: not the call's last launch)
    const uint8_t *tail_src;
    uint32_t tail_bytes;
    uint32_t os_hop;       // window hop L (0: block-circular form)
    uint32_t os_head;      // samples taken from `hist` (os_hist in the first launch of a call, else 0)
    uint32_t os_valid;     // launch coordinates >= this are beyond the buffer's end: zero-filled
    uint32_t os_zend;      // outputs are kept for z positions < this (the call's length)
    uint32_t os_zero_head; // stream start: the history is silence
};

// May this chain launch start while earlier overlappable launches of the stream drain?  (common.cuh,
// OverlapWindow.)  Spans are conservative: whole decimate blocks on the output side.
inline bool chain_may_overlap(hzsdr_ctx *ctx, ChainParams &prm, uint32_t n_fft, int sample_bytes) {
    const size_t count = (size_t)prm.nblocks * n_fft;
    const size_t first = (size_t)(prm.z0 >> prm.db_log2) * prm.M;
    const size_t last = (size_t)(((prm.z0 + count - 1) >> prm.db_log2) + 1) * prm.M;
    const bool ok = ctx->overlap.admit(OverlapWindow::span(prm.src, count * (size_t)sample_bytes),
                                       OverlapWindow::span(prm.dst + first, (last - first) * sizeof(float2)),
                                       ctx->overlap_pred_ok());
    prm.done = ctx->overlap_done + ctx->overlap.slot();
    ctx->overlap_launched();
    return ok;
}

template <int N> int launch_fft(hzsdr_ctx *ctx, int dir, const float2 *src, float2 *dst, size_t batch, const float2 *tw);
template <int N> int launch_convolve(hzsdr_ctx *ctx, const float2 *src, float2 *dst, size_t nblocks, const float2 *tw, const float2 *H, size_t src_stride, size_t h_stride);
template <int N> int launch_chain(hzsdr_ctx *ctx, int fmt, const ChainParams &prm, const NcoTable &nco);
// chain1024.cu: warp-per-block specialisation for N = 1024 (prm.tw = the [31][32] table below)
int launch_chain1024(hzsdr_ctx *ctx, int fmt, const ChainParams &prm, const NcoTable &nco);
constexpr int kChain1024TableLen = 32 * 32 + 15 * 32 + 8 * 32;
void chain1024_twiddles(float2 *host_out /* kChain1024TableLen */);
void chain1024_split_twiddles(float2 *host_out /* 32*32 */, uint64_t dp_nom, float scale);
// chain16k.cu: CTA-per-block specialisation for N = 16384 with a decimation factor that is a multiple of 16
int launch_chain16k(hzsdr_ctx *ctx, int fmt, const ChainParams &prm, const NcoTable &nco);
void chain16k_twiddles(float2 *tw2 /* 31*32 */, float2 *tw3 /* 15*1024 */);
void chain16k_permute_filter(const float2 *H /* 16384 */, float2 *Hp /* 16384 */);
// chaink.cu: one CTA of K warps per block for N = K * 1024, K = 2, 4, 8 (prm.tw = [32][32] W_1024 table,
// prm.tw3 = [K-1][1024] W_N^{m k1}, prm.tw1k = the filter by sub-convolution; both from chaink_tables)
int launch_chaink(hzsdr_ctx *ctx, int fmt, int k, const ChainParams &prm, const NcoTable &nco);
void chaink_tables(int k, const float2 *H /* N */, float2 *twn /* (K-1)*1024 */, float2 *hp /* N */);
// one launch over prm.nstreams streams of prm.nblocks blocks each, described by prm.streams (device memory)
int launch_chain1024_batch(hzsdr_ctx *ctx, int fmt, const ChainParams &prm, const BatchTable *tbl, bool may_overlap);

// bigfft.cu: 2^15 .. 2^20 points as N1 x N2 through a scratch buffer (dir: FFT_FWD / FFT_BWD)
bool bigfft_len_ok(size_t n);
void bigfft_factors(size_t n, int *n1, int *n2);
int launch_bigfft(hzsdr_ctx *ctx, size_t n, int dir, const float2 *src, float2 *dst, float2 *scratch, size_t batch,
                  const float2 *tw1, const float2 *tw2);

// fft.cu: a transform of any supported length (direction: HZSDR_FFT_FORWARD / HZSDR_FFT_BACKWARD)
int fft_any(hzsdr_ctx *ctx, size_t n, int direction, const float2 *src, float2 *dst, size_t batch);

#ifdef HZ_FFT_N
#ifndef HZ_P16_RESIDENT
#define HZ_P16_RESIDENT 512
#endif
// at most 128 registers per thread (>= 512 resident threads per SM): without the cap several lengths
// compile to 170-250 registers and run 8 warps per SM
template <int N>
struct ChainCta {
    static constexpr int resident = FftCfg<N>::P <= 16 ? HZ_P16_RESIDENT : 512;  // threads per SM the register budget is cut for
    static constexpr int min_ctas = resident / FftCta<N>::threads >= 1 ? resident / FftCta<N>::threads : 1;
};

// first-pass gather pattern from global memory: v[i*R1 + r] = x[(t + T*i) + r*N/R1]
template <int N, int P, int R1>
__device__ __forceinline__ void load_first_pass(float2 (&v)[P], const float2 *__restrict__ x, int t, bool active) {
    constexpr int T = N / P;
    static_for<P / R1>([&](auto I) {
        constexpr int i = decltype(I)::value;
        static_for<R1>([&](auto RR) {
            constexpr int r = decltype(RR)::value;
            v[i * R1 + r] = active ? x[(t + T * i) + r * (N / R1)] : make_float2(0.f, 0.f);
        });
    });
}

// =================================================================================================
// K6a  batched FFT (fft.Plan.Transform).  16 B/sample of HBM traffic, 5*N*log2(N) flop/transform.
// =================================================================================================
template <int N, int DIR>
__global__ void __launch_bounds__(FftCta<N>::threads, ChainCta<N>::min_ctas) k_fft(const float2 *__restrict__ src, float2 *__restrict__ dst,
                                                             uint32_t batch, const float2 *__restrict__ tw) {
    using C = FftCfg<N>;
    constexpr int P = C::P, T = C::T, F = FftCta<N>::F, RL = C::RL;
    extern __shared__ float2 smem[];
    const int f = threadIdx.x / T, t = threadIdx.x % T;
    float2 *sm = smem + (size_t)f * smem_elems(N);
    for (uint32_t base = blockIdx.x * F; base < batch; base += gridDim.x * F) {
        const uint32_t b = base + f;
        const bool active = b < batch;
        float2 v[P];
        load_first_pass<N, P, C::R1>(v, src + (size_t)b * N, t, active);
        fft_regs<N, P, C::R1, C::R2, C::R3, DIR>(v, sm, tw, t);
        if (active) {
            float2 *y = dst + (size_t)b * N;
            constexpr int NS = N / RL;
            static_for<P / RL>([&](auto I) {
                constexpr int i = decltype(I)::value;
                static_for<RL>([&](auto QQ) {
                    constexpr int q = decltype(QQ)::value;
                    y[pass_out_index<N, P, RL, NS>(t, i, q)] = v[i * RL + bitrev(q, ilog2(RL))];
                });
            });
        }
    }
}

// after a forward transform: multiply by the filter and re-order for the inverse's first pass
//   w[i*RL + r] = X[j + r*N/RL] * H[j + r*N/RL],   X[...] sits in v[i*RL + bitrev(r)]
template <int N, int P, int RL>
__device__ __forceinline__ void spectrum_multiply(float2 (&v)[P], const float2 *__restrict__ H, int t) {
    constexpr int T = N / P;
    float2 w[P];
    static_for<P / RL>([&](auto I) {
        constexpr int i = decltype(I)::value;
        const int j = t + T * i;
        static_for<RL>([&](auto RR) {
            constexpr int r = decltype(RR)::value;
            const float2 h = __ldg(H + j + r * (N / RL));
            w[i * RL + r] = cmul(v[i * RL + bitrev(r, ilog2(RL))], h);
        });
    });
    static_for<P>([&](auto I) { v[decltype(I)::value] = w[decltype(I)::value]; });
}

// inverse transform with the pass order reversed, so that its first radix equals the forward's last
template <int N>
__device__ __forceinline__ void ifft_regs_reversed(float2 (&v)[FftCfg<N>::P], float2 *sm, const float2 *__restrict__ tw,
                                                   int t) {
    using C = FftCfg<N>;
    if constexpr (C::R3 > 1)
        fft_regs<N, C::P, C::R3, C::R2, C::R1, FFT_BWD>(v, sm, tw, t);
    else if constexpr (C::R2 > 1)
        fft_regs<N, C::P, C::R2, C::R1, 1, FFT_BWD>(v, sm, tw, t);
    else
        fft_regs<N, C::P, C::R1, 1, 1, FFT_BWD>(v, sm, tw, t);
}

// =================================================================================================
// K6b  ConvolveFreq over consecutive blocks (stream/convolution.go:62-80): 16 B/sample of HBM,
//      2 FFTs + 6N flop per block, spectrum stays in registers.
// =================================================================================================
template <int N>
__global__ void __launch_bounds__(FftCta<N>::threads, ChainCta<N>::min_ctas) k_convolve(const float2 *__restrict__ src, float2 *__restrict__ dst,
                                                                  uint32_t nblocks, const float2 *__restrict__ tw,
                                                                  const float2 *__restrict__ H, uint32_t src_stride, uint32_t h_stride) {
    using C = FftCfg<N>;
    constexpr int P = C::P, T = C::T, F = FftCta<N>::F, R1 = C::R1;
    extern __shared__ float2 smem[];
    const int f = threadIdx.x / T, t = threadIdx.x % T;
    float2 *sm = smem + (size_t)f * smem_elems(N);
    for (uint32_t base = blockIdx.x * F; base < nblocks; base += gridDim.x * F) {
        const uint32_t b = base + f;
        const bool active = b < nblocks;
        float2 v[P];
        load_first_pass<N, P, R1>(v, src + (size_t)b * src_stride, t, active);  // stride < N: overlap-save windows
        fft_regs<N, P, C::R1, C::R2, C::R3, FFT_FWD>(v, sm, tw, t);
        spectrum_multiply<N, P, C::RL>(v, H + (active ? (size_t)b * h_stride : 0), t);  // h_stride = N: one spectrum per block
        ifft_regs_reversed<N>(v, sm, tw, t);
        if (active) {
            float2 *y = dst + (size_t)b * N;
            // the reversed inverse ends with radix R1, Ns = N/R1: value q of item i is y[j + q*N/R1]
            static_for<P / R1>([&](auto I) {
                constexpr int i = decltype(I)::value;
                static_for<R1>([&](auto QQ) {
                    constexpr int q = decltype(QQ)::value;
                    y[(t + T * i) + q * (N / R1)] = v[i * R1 + bitrev(q, ilog2(R1))];
                });
            });
        }
    }
}

// =================================================================================================
// Fused chain: raw -> toC64 -> NCO mix -> FFT_N -> xH -> IFFT_N -> decimate.
// Algorithmic HBM bytes per input sample: raw bytes + 8 * kept/total (2.80 B for i8 / N=1024 / D=10).
// =================================================================================================

// floor(x / d) for x < 2^24: float estimate (never too large: inv_d is rounded down) + 1 fix-up
__device__ __forceinline__ uint32_t chain_udiv(uint32_t x, uint32_t d, float inv_d) {
    uint32_t q = __float2uint_rz(__uint2float_rz(x) * inv_d);
    if (x - q * d >= d) q++;
    return q;
}

// raw word of sample j, Pluto LSB shift applied
template <int FMT>
__device__ __forceinline__ uint32_t chain_load_raw(const uint8_t *__restrict__ src, uint32_t j, int lsb_shift) {
    if constexpr (FMT == HZSDR_FORMAT_I16) {
        uint32_t w = __ldg(reinterpret_cast<const uint32_t *>(src) + j);
        if (lsb_shift) w = ((w & 0xffff0000u) << lsb_shift) | (((w & 0xffffu) << lsb_shift) & 0xffffu);
        return w;
    } else {
        return (uint32_t)__ldg(reinterpret_cast<const uint16_t *>(src) + j);
    }
}

template <int N, int FMT>
__global__ void __launch_bounds__(FftCta<N>::threads, ChainCta<N>::min_ctas) k_chain(const __grid_constant__ ChainParams prm,
                                                                                     const __grid_constant__ NcoTable nco) {
    using C = FftCfg<N>;
    constexpr int P = C::P, T = C::T, F = FftCta<N>::F, R1 = C::R1;
    extern __shared__ float2 smem[];
    const int f = threadIdx.x / T, t = threadIdx.x % T;
    float2 *sm = smem + (size_t)f * smem_elems(N);
    const float sc = RawTraits<FMT>::scale();
    for (uint32_t base = blockIdx.x * F; base < prm.nblocks; base += gridDim.x * F) {
        const uint32_t b = base + f;
        const bool active = b < prm.nblocks;
        const uint32_t s0 = b * N;  // launch-relative index of the block's first sample
        float2 v[P];
        // ---- Convert + Shift: a coalesced walk over the block (thread t takes samples t + T k) into
        // the exchange buffer, as a compact loop.  Inside one accumulator segment the phase is linear in
        // the sample index, so the rotation advances by a constant complex step and is re-anchored with
        // an exact evaluation every 4 samples; a block that straddles segments evaluates every sample.
        fft_sync<T>();  // the previous block's last exchange has been read
        if (active) {
            const int si = nco_find(nco, s0);
            const uint32_t seg_j0 = nco.seg[si].j0, seg_end = seg_j0 + nco.seg[si].count;
            if (s0 + (uint32_t)N <= seg_end) {
                const uint64_t dp = nco.seg[si].dp;
                const uint64_t ph0 = nco.seg[si].p0 + (uint64_t)(s0 + t - seg_j0 + 1) * dp;
                const float2 step = nco_rot((uint64_t)T * dp);
                float2 rot = make_float2(0.f, 0.f);
#pragma unroll 4
                for (int k = 0; k < P; ++k) {
                    if ((k & 3) == 0) {
                        rot = nco_rot(ph0 + (uint64_t)(T * k) * dp);
                        rot = mul2(rot, make_float2(sc, sc));
                    }
                    const float2 x = RawTraits<FMT>::unscaled2(chain_load_raw<FMT>(prm.src, s0 + t + T * k, prm.lsb_shift));
                    sm[smem_pad(t + T * k)] = cmul(x, rot);
                    rot = cmul(rot, step);
                }
            } else {
                NcoCursor cur;
#pragma unroll 1
                for (int k = 0; k < P; ++k) {
                    const uint32_t j = s0 + t + T * k;
                    cur.seek(nco, j);
                    const float2 rot = mul2(nco_rot(cur.phase(j)), make_float2(sc, sc));
                    sm[smem_pad(t + T * k)] = cmul(RawTraits<FMT>::unscaled2(chain_load_raw<FMT>(prm.src, j, prm.lsb_shift)), rot);
                }
            }
        }
        fft_sync<T>();
        if (active) {
            pass_gather<N, P, R1>(v, sm, t);
        } else {
            static_for<P>([&](auto I) { v[decltype(I)::value] = make_float2(0.f, 0.f); });
        }
        // ---- ConvolutionReader: FFT, xH, IFFT ----
        fft_regs<N, P, C::R1, C::R2, C::R3, FFT_FWD>(v, sm, prm.tw, t);
        spectrum_multiply<N, P, C::RL>(v, prm.H, t);
        ifft_regs_reversed<N>(v, sm, prm.tw, t);
        // ---- DecimateReader: keep z[q*DB + D*i], i < M ----
        if constexpr (T > 1) {
            if ((uint32_t)N <= (1u << prm.db_log2)) {  // uniform: the block lies inside one decimate block
                // z in natural order through the exchange buffer, then the transform's threads copy the kept
                // samples out, coalesced (testing every sample in registers costs ~12 instructions each)
                fft_sync<T>();  // the inverse's last gather is done
                if (active) {
                    static_for<P / R1>([&](auto I) {
                        constexpr int i = decltype(I)::value;
                        static_for<R1>([&](auto QQ) {
                            constexpr int q = decltype(QQ)::value;
                            sm[smem_pad((t + T * i) + q * (N / R1))] = v[i * R1 + bitrev(q, ilog2(R1))];
                        });
                    });
                }
                fft_sync<T>();
                if (active) {
                    const uint32_t g0 = prm.z0 + s0, p0 = g0 & ((1u << prm.db_log2) - 1u);
                    const uint32_t o0 = chain_udiv(p0 + prm.D - 1u, prm.D, prm.inv_d);  // first kept output at or after the block's start
                    const uint32_t pos0 = o0 * prm.D - p0;
                    uint32_t cnt = 0;
                    if (pos0 < (uint32_t)N && o0 < prm.M) {
                        cnt = chain_udiv((uint32_t)N - 1u - pos0, prm.D, prm.inv_d) + 1u;
                        if (cnt > prm.M - o0) cnt = prm.M - o0;
                    }
                    float2 *out = prm.dst + (size_t)(g0 >> prm.db_log2) * prm.M + o0;
                    for (uint32_t k = t; k < cnt; k += T) out[k] = sm[smem_pad((int)(pos0 + k * prm.D))];
                }
                continue;  // (the barrier at the top of the loop covers the buffer's reuse)
            }
        }
        if (active) {
            const uint32_t db_mask = (1u << prm.db_log2) - 1u;
            static_for<P / R1>([&](auto I) {
                constexpr int i = decltype(I)::value;
                static_for<R1>([&](auto QQ) {
                    constexpr int q = decltype(QQ)::value;
                    const uint32_t g = prm.z0 + s0 + (t + T * i) + q * (N / R1);
                    const uint32_t p = g & db_mask;
                    uint32_t o = __float2uint_rz(__uint2float_rz(p) * prm.inv_d);
                    uint32_t rem = p - o * prm.D;
                    if (rem >= prm.D) {
                        rem -= prm.D;
                        o++;
                    }
                    if (rem == 0 && o < prm.M) {
                        const size_t out = (size_t)(g >> prm.db_log2) * prm.M + o;
                        prm.dst[out] = v[i * R1 + bitrev(q, ilog2(R1))];
                    }
                });
            });
        }
    }
}

template <int N>
static int set_smem_attr(const void *fn) {
    if (FftCta<N>::smem_bytes > 48 * 1024)
        HZ_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FftCta<N>::smem_bytes));
    return HZSDR_OK;
}

template <int N>
static int fft_grid(const hzsdr_ctx *ctx, size_t batch) {
    constexpr int F = FftCta<N>::F;
    size_t need = (batch + F - 1) / F;
    // resident CTAs per SM, bounded by shared memory and by ~2048 threads
    size_t per_sm = 2048 / FftCta<N>::threads;
    const size_t by_smem = FftCta<N>::smem_bytes ? (200 * 1024) / FftCta<N>::smem_bytes : per_sm;
    if (by_smem < per_sm) per_sm = by_smem;
    if (per_sm < 1) per_sm = 1;
    size_t cap = (size_t)ctx->sm_count * per_sm;
    return (int)(need < cap ? need : cap);
}

template <int N>
int launch_fft(hzsdr_ctx *ctx, int dir, const float2 *src, float2 *dst, size_t batch, const float2 *tw) {
    const int grid = fft_grid<N>(ctx, batch);
    const size_t smem = FftCfg<N>::T > 1 ? FftCta<N>::smem_bytes : 0;
    if (dir == HZSDR_FFT_FORWARD) {
        int rc = set_smem_attr<N>((const void *)k_fft<N, FFT_FWD>);
        if (rc) return rc;
        k_fft<N, FFT_FWD><<<grid, FftCta<N>::threads, smem, ctx->stream>>>(src, dst, (uint32_t)batch, tw);
    } else {
        int rc = set_smem_attr<N>((const void *)k_fft<N, FFT_BWD>);
        if (rc) return rc;
        k_fft<N, FFT_BWD><<<grid, FftCta<N>::threads, smem, ctx->stream>>>(src, dst, (uint32_t)batch, tw);
    }
    HZ_CHECK_LAUNCH();
    return HZSDR_OK;
}

template <int N>
int launch_convolve(hzsdr_ctx *ctx, const float2 *src, float2 *dst, size_t nblocks, const float2 *tw,
                           const float2 *H, size_t src_stride, size_t h_stride) {
    int rc = set_smem_attr<N>((const void *)k_convolve<N>);
    if (rc) return rc;
    const size_t smem = FftCfg<N>::T > 1 ? FftCta<N>::smem_bytes : 0;
    k_convolve<N><<<fft_grid<N>(ctx, nblocks), FftCta<N>::threads, smem, ctx->stream>>>(src, dst, (uint32_t)nblocks, tw, H, (uint32_t)src_stride, (uint32_t)h_stride);
    HZ_CHECK_LAUNCH();
    return HZSDR_OK;
}

template <int N>
int launch_chain(hzsdr_ctx *ctx, int fmt, const ChainParams &prm, const NcoTable &nco) {
    const size_t smem = FftCfg<N>::T > 1 ? FftCta<N>::smem_bytes : 0;
    const int grid = fft_grid<N>(ctx, prm.nblocks);
    int rc;
    switch (fmt) {
        case HZSDR_FORMAT_U8:
            if ((rc = set_smem_attr<N>((const void *)k_chain<N, HZSDR_FORMAT_U8>))) return rc;
            k_chain<N, HZSDR_FORMAT_U8><<<grid, FftCta<N>::threads, smem, ctx->stream>>>(prm, nco);
            break;
        case HZSDR_FORMAT_I8:
            if ((rc = set_smem_attr<N>((const void *)k_chain<N, HZSDR_FORMAT_I8>))) return rc;
            k_chain<N, HZSDR_FORMAT_I8><<<grid, FftCta<N>::threads, smem, ctx->stream>>>(prm, nco);
            break;
        default:
            if ((rc = set_smem_attr<N>((const void *)k_chain<N, HZSDR_FORMAT_I16>))) return rc;
            k_chain<N, HZSDR_FORMAT_I16><<<grid, FftCta<N>::threads, smem, ctx->stream>>>(prm, nco);
            break;
    }
    HZ_CHECK_LAUNCH();
    return HZSDR_OK;
}


template int launch_fft<HZ_FFT_N>(hzsdr_ctx *, int, const float2 *, float2 *, size_t, const float2 *);
template int launch_convolve<HZ_FFT_N>(hzsdr_ctx *, const float2 *, float2 *, size_t, const float2 *, const float2 *, size_t, size_t);
template int launch_chain<HZ_FFT_N>(hzsdr_ctx *, int, const ChainParams &, const NcoTable &);
#endif  // HZ_FFT_N

}  // namespace hz
