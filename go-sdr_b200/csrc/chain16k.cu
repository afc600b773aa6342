// chain16k.cu -- the fused chain specialised for ConvolutionReader blocks of N = 16384 with a
// decimation factor that is a multiple of 16 (BASELINE config 3: 4095-tap lowpass as a 16384-bin
// frequency-domain filter, Decimate x16).
//
// One CTA of 512 threads owns a 16384-sample block (32 points per thread), one CTA per SM:
//   prefetch  the NEXT block's raw samples stream into shared memory with cp.async while this block
//             computes -- with a single resident CTA there is nobody else to hide HBM latency.
//   stage A   raw -> float -> NCO mix in the first pass's stride-512 pattern; one sincos per thread
//             per block, the other 31 rotations from two tiny per-block tables (nco.cuh).
//   forward   16384 = 32 x 32 x 16 Stockham; the two radix-32 passes share one instruction stream.
//   x H, fold the last pass leaves a thread holding X[j + 1024 q], q < 16 -- exactly the 16 bins
//             that alias onto output bin j when only every 16th output sample is kept:
//                 z[16 m] = IDFT_1024( sum_q X[j + 1024 q] H[j + 1024 q] )[m]
//             so the filter multiply and the fold are 16 complex FMAs in registers and the inverse
//             shrinks from 16384 to 1024 points.
//   inverse   1024 = 8 x 8 x 8 x 2 by 4 of the 16 warps (named barrier), forward code on re/im-
//             swapped data; the other 12 warps run ahead into the next block.
//   stage C   the same 4 warps copy the kept samples z[q*32768 + D*i] out.
//
// Algorithmic HBM bytes per input sample: raw bytes + 8/D (4.5 B for i16, D = 16).
#include "common.cuh"
#include "fft.cuh"
#include "fft_kernels.cuh"
#include "nco.cuh"

namespace hz {

constexpr int kC16Threads = 512;
constexpr int kC16N = 16384;
constexpr int kC16InvThreads = 128;

__device__ __forceinline__ int xpad(int a) { return a + (a >> 5); }  // exchange buffer, 64-bit accesses
__device__ __forceinline__ int ypad(int a) { return a + (a >> 3); }  // inverse workspace, stride-8 scatter

struct Chain16kSmem {
    float2 x[kC16N + kC16N / 32 + 8];  // forward exchange buffer (padded)
    float2 tw2[31][32];                // W_1024^{r*l}: twiddles of the second radix-32 pass
    float2 y[1024 + 128 + 8];          // folded spectrum -> inverse workspace (padded)
    float2 rot[16];                    // NCO step tables of the current block
    uint32_t raw[kC16N];               // the block's raw samples (prefetched); 2-byte formats use half
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

template <int FMT>
__device__ __forceinline__ void c16_prefetch(Chain16kSmem &S, const uint8_t *__restrict__ src, uint32_t block) {
    constexpr int kBytes = kC16N * RawTraits<FMT>::bytes;
    const uint8_t *g = src + (size_t)block * kBytes;
    uint8_t *s = reinterpret_cast<uint8_t *>(S.raw);
#pragma unroll
    for (int k = 0; k < kBytes / 16 / kC16Threads; k++) {
        const int off = (k * kC16Threads + threadIdx.x) * 16;
        cp_async16(s + off, g + off);
    }
    cp_async_commit();
}

template <int FMT>
__device__ __forceinline__ uint32_t c16_raw(const Chain16kSmem &S, int idx, int lsb_shift) {
    if constexpr (RawTraits<FMT>::bytes == 4) {
        uint32_t w = S.raw[idx];
        if (lsb_shift) w = ((w & 0xffff0000u) << lsb_shift) | (((w & 0xffffu) << lsb_shift) & 0xffffu);
        return w;
    } else {
        return (uint32_t) reinterpret_cast<const uint16_t *>(S.raw)[idx];
    }
}

template <int FMT>
__global__ void __launch_bounds__(kC16Threads, 1) k_chain16k(const __grid_constant__ ChainParams prm,
                                                              const __grid_constant__ NcoTable nco) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Chain16kSmem &S = *reinterpret_cast<Chain16kSmem *>(smem_raw);
    const int t = threadIdx.x;

    asm volatile("griddepcontrol.launch_dependents;");  // see chain1024.cu: consecutive buffers are independent

    uint32_t b = blockIdx.x;
    if (b < prm.nblocks) c16_prefetch<FMT>(S, prm.src, b);
    for (int i = t; i < 31 * 32; i += kC16Threads) (&S.tw2[0][0])[i] = __ldg(prm.tw + i);

    const uint32_t db_mask = (1u << prm.db_log2) - 1u;
    const float sc = RawTraits<FMT>::scale();

    for (; b < prm.nblocks; b += gridDim.x) {
        const uint32_t s0 = b * (uint32_t)kC16N;
        float2 v[32];

        // ------------------------------------------------------------------ stage A
        const int si = nco_find(nco, s0);
        const uint32_t seg_j0 = nco.seg[si].j0, seg_end = seg_j0 + nco.seg[si].count;
        const uint64_t seg_p0 = nco.seg[si].p0, seg_dp = nco.seg[si].dp;
        const bool fast = s0 + (uint32_t)kC16N <= seg_end;
        if (fast) {  // phase(s0 + t + 512 r) = ph_t + 512 r dP, r = 8a + bb
            if (t < 8)
                S.rot[t] = nco_rot((uint64_t)(512u * t) * seg_dp);
            else if (t < 12)
                S.rot[t] = nco_rot((uint64_t)(4096u * (t - 8)) * seg_dp);
        }
        cp_async_wait_all();
        __syncthreads();  // raw block landed, rot tables (and, first time, tw2) visible
        if (fast) {
            float2 r0 = nco_rot(seg_p0 + (uint64_t)(s0 + t - seg_j0 + 1) * seg_dp);
            r0.x *= sc;
            r0.y *= sc;
            static_for<4>([&](auto AA) {
                constexpr int a = decltype(AA)::value;
                const float2 ra = a == 0 ? r0 : cmul(r0, S.rot[8 + a]);
                static_for<8>([&](auto BB) {
                    constexpr int bb = decltype(BB)::value;
                    const float2 rot = bb == 0 ? ra : cmul(ra, S.rot[bb]);
                    v[8 * a + bb] = cmul(RawTraits<FMT>::unscaled(c16_raw<FMT>(S, t + 512 * (8 * a + bb), prm.lsb_shift)), rot);
                });
            });
        } else {  // block straddles accumulator segments: compact per-sample loop through the exchange buffer
            NcoCursor cur;
#pragma unroll 1
            for (int r = 0; r < 32; ++r) {
                const uint32_t j = s0 + t + 512u * r;
                cur.seek(nco, j);
                float2 rot = nco_rot(cur.phase(j));
                rot.x *= sc;
                rot.y *= sc;
                S.x[xpad(t + 512 * r)] = cmul(RawTraits<FMT>::unscaled(c16_raw<FMT>(S, t + 512 * r, prm.lsb_shift)), rot);
            }
            static_for<32>([&](auto RR) {  // own elements only: no barrier needed
                constexpr int r = decltype(RR)::value;
                v[r] = S.x[xpad(t + 512 * r)];
            });
        }

        // ------------------------------------------------------------------ forward: two radix-32 passes
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
            fft_reg<32, FFT_FWD, 0, 32>(v);
            __syncthreads();  // everyone is done reading x (and, pass 0, the raw block)
            if (pass == 0) {
                // raw is free: start streaming the next block in behind the computation
                if (b + gridDim.x < prm.nblocks) c16_prefetch<FMT>(S, prm.src, b + gridDim.x);
                static_for<32>([&](auto QQ) {  // Ns = 1: index 32 t + q, padded = 33 t + q
                    constexpr int q = decltype(QQ)::value;
                    S.x[33 * t + q] = v[bitrev(q, 5)];
                });
            } else {
                const int base = (t >> 5) * 1056 + (t & 31);  // Ns = 32: index 1024 (t/32) + t%32 + 32 q, padded
                static_for<32>([&](auto QQ) {
                    constexpr int q = decltype(QQ)::value;
                    S.x[base + 33 * q] = v[bitrev(q, 5)];
                });
            }
            __syncthreads();
            if (pass == 0) {
                const int gbase = t + (t >> 5);  // element t + 512 r, padded = t + t/32 + 528 r
                static_for<32>([&](auto RR) {
                    constexpr int r = decltype(RR)::value;
                    v[r] = S.x[gbase + 528 * r];
                });
                static_for<31>([&](auto RR) {  // W_1024^{r (t mod 32)}
                    constexpr int r = decltype(RR)::value + 1;
                    const float2 w = S.tw2[r - 1][t & 31];
                    v[r] = tw_mul<FFT_FWD>(v[r], w.x, w.y);
                });
            }
        }

        // ------------------------------------------------------------------ third pass (radix 16, Ns = 1024), x H, fold
        // items j = t and t + 512; item j gathers x[j + 1024 r], twiddle W_16384^{r j}
        float2 yf[2];
#pragma unroll 1
        for (int i = 0; i < 2; ++i) {
            const int j = t + 512 * i;
            const int gbase = j + (j >> 5);  // padded index of j + 1024 r = j + j/32 + 1056 r
            float2 w[16];
            static_for<16>([&](auto RR) {
                constexpr int r = decltype(RR)::value;
                w[r] = S.x[gbase + 1056 * r];
            });
            static_for<15>([&](auto RR) {
                constexpr int r = decltype(RR)::value + 1;
                const float2 tw = __ldg(prm.tw3 + (r - 1) * 1024 + j);
                w[r] = tw_mul<FFT_FWD>(w[r], tw.x, tw.y);
            });
            fft_reg<16, FFT_FWD, 0, 16>(w);
            // X[j + 1024 q] = w[bitrev(q)];  Yf[j] = sum_q X[j + 1024 q] H[j + 1024 q]
            float2 acc = make_float2(0.f, 0.f);
            static_for<16>([&](auto QQ) {
                constexpr int q = decltype(QQ)::value;
                const float2 h = __ldg(prm.H + j + 1024 * q);
                const float2 x = w[bitrev(q, 4)];
                acc.x = fmaf(x.x, h.x, acc.x);
                acc.x = fmaf(-x.y, h.y, acc.x);
                acc.y = fmaf(x.x, h.y, acc.y);
                acc.y = fmaf(x.y, h.x, acc.y);
            });
            yf[i] = acc;
        }
        // swapped, so that forward passes compute the inverse (IDFT(Y) = swap(DFT(swap(Y))))
        S.y[ypad(t)] = make_float2(yf[0].y, yf[0].x);
        S.y[ypad(t + 512)] = make_float2(yf[1].y, yf[1].x);
        __syncthreads();

        // ------------------------------------------------------------------ inverse 1024 = 8 x 8 x 8 x 2 + stage C (4 warps)
        if (t < kC16InvThreads) {
            float2 u[8];
            // three radix-8 passes, Ns = 1, 8, 64; item j = t gathers y[j + 128 r]
#pragma unroll 1
            for (int pass = 0; pass < 3; ++pass) {
                const int ns_log2 = 3 * pass, ns = 1 << ns_log2;
                const int gb = t + (t >> 3);  // padded index of t + 128 r = t + t/8 + 144 r
                static_for<8>([&](auto RR) {
                    constexpr int r = decltype(RR)::value;
                    u[r] = S.y[gb + 144 * r];
                });
                if (pass > 0) {  // W_{8 ns}^{r (t mod ns)} = W_1024^{r (t mod ns) (128 / ns)}
                    const int k = (t & (ns - 1)) << (7 - ns_log2);
                    static_for<7>([&](auto RR) {
                        constexpr int r = decltype(RR)::value + 1;
                        const float2 tw = __ldg(prm.tw1k + r * k);
                        u[r] = tw_mul<FFT_FWD>(u[r], tw.x, tw.y);
                    });
                }
                fft_reg<8, FFT_FWD, 0, 8>(u);
                named_bar_sync(1, kC16InvThreads);  // the 128 inverse threads are done reading y
                const int obase = ((t >> ns_log2) << (ns_log2 + 3)) + (t & (ns - 1));
                static_for<8>([&](auto QQ) {
                    constexpr int q = decltype(QQ)::value;
                    S.y[ypad(obase + (q << ns_log2))] = u[bitrev(q, 3)];
                });
                named_bar_sync(1, kC16InvThreads);
            }
            // radix-2 pass, Ns = 512: items j = t + 128 i: (y[j], y[j+512] W_1024^j) -> y[j], y[j+512]
            static_for<4>([&](auto II) {
                constexpr int i = decltype(II)::value;
                const int j = t + 128 * i;
                u[2 * i] = S.y[ypad(j)];
                u[2 * i + 1] = S.y[ypad(j + 512)];
            });
            named_bar_sync(1, kC16InvThreads);
            static_for<4>([&](auto II) {
                constexpr int i = decltype(II)::value;
                const int j = t + 128 * i;
                const float2 tw = __ldg(prm.tw1k + j);
                const float2 a = u[2 * i], bq = tw_mul<FFT_FWD>(u[2 * i + 1], tw.x, tw.y);
                S.y[ypad(j)] = make_float2(a.x + bq.x, a.y + bq.y);
                S.y[ypad(j + 512)] = make_float2(a.x - bq.x, a.y - bq.y);
            });
            named_bar_sync(1, kC16InvThreads);

            // stage C: y[m] = swap(z[16 m]).  Keep z[g], (g mod DB) = D*i, i < M  (D is a multiple of 16)
            const uint32_t g0 = prm.z0 + s0;
            const uint32_t p0 = g0 & db_mask;
            const uint32_t o0 = (p0 + prm.D - 1u) / prm.D;
            const uint32_t pos0 = o0 * prm.D - p0;
            uint32_t cnt = 0;
            if (pos0 < (uint32_t)kC16N && o0 < prm.M) {
                cnt = ((uint32_t)kC16N - 1u - pos0) / prm.D + 1u;
                if (cnt > prm.M - o0) cnt = prm.M - o0;
            }
            float2 *out = prm.dst + (size_t)(g0 >> prm.db_log2) * prm.M + o0;
            for (uint32_t k = t; k < cnt; k += kC16InvThreads) {
                const uint32_t m = (pos0 + k * prm.D) >> 4;
                const float2 z = S.y[ypad((int)m)];
                out[k] = make_float2(z.y, z.x);
            }
        }
        // No barrier here: the next write to y is after three more __syncthreads, which the inverse
        // warps only reach once they are done with it.
    }
    cp_async_wait_all();
}

template <int FMT>
static int launch16(hzsdr_ctx *ctx, const ChainParams &prm, const NcoTable &nco) {
    static bool attr_set = false;
    const size_t smem = sizeof(Chain16kSmem);
    if (!attr_set) {
        HZ_CUDA(cudaFuncSetAttribute((const void *)k_chain16k<FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const int grid = (int)(prm.nblocks < (uint32_t)ctx->sm_count ? prm.nblocks : (uint32_t)ctx->sm_count);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kC16Threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    HZ_CUDA(cudaLaunchKernelEx(&cfg, k_chain16k<FMT>, prm, nco));
    return HZSDR_OK;
}

// prm.tw = [31][32] W_1024^{r l}; prm.tw3 = [15][1024] W_16384^{r j}; prm.tw1k = W_1024^m, m < 1024
int launch_chain16k(hzsdr_ctx *ctx, int fmt, const ChainParams &prm, const NcoTable &nco) {
    switch (fmt) {
        case HZSDR_FORMAT_U8: return launch16<HZSDR_FORMAT_U8>(ctx, prm, nco);
        case HZSDR_FORMAT_I8: return launch16<HZSDR_FORMAT_I8>(ctx, prm, nco);
        default: return launch16<HZSDR_FORMAT_I16>(ctx, prm, nco);
    }
}

void chain16k_twiddles(float2 *tw2 /* 31*32 */, float2 *tw3 /* 15*1024 */, float2 *tw1k /* 1024 */) {
    auto w = [](double num, double den) {
        const double a = 2.0 * M_PI * num / den;
        return make_float2((float)cos(a), (float)sin(a));
    };
    for (int r = 1; r < 32; r++)
        for (int l = 0; l < 32; l++) tw2[(r - 1) * 32 + l] = w(r * l, 1024.0);
    for (int r = 1; r < 16; r++)
        for (int j = 0; j < 1024; j++) tw3[(r - 1) * 1024 + j] = w((double)r * j, 16384.0);
    for (int m = 0; m < 1024; m++) tw1k[m] = w(m, 1024.0);
}

}  // namespace hz
