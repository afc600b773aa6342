// chain16k.cu -- the fused chain specialised for ConvolutionReader blocks of N = 16384 with a
// decimation factor that is a multiple of 16 (BASELINE config 3: 4095-tap lowpass as a 16384-bin
// frequency-domain filter, Decimate x16).
//
// One CTA per SM owns a 16384-sample block at a time: 16 worker warps + 1 inverse warp.
//
//   16384 = 16 x 1024, decimation in frequency:
//       X[k + 16 q] = sum_m  W_16384^{m k} ( sum_c x[m + 1024 c] W_16^{c k} )  W_1024^{m q}
//   so after ONE radix-16 pass over the whole block (pass 1, every thread two butterflies, inputs
//   straight from the prefetched raw samples: convert + NCO mix fused in) the block falls apart
//   into 16 independent 1024-point transforms, one per worker warp, which run exactly like the
//   N = 1024 kernel (chain1024.cu): two radix-32 passes, exchange through the warp's own shared
//   memory region, __syncwarp only.  The workers meet at two CTA barriers per block instead of
//   six, and between them every warp is on its own schedule.
//
//   x H, fold  warp k ends with lane l holding X[k + 16 (l + 32 q)], q < 32.  Keeping every 16th
//       output sample aliases the bins j + 1024 c, c < 16, onto bin j of a 1024-point spectrum:
//           z[16 n] = IDFT_1024( sum_c X[j + 1024 c] H[j + 1024 c] )[n]
//       and those 16 bins all sit in ONE lane (q = p + 2 c): 32 complex FMAs in registers leave two
//       folded bins per lane.  The filter is stored permuted (chain16k_permute_filter) so that
//       these reads are coalesced.
//   inverse    the 17th warp turns the folded spectrum into the kept samples (a warp-local
//       1024-point transform on re/im-swapped data + the DecimateReader copy) while the workers
//       are already in the next block; folded spectra are double buffered.
//   prefetch   the NEXT block's raw samples stream into shared memory with cp.async behind the
//       computation -- with a single resident CTA there is nobody else to hide HBM latency.
//
// Barriers: 1 = workers (512 threads): raw samples landed, x free;  2 = workers + inverse warp
// (544): pass 1 scattered -- the inverse warp only *arrives* here, once per block, when the
// folded-spectrum buffer the workers will write next is free;  3, 4 = folded spectrum
// buffer 0 / 1 full (workers arrive, inverse warp waits).
//
// Algorithmic HBM bytes per input sample: raw bytes + 8/D (4.5 B for i16, D = 16).
//
// OS = true: OVERLAP-SAVE (BASELINE config 3 as worded; an extension -- the reference's ConvolutionReader
// is block-circular, stream/convolution.go:57-81).  The same pipeline runs on windows of 16384 samples
// that start every L = prm.os_hop samples (a multiple of 16, <= N - taps + 1; 12288 for 4095 taps), in
// coordinates that begin os_hist = N - L samples BEFORE the launch's first new sample: the first
// prm.os_head samples come from the chain's carried raw history (prm.hist), the rest from the buffer, and
// whatever lies beyond the buffer's end is zero-filled (a causal FIR's outputs do not depend on it).  Only
// window positions >= os_hist are free of circular wrap-around, so stage C keeps z[16 n] for n >= os_hist/16:
// L/16 of the 1024 folded outputs per window.  True linear convolution z[n] = sum_k h[k] y[n-k], history
// carried across calls; per new sample N/L = 1.33x the work of the block-circular form.
#include "common.cuh"
#include "fft.cuh"
#include "fft_kernels.cuh"
#include "nco.cuh"

namespace hz {

constexpr int kC16Workers = 512;  // 16 warps
constexpr int kC16Threads = 544;  // + the inverse warp
constexpr int kC16N = 16384;
constexpr int kBarWorkers = 1, kBarScatter = 2, kBarFull = 3;

__device__ __forceinline__ int xpad(int a) { return a + (a >> 5); }  // 1024-point regions, 64-bit accesses

struct Chain16kSmem {
    float2 x[16][1056];   // x[k]: DFT-16 output k of every m (padded m + m/32); then warp k's exchange space
    float2 tw2[31][32];   // W_1024^{r*l}: twiddles between the two radix-32 passes
    float2 y[2][1056];    // folded spectra (padded), double buffered; the inverse warp transforms in place
    float2 rot[24];       // NCO step tables: rot[c] = e^{i 1024 c dP}, c < 16
    uint32_t raw[kC16N];  // the block's raw samples (prefetched); 2-byte formats use half
};

__device__ __forceinline__ void bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

__device__ __forceinline__ void cp_async16_zfill(void *smem, const void *gmem, uint32_t src_bytes) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}

template <int FMT, bool OS>
__device__ __forceinline__ void c16_prefetch(Chain16kSmem &S, const ChainParams &prm, uint32_t block) {
    constexpr int kSb = RawTraits<FMT>::bytes;
    constexpr int kBytes = kC16N * kSb;
    uint8_t *s = reinterpret_cast<uint8_t *>(S.raw);
    if constexpr (!OS) {
        const uint8_t *g = prm.src + (size_t)block * kBytes;
#pragma unroll
        for (int k = 0; k < kBytes / 16 / kC16Workers; k++) {
            const int off = (k * kC16Workers + threadIdx.x) * 16;
            cp_async16(s + off, g + off);
        }
    } else {
        // window `block` = launch coordinates [block * hop, + 16384): history | buffer | zeros
        const uint32_t e0 = block * prm.os_hop;
#pragma unroll
        for (int k = 0; k < kBytes / 16 / kC16Workers; k++) {
            const int off = (k * kC16Workers + threadIdx.x) * 16;
            const uint32_t e = e0 + (uint32_t)off / kSb;  // first sample of the 16-byte chunk (boundaries are multiples of 8 samples)
            if (e < prm.os_head)
                cp_async16(s + off, prm.hist + (size_t)e * kSb);
            else if (e < prm.os_valid)
                cp_async16(s + off, prm.src + (size_t)e * kSb);
            else
                cp_async16_zfill(s + off, prm.hist, 0u);
        }
    }
    cp_async_commit();
}

// LSB: the Pluto LSB->MSB shift (iq_i16.go:103-111), compiled in only for chains that ask for it
template <int FMT, bool LSB>
__device__ __forceinline__ uint32_t c16_raw(const Chain16kSmem &S, int idx, int lsb_shift) {
    if constexpr (RawTraits<FMT>::bytes == 4) {
        uint32_t w = S.raw[idx];
        if constexpr (LSB) w = ((w & 0xffff0000u) << lsb_shift) | (((w & 0xffffu) << lsb_shift) & 0xffffu);
        return w;
    } else {
        return (uint32_t) reinterpret_cast<const uint16_t *>(S.raw)[idx];
    }
}

// the 17th warp: folded spectrum -> kept output samples, one block behind the workers
template <bool OS>
__device__ __forceinline__ void c16_inverse_warp(Chain16kSmem &S, const ChainParams &prm, uint32_t n_it) {
    const int lane = threadIdx.x & 31;
    const uint32_t db_mask = (1u << prm.db_log2) - 1u;
    // Arrival p on barrier 2 tells the workers that y[p & 1] is free (inverse(p - 2) is done).  A
    // hardware barrier has ONE counter, so the warp must never be two arrivals ahead: arrival p + 1
    // is issued right after "full(p)" completes, which implies that phase p of barrier 2 is over
    // and, the warp being sequential, that inverse(p - 1) is done.
    if (n_it) bar_arrive(kBarScatter, kC16Threads);
    uint32_t b = blockIdx.x;
    for (uint32_t it = 0; it < n_it; ++it, b += gridDim.x) {
        float2 *y = S.y[it & 1];
        bar_sync(kBarFull + (int)(it & 1u), kC16Threads);
        if (it + 1u < n_it) bar_arrive(kBarScatter, kC16Threads);  // y[(it + 1) & 1] is free
        // 1024 = 32 x 32 forward passes on swapped data: IDFT(Y) = swap(DFT(swap(Y)))
        float2 v[32];
        static_for<32>([&](auto RR) {
            constexpr int r = decltype(RR)::value;
            v[r] = y[lane + 33 * r];
        });
        fft_reg<32, FFT_FWD, 0, 32>(v);
        __syncwarp();
        static_for<32>([&](auto QQ) {
            constexpr int q = decltype(QQ)::value;
            y[lane * 33 + q] = v[bitrev(q, 5)];
        });
        __syncwarp();
        static_for<32>([&](auto RR) {
            constexpr int r = decltype(RR)::value;
            v[r] = y[lane + 33 * r];
        });
        // the twiddles between the passes ride in the first butterfly stage (fft.cuh, fft_reg_pre): 16 packed instructions less
        fft_reg_pre<32, FFT_FWD, 0, 32, true>(v, [&](auto RR) {
            constexpr int r = decltype(RR)::value;
            const float2 w = S.tw2[(r > 0 ? r : 1) - 1][lane];
            return make_float2(w.x, FFT_FWD < 0 ? -w.y : w.y);
        });
        __syncwarp();
        static_for<32>([&](auto QQ) {
            constexpr int q = decltype(QQ)::value;
            y[lane + 33 * q] = v[bitrev(q, 5)];  // natural order: element lane + 32 q
        });
        __syncwarp();

        if constexpr (OS) {
            // stage C, overlap-save: y[n] = swap(z'[16 n]) of the WINDOW; window position p >= os_hist is stream
            // sample g = z0 + b * hop + (p - os_hist).  Keep z[g], (g mod DB) = D*i, i < M, g < os_zend.  A
            // window's valid range can straddle a DecimateReader block (hop does not divide 32768): walk the
            // (at most two) blocks it touches.
            const uint32_t hist = (uint32_t)kC16N - prm.os_hop;
            const uint32_t g0 = prm.z0 + b * prm.os_hop;
            uint32_t g_end = g0 + prm.os_hop;
            if (g_end > prm.os_zend) g_end = prm.os_zend;
            for (uint32_t g_lo = g0; g_lo < g_end;) {
                const uint32_t blk = g_lo >> prm.db_log2;
                uint32_t blk_end = (blk + 1u) << prm.db_log2;
                if (blk_end > g_end) blk_end = g_end;
                const uint32_t p0 = g_lo & db_mask;
                const uint32_t o0 = (p0 + prm.D - 1u) / prm.D;
                const uint32_t gk = (blk << prm.db_log2) + o0 * prm.D;  // first kept sample at or after g_lo
                if (o0 < prm.M && gk < blk_end) {
                    uint32_t cnt = (blk_end - 1u - gk) / prm.D + 1u;
                    if (cnt > prm.M - o0) cnt = prm.M - o0;
                    float2 *out = prm.dst + (size_t)blk * prm.M + o0;
                    for (uint32_t k = lane; k < cnt; k += 32u) {
                        const uint32_t n = (hist + (gk - g0) + k * prm.D) >> 4;
                        const float2 z = y[n + (n >> 5)];
                        out[k] = make_float2(z.y, z.x);
                    }
                }
                g_lo = blk_end;
            }
        } else {
        // stage C: y[n] = swap(z[16 n]).  Keep z[g], (g mod DB) = D*i, i < M  (D is a multiple of 16)
        const uint32_t s0 = b * (uint32_t)kC16N;
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
        for (uint32_t k = lane; k < cnt; k += 32u) {
            const uint32_t n = (pos0 + k * prm.D) >> 4;
            const float2 z = y[n + (n >> 5)];
            out[k] = make_float2(z.y, z.x);
        }
        }
        __syncwarp();
    }
}

template <int FMT, bool LSB, bool OS>
__global__ void __launch_bounds__(kC16Threads, 1) k_chain16k(const __grid_constant__ ChainParams prm,
                                                              const __grid_constant__ NcoTable nco) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Chain16kSmem &S = *reinterpret_cast<Chain16kSmem *>(smem_raw);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;

    overlap_trigger();  // see chain1024.cu: consecutive buffers are independent

    const uint32_t n_it = prm.nblocks > blockIdx.x ? (prm.nblocks - blockIdx.x + gridDim.x - 1u) / gridDim.x : 0u;
    if (warp == 16) {
        c16_inverse_warp<OS>(S, prm, n_it);
        return;
    }

    uint32_t b = blockIdx.x;
    if constexpr (OS) {
        // the call's first window reads the history the PREVIOUS launch wrote (its tail CTA, below): this CTA -- and
        // only this one -- waits for that launch to have completed; every other window depends on nothing earlier
        if (prm.os_head && blockIdx.x == 0) asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    if (n_it) c16_prefetch<FMT, OS>(S, prm, b);
    for (int i = t; i < 31 * 32; i += kC16Workers) (&S.tw2[0][0])[i] = __ldg(prm.tw + i);

    const float sc = RawTraits<FMT>::scale();
    const float2 *hp = prm.tw1k + warp * 1024 + lane;  // permuted filter: hp[32 q] = H[warp + 16 (lane + 32 q)]
    uint64_t rot_dp = 0;
    bool rot_ok = false;

    for (uint32_t it = 0; it < n_it; ++it, b += gridDim.x) {
        const uint32_t s0 = OS ? b * prm.os_hop : b * (uint32_t)kC16N;  // launch coordinates of the block's / window's first sample
        // overlap-save at stream start: the history in front of the first window is silence, whatever the format's zero code is
        const bool zero_head = OS && b == 0u && prm.os_zero_head != 0u;

        // ------------------------------------------------------------------ NCO tables of this block
        const int si = nco_find(nco, s0);
        const uint32_t seg_j0 = nco.seg[si].j0, seg_end = seg_j0 + nco.seg[si].count;
        const uint64_t seg_p0 = nco.seg[si].p0, seg_dp = nco.seg[si].dp;
        const bool fast = s0 + (uint32_t)kC16N <= seg_end;
        if (fast && (!rot_ok || rot_dp != seg_dp)) {  // uniform over the CTA; the last readers passed barrier 2
            if (t < 16) S.rot[t] = nco_rot((uint64_t)(1024u * t) * seg_dp);
            rot_dp = seg_dp;
            rot_ok = true;
        }
        cp_async_wait_all();
        bar_sync(kBarWorkers, kC16Workers);  // raw block landed; rot (and, first time, tw2) visible; x free

        // ------------------------------------------------------------------ pass 1: radix 16 over c, n = m + 1024 c
        // two butterflies per thread (m = t, t + 512); phase(s0 + m + 1024 c) = ph_m + 1024 c dP
#pragma unroll 1
        for (int i = 0; i < 2; ++i) {
            const int m = t + 512 * i;
            const int mp = xpad(m);
            float2 w[16];
            if (fast) {
                float2 r0 = nco_rot(seg_p0 + (uint64_t)(s0 + m - seg_j0 + 1) * seg_dp);
                r0 = mul2(r0, make_float2(sc, sc));
                static_for<16>([&](auto CC) {
                    constexpr int c = decltype(CC)::value;
                    const float2 rot = c == 0 ? r0 : cmul(r0, S.rot[c]);
                    w[c] = cmul(RawTraits<FMT>::unscaled2(c16_raw<FMT, LSB>(S, m + 1024 * c, prm.lsb_shift)), rot);
                });
                if constexpr (OS) {
                    if (zero_head) {  // os_hist = 4096 = rows c < 4 of every column m
                        const uint32_t zc = ((uint32_t)kC16N - prm.os_hop) >> 10;
                        static_for<16>([&](auto CC) {
                            constexpr int c = decltype(CC)::value;
                            if ((uint32_t)c < zc) w[c] = make_float2(0.f, 0.f);
                        });
                    }
                }
            } else {  // block straddles accumulator segments: per-sample phase, through the thread's own x slots
                NcoCursor cur;
#pragma unroll 1
                for (int c = 0; c < 16; ++c) {
                    const uint32_t j = s0 + m + 1024u * c;
                    cur.seek(nco, j);
                    float2 rot = nco_rot(cur.phase(j));
                    rot = mul2(rot, make_float2(sc, sc));
                    float2 xv = cmul(RawTraits<FMT>::unscaled2(c16_raw<FMT, LSB>(S, m + 1024 * c, prm.lsb_shift)), rot);
                    if (OS && (zero_head && (uint32_t)(m + 1024 * c) < (uint32_t)kC16N - prm.os_hop)) xv = make_float2(0.f, 0.f);
                    S.x[c][mp] = xv;
                }
                static_for<16>([&](auto CC) {
                    constexpr int c = decltype(CC)::value;
                    w[c] = S.x[c][mp];
                });
            }
            fft_reg<16, FFT_FWD, 0, 16>(w);
            static_for<15>([&](auto KK) {  // W_16384^{m k}
                constexpr int k = decltype(KK)::value + 1;
                const float2 tw = __ldg(prm.tw3 + (k - 1) * 1024 + m);
                w[bitrev(k, 4)] = tw_mul<FFT_FWD>(w[bitrev(k, 4)], tw.x, tw.y);
            });
            static_for<16>([&](auto KK) {
                constexpr int k = decltype(KK)::value;
                S.x[k][mp] = w[bitrev(k, 4)];
            });
        }
        bar_sync(kBarScatter, kC16Threads);  // x complete, raw consumed; y[it & 1] is free again
        if (it + 1u < n_it) c16_prefetch<FMT, OS>(S, prm, b + gridDim.x);

        // ------------------------------------------------------------------ sub-transform `warp`: 1024 = 32 x 32, warp-local
        float2 *buf = S.x[warp];
        float2 v[32];
        static_for<32>([&](auto RR) {
            constexpr int r = decltype(RR)::value;
            v[r] = buf[lane + 33 * r];  // element lane + 32 r
        });
        fft_reg<32, FFT_FWD, 0, 32>(v);
        __syncwarp();
        static_for<32>([&](auto QQ) {
            constexpr int q = decltype(QQ)::value;
            buf[lane * 33 + q] = v[bitrev(q, 5)];  // Ns = 1: index 32 lane + q
        });
        __syncwarp();
        static_for<32>([&](auto RR) {
            constexpr int r = decltype(RR)::value;
            v[r] = buf[lane + 33 * r];
        });
        fft_reg_pre<32, FFT_FWD, 0, 32, true>(v, [&](auto RR) {
            constexpr int r = decltype(RR)::value;
            const float2 w = S.tw2[(r > 0 ? r : 1) - 1][lane];
            return make_float2(w.x, FFT_FWD < 0 ? -w.y : w.y);
        });

        // ------------------------------------------------------------------ x H, fold 16:1 (fft/convolution.go:187-189)
        // v[bitrev(q)] = X[warp + 16 (lane + 32 q)]; bins q = p + 2 c alias onto folded bin warp + 16 (lane + 32 p)
        float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
        static_for<16>([&](auto CC) {
            constexpr int c = decltype(CC)::value;
            acc0 = cmac_swapped(acc0, v[bitrev(2 * c, 5)], __ldg(hp + 32 * (2 * c)));
            acc1 = cmac_swapped(acc1, v[bitrev(2 * c + 1, 5)], __ldg(hp + 32 * (2 * c + 1)));
        });
        float2 *y = S.y[it & 1];
        const int j0 = warp + 16 * lane;
        y[xpad(j0)] = acc0;
        y[xpad(j0 + 512)] = acc1;
        __threadfence_block();
        bar_arrive(kBarFull + (int)(it & 1u), kC16Threads);
    }
    cp_async_wait_all();
    if constexpr (OS) {
        // the CTA that ran the launch's last window saves the buffer's last os_hist raw samples for the next call
        // (ping-pong with the history this launch read; the wait covers the previous launch's read of the other half)
        if (prm.hist_out && blockIdx.x == (prm.nblocks - 1u) % gridDim.x) {
            asm volatile("griddepcontrol.wait;" ::: "memory");
            const uint4 *ts = reinterpret_cast<const uint4 *>(prm.tail_src);
            uint4 *td = reinterpret_cast<uint4 *>(prm.hist_out);
            for (uint32_t i = t; i < prm.tail_bytes / 16u; i += kC16Workers) td[i] = __ldg(ts + i);
        }
    }
    if (t == 0) overlap_join(prm.done);  // see the end of k_chain1024
}

template <int FMT, bool LSB, bool OS>
static int launch16(hzsdr_ctx *ctx, const ChainParams &prm_in, const NcoTable &nco) {
    ChainParams prm = prm_in;
    static PerDevice attr_set;
    const size_t smem = sizeof(Chain16kSmem);
    int rc = attr_set.once(ctx->device, [&](int &) {
        HZ_CUDA(cudaFuncSetAttribute((const void *)k_chain16k<FMT, LSB, OS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        return (int)HZSDR_OK;
    });
    if (rc) return rc;
    const int grid = (int)(prm.nblocks < (uint32_t)ctx->sm_count ? prm.nblocks : (uint32_t)ctx->sm_count);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kC16Threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    constexpr int sb = FMT == HZSDR_FORMAT_I16 ? 4 : 2;
    bool may;
    if constexpr (OS) {
        // spans, conservatively: the buffer part this launch reads and the whole output range of the call.  The carried
        // history is the library's own memory: the CTAs that read / write it wait for the previous launch themselves.
        const size_t out_bytes = (size_t)((prm.os_zend >> prm.db_log2) * prm.M) * sizeof(float2);
        may = ctx->overlap.admit(OverlapWindow::span(prm.src + (size_t)prm.os_head * sb, (size_t)(prm.os_valid - prm.os_head) * sb),
                                 OverlapWindow::span(prm.dst, out_bytes), ctx->overlap_pred_ok());
        prm.done = ctx->overlap_done + ctx->overlap.slot();
        ctx->overlap_launched();
    } else {
        may = chain_may_overlap(ctx, prm, (uint32_t)kC16N, sb);
    }
    overlap_launch_config(cfg, attr, may);
    HZ_CUDA(cudaLaunchKernelEx(&cfg, k_chain16k<FMT, LSB, OS>, prm, nco));
    return HZSDR_OK;
}

// prm.tw = [31][32] W_1024^{r l}; prm.tw3 = [15][1024] W_16384^{k m}; prm.tw1k = the permuted filter
int launch_chain16k(hzsdr_ctx *ctx, int fmt, const ChainParams &prm, const NcoTable &nco) {
    if (prm.os_hop) {  // overlap-save windows (prm.nblocks of them)
        switch (fmt) {
            case HZSDR_FORMAT_U8: return launch16<HZSDR_FORMAT_U8, false, true>(ctx, prm, nco);
            case HZSDR_FORMAT_I8: return launch16<HZSDR_FORMAT_I8, false, true>(ctx, prm, nco);
            default:
                return prm.lsb_shift ? launch16<HZSDR_FORMAT_I16, true, true>(ctx, prm, nco)
                                     : launch16<HZSDR_FORMAT_I16, false, true>(ctx, prm, nco);
        }
    }
    switch (fmt) {
        case HZSDR_FORMAT_U8: return launch16<HZSDR_FORMAT_U8, false, false>(ctx, prm, nco);
        case HZSDR_FORMAT_I8: return launch16<HZSDR_FORMAT_I8, false, false>(ctx, prm, nco);
        default:
            return prm.lsb_shift ? launch16<HZSDR_FORMAT_I16, true, false>(ctx, prm, nco)
                                 : launch16<HZSDR_FORMAT_I16, false, false>(ctx, prm, nco);
    }
}

void chain16k_twiddles(float2 *tw2 /* 31*32 */, float2 *tw3 /* 15*1024 */) {
    auto w = [](double num, double den) {
        const double a = 2.0 * M_PI * num / den;
        return make_float2((float)cos(a), (float)sin(a));
    };
    for (int r = 1; r < 32; r++)
        for (int l = 0; l < 32; l++) tw2[(r - 1) * 32 + l] = w(r * l, 1024.0);
    for (int k = 1; k < 16; k++)
        for (int m = 0; m < 1024; m++) tw3[(k - 1) * 1024 + m] = w((double)k * m, 16384.0);
}

// Hp[(k*32 + q)*32 + l] = H[k + 16 (l + 32 q)]: the order warp k, lane l reads its bins in
void chain16k_permute_filter(const float2 *H /* 16384 */, float2 *Hp /* 16384 */) {
    for (int k = 0; k < 16; k++)
        for (int q = 0; q < 32; q++)
            for (int l = 0; l < 32; l++) Hp[(k * 32 + q) * 32 + l] = H[k + 16 * (l + 32 * q)];
}

}  // namespace hz
