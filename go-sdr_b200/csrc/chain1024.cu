// chain1024.cu -- the fused chain specialised for ConvolutionReader blocks of N = 1024
// (BASELINE config 2: 255-tap lowpass as a 1024-bin frequency-domain filter).
//
// One WARP owns one 1024-sample block end to end; nothing is shared between warps except two
// read-only tables, so the only barrier in the steady state is __syncwarp.
//
//   stage A  32 raw samples per lane (stride-32 pattern of the first radix-32 pass) -> float ->
//            NCO mix.  One accurate sincos per lane per block; the other 31 rotations come from
//            two tiny per-block tables (e^{i*32*b*dP}, b < 8 and e^{i*256*a*dP}, a < 4) by complex
//            multiplication, depth <= 2, because inside one accumulator segment the phase is
//            linear in the sample index (nco.cuh).
//   loop x4  radix-32 butterflies in registers -> padded shared-memory exchange -> gather ->
//            {twiddle | xH}.  The inverse transform runs on re/im-swapped data, so all four
//            passes execute the SAME forward code: IDFT(Y) = swap(DFT(swap(Y))).  That keeps the
//            hot loop at ~14 KB of SASS -- inside the instruction cache, which the fully unrolled
//            generic kernel (80 KB) was not (26% of its stall samples were instruction fetch).
//   stage C  DecimateReader: the lanes copy only the kept samples z[q*32768 + D*i] out of shared
//            memory, coalesced.
//
// Algorithmic HBM bytes: raw bytes in + 8 B per kept sample out (2.80 B/sample for i8, D = 10).
// The kernel is FP32-issue bound (about 90 lane-instructions per sample), not HBM bound.
#include "common.cuh"
#include "fft.cuh"
#include "fft_kernels.cuh"
#include "nco.cuh"

namespace hz {

#ifndef HZ_C1024_WARPS
#define HZ_C1024_WARPS 4
#endif
#ifndef HZ_C1024_CTAS
#define HZ_C1024_CTAS 4
#endif
constexpr int kC1024Warps = HZ_C1024_WARPS;      // warps per CTA
constexpr int kC1024MinCtas = HZ_C1024_CTAS;     // resident CTAs per SM the register budget is cut for
constexpr int kC1024Threads = 32 * kC1024Warps;
constexpr int kC1024TwFull = 32 * 32;                 // table layout in global memory (complex entries):
constexpr int kC1024TwB = 15 * 32, kC1024TwC = 8 * 32;  // [tw | twB | twC]

struct Chain1024Smem {
    float2 tw[32][32];             // tw[r][lane] = W_1024^{r*lane}  (cos, sin), forward sign applied in tw_mul; SPLIT
                                   // kernels get it multiplied by conj(e^{i lane dP}) * scale (row 0 included), see below
    float2 H[32][32];              // H[r][lane]    = filter[32*r + lane]
    float2 twB[15][32];            // pruned inverse, 2nd radix-16 pass: W_256^{r*(lane&15)}, r = 1..15
    float2 twC[8][32];             // pruned inverse, radix-2 pass:      W_512^{lane + 32 i}, i < 8
    float2 rot[kC1024Warps][32];   // per-warp NCO step tables of the current segment
    float2 buf[kC1024Warps][1056]; // per-warp exchange buffer, index padded a + a/32
};

// LSB: the Pluto LSB->MSB shift (iq_i16.go:103-111) is compiled in only for chains that ask for it --
// as a run-time test it costs four predicated-off instructions per sample in every i16 chain
template <int FMT, bool LSB>
__device__ __forceinline__ uint32_t c1024_load_raw(const uint8_t *__restrict__ src, uint32_t j, int lsb_shift) {
    if constexpr (FMT == HZSDR_FORMAT_I16) {
        uint32_t w = __ldg(reinterpret_cast<const uint32_t *>(src) + j);
        if constexpr (LSB) w = ((w & 0xffff0000u) << lsb_shift) | (((w & 0xffffu) << lsb_shift) & 0xffffu);
        return w;
    } else {
        return (uint32_t)__ldg(reinterpret_cast<const uint16_t *>(src) + j);
    }
}

// raw word -> exact unscaled complex float; the format's scale is folded into the NCO rotation
// (exact for i8's 2^-7; <= 1 ulp from the reference's division for u8 / i16, inside the 1e-5 bar)
template <int FMT>
__device__ __forceinline__ float2 c1024_to_float(uint32_t w) {
    if constexpr (FMT == HZSDR_FORMAT_C64)
        return make_float2(0.f, 0.f);  // never used: complex64 input skips stage A's conversion
    else
        return RawTraits<FMT>::unscaled2(w);
}
template <int FMT>
__device__ __forceinline__ float c1024_fold_scale() {
    if constexpr (FMT == HZSDR_FORMAT_C64)
        return 1.0f;
    else
        return RawTraits<FMT>::scale();
}

__device__ __forceinline__ uint32_t udiv_small(uint32_t x, uint32_t d, float inv_d) {
    // floor(x / d) for x < 2^24: float estimate (never too large: inv_d is rounded down) + 1 fix-up
    uint32_t q = __float2uint_rz(__uint2float_rz(x) * inv_d);
    if (x - q * d >= d) q++;
    return q;
}

// BATCH = false: one stream, its segment table in the kernel parameters (`nco`).
// BATCH = true : prm.nstreams streams x prm.nblocks blocks (channelizer); source, destination and
//                segment table of each stream come from prm.streams[] in device memory.
// SPLIT: the mixer rotation of sample lane + 32 r of a block is factored
//            e^{i phi_b} * e^{i lane dP} * e^{i 32 r dP}
//        and only the last factor (32 values per segment, shared by the warp) is multiplied onto
//        the samples.  e^{i lane dP} is constant along the first pass's transform direction, so it
//        (and the format's scale) rides in the twiddle table that follows that pass -- the host
//        builds prm.tw with it (chain1024_split_twiddles, dP = prm.dp_nom) -- and the block's phase
//        e^{i phi_b} commutes with everything linear, so it is applied to the ~1024/D samples that
//        survive the decimation instead of to all 1024.  54 fewer packed fp32 instructions per lane
//        and block (-6%).  dP differs between accumulator segments by <= 9e-9 relative (the fp64 step
//        is rounded to the accumulator's binade); the table uses the launch's dominant segment, so
//        the e^{i lane dP} factor can be off by <= 31 * dP * 9e-9 turns < 2.2e-7 rad in the others.
//        Even decimation factors only.  BATCH + SPLIT: every stream has its own table (StreamDesc::tw), so
//        a CTA works through a CONTIGUOUS range of prm.per_cta blocks -- nearly always one stream -- and
//        reloads the table behind a CTA barrier when the range crosses into the next stream.
template <int FMT, bool BATCH, bool LSB, bool SPLIT>
__global__ void __launch_bounds__(kC1024Threads, kC1024MinCtas) k_chain1024(const __grid_constant__ ChainParams prm,
                                                                 const __grid_constant__ NcoTable nco) {
    static_assert(!SPLIT || FMT != HZSDR_FORMAT_C64, "SPLIT: raw input");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Chain1024Smem &S = *reinterpret_cast<Chain1024Smem *>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // Let the next chain launch in the stream start filling SMs as this one drains (programmatic
    // dependent launch): consecutive buffers are independent -- all carried state (the NCO time)
    // lives on the host -- so this kernel never waits on its predecessor either.
    if constexpr (!BATCH) overlap_trigger();  // (a batched launch's spans are not tracked: its successor waits for it)

    // raw samples of a block -> L2 ahead of time (one 128-byte line per lane, no registers held): issued
    // for the warp's first block before the table prologue, and for its next block while the current
    // one is being transformed, so stage A's loads find them in L2 instead of waiting on HBM
    constexpr uint32_t kRawBlockBytes = 1024u * (FMT == HZSDR_FORMAT_C64 ? 8u : FMT == HZSDR_FORMAT_I16 ? 4u : 2u);
    auto prefetch_raw = [&](const uint8_t *block_src) {
#pragma unroll
        for (uint32_t o = 0; o < kRawBlockBytes; o += 32u * 128u)
            if (o + 128u * lane < kRawBlockBytes) asm volatile("prefetch.global.L2 [%0];" ::"l"(block_src + o + 128u * lane));
    };
    if constexpr (!BATCH) {
        const uint32_t gb0 = blockIdx.x * kC1024Warps + warp;
        if (gb0 < prm.nblocks) prefetch_raw(prm.src + (size_t)gb0 * kRawBlockBytes);
    }

    // tables -> shared memory, every load in flight before the first store.  Global layout
    // [tw | twB | twC] and H; shared layout tw, H, twB, twC.
    {
        constexpr int kTwVec = kC1024TwFull / 2, kHVec = 1024 / 2, kBcVec = (kC1024TwB + kC1024TwC) / 2;  // float4 = 2 complex
        constexpr int kAll = kTwVec + kHVec + kBcVec;
        constexpr int kPer = (kAll + kC1024Threads - 1) / kC1024Threads;
        const float4 *gtw = reinterpret_cast<const float4 *>(prm.tw);
        // (a SPLIT launch's 32 x 32 table is one of several cached per chain; twB / twC are shared)
        const float4 *gbc = reinterpret_cast<const float4 *>(prm.tw_bc ? prm.tw_bc : prm.tw + kC1024TwFull);
        const float4 *gh = reinterpret_cast<const float4 *>(prm.H);
        float4 *stw = reinterpret_cast<float4 *>(&S.tw[0][0]);
        float4 *sh = reinterpret_cast<float4 *>(&S.H[0][0]);
        float4 *sbc = reinterpret_cast<float4 *>(&S.twB[0][0]);  // twB and twC are contiguous
        float4 t[kPer];
#pragma unroll
        for (int u = 0; u < kPer; u++) {
            const int i = threadIdx.x + u * kC1024Threads;
            if (i < kTwVec)
                t[u] = __ldg(gtw + i);
            else if (i < kTwVec + kHVec)
                t[u] = __ldg(gh + (i - kTwVec));
            else if (i < kAll)
                t[u] = __ldg(gbc + (i - kTwVec - kHVec));
        }
#pragma unroll
        for (int u = 0; u < kPer; u++) {
            const int i = threadIdx.x + u * kC1024Threads;
            if (i < kTwVec)
                stw[i] = t[u];
            else if (i < kTwVec + kHVec)
                sh[i - kTwVec] = t[u];
            else if (i < kAll)
                sbc[i - kTwVec - kHVec] = t[u];
        }
    }
    __syncthreads();

    float2 *buf = S.buf[warp];
    float2 *rt = S.rot[warp];
    const uint32_t db_mask = (1u << prm.db_log2) - 1u;
    const uint32_t nwarps = gridDim.x * kC1024Warps;

    // carried from block to block of this warp: the accumulator segment found last (a steady-state
    // buffer has 1-3 of them, so the next block is nearly always in the same one) and the step for
    // which the rotation tables in `rt` were built
    uint32_t seg_j0 = 1u, seg_end = 0u, seg_stream = 0xffffffffu;
    uint64_t seg_p0 = 0, seg_dp = 0, rt_dp = 0;
    bool rt_ok = false;

    const uint32_t total_blocks = BATCH ? prm.nblocks * prm.nstreams : prm.nblocks;
    // block walk: chip-wide stride, or (BATCH + SPLIT) the CTA's own contiguous range, a group of
    // kC1024Warps consecutive blocks per iteration (per_cta and nblocks are multiples of the group, so a
    // group never straddles two streams and the table switch below is uniform over the CTA)
    constexpr bool kRanged = BATCH && SPLIT;
    const uint32_t gb_first = kRanged ? blockIdx.x * prm.per_cta + warp : blockIdx.x * kC1024Warps + warp;
    const uint32_t gb_step = kRanged ? (uint32_t)kC1024Warps : nwarps;
    const uint32_t gb_end = kRanged ? min(total_blocks + (uint32_t)kC1024Warps - 1u, (blockIdx.x + 1u) * prm.per_cta + warp) : total_blocks;
    uint32_t table_stream = 0xffffffffu;
    for (uint32_t gb = gb_first; gb < gb_end; gb += gb_step) {
        uint32_t b = gb;
        const uint8_t *src = prm.src;
        float2 *dst = prm.dst;
        const StreamDesc *sd = nullptr;
        uint32_t st = 0;
        uint64_t dp_nom = prm.dp_nom;
        if constexpr (kRanged) {
            const uint32_t group = gb - warp;  // first block of this iteration's group: the same in every warp
            if (group >= total_blocks) break;
            const uint32_t gst = group / prm.nblocks;
            if (gst != table_stream) {
                __syncthreads();  // everybody is done with the previous stream's table
                const float4 *g = reinterpret_cast<const float4 *>(prm.streams[gst].tw);
                float4 *d = reinterpret_cast<float4 *>(&S.tw[0][0]);
                float4 t[kC1024TwFull / 2 / kC1024Threads];
#pragma unroll
                for (int u = 0; u < kC1024TwFull / 2 / kC1024Threads; u++) t[u] = __ldg(g + threadIdx.x + u * kC1024Threads);
#pragma unroll
                for (int u = 0; u < kC1024TwFull / 2 / kC1024Threads; u++) d[threadIdx.x + u * kC1024Threads] = t[u];
                table_stream = gst;
                __syncthreads();
            }
            if (gb >= total_blocks) continue;  // (cannot happen while total_blocks is a multiple of the group)
        }
        if constexpr (BATCH) {
            st = gb / prm.nblocks;
            b = gb - st * prm.nblocks;
            sd = prm.streams + st;
            src = sd->src;
            dst = reinterpret_cast<float2 *>(sd->dst);
            if constexpr (SPLIT) dp_nom = sd->dp_nom;
        }
        const uint32_t s0 = b * 1024u;  // index of the block's first sample within its stream's buffer
        float2 v[32];
        float2 blk = make_float2(1.0f, 0.0f);  // SPLIT: e^{i phi_b}, applied to the kept outputs in stage C

        // ------------------------------------------------------------------ stage A
        if constexpr (FMT == HZSDR_FORMAT_C64) {
            // ConvolutionReader on its own (hzsdr_convolve_freq): complex64 in, no conversion, no mixer
            const float2 *x = reinterpret_cast<const float2 *>(src) + s0 + lane;
            static_for<32>([&](auto RR) {
                constexpr int r = decltype(RR)::value;
                v[r] = ld_stream_f2(x + 32 * r);
            });
        } else {
        if (!(s0 >= seg_j0 && s0 < seg_end && (!BATCH || st == seg_stream))) {
            if constexpr (BATCH) {
                const int si = nco_find(*sd, s0);
                seg_j0 = sd->seg[si].j0, seg_end = seg_j0 + sd->seg[si].count;
                seg_p0 = sd->seg[si].p0, seg_dp = sd->seg[si].dp;
            } else {
                const int si = nco_find(nco, s0);
                seg_j0 = nco.seg[si].j0, seg_end = seg_j0 + nco.seg[si].count;
                seg_p0 = nco.seg[si].p0, seg_dp = nco.seg[si].dp;
            }
            seg_stream = st;
        }
        if (SPLIT && s0 + 1024u <= seg_end) {
            // whole block inside one linear segment: phase(s0 + lane + 32 r) = phi_b + lane dP + 32 r dP
            if (!rt_ok || rt_dp != seg_dp) {
                __syncwarp();
                rt[lane] = nco_rot((uint64_t)(32u * lane) * seg_dp);  // e^{i 32 r dP}, r = lane
                __syncwarp();
                rt_dp = seg_dp;
                rt_ok = true;
            }
            blk = nco_rot(seg_p0 + (uint64_t)(s0 - seg_j0 + 1) * seg_dp);
            static_for<4>([&](auto AA) {
                constexpr int a = decltype(AA)::value;
                uint32_t raw[8];
                static_for<8>([&](auto BB) {
                    constexpr int bb = decltype(BB)::value;
                    raw[bb] = c1024_load_raw<FMT, LSB>(src, s0 + lane + 32u * (8 * a + bb), prm.lsb_shift);
                });
                static_for<8>([&](auto BB) {
                    constexpr int r = 8 * a + decltype(BB)::value;
                    v[r] = c1024_to_float<FMT>(raw[r & 7]);  // x e^{i 32 r dP} happens inside the first pass
                });
            });
        } else if (!SPLIT && s0 + 1024u <= seg_end) {
            // whole block inside one linear segment: phase(s0 + lane + 32 r) = ph + 32 r dP
            if (!rt_ok || rt_dp != seg_dp) {
                // rt[b] = e^{i 32 b dP}, b < 8; rt[8 + a] = e^{i 256 a dP}, a < 4: one evaluation, 12 lanes keep it
                const float2 t = nco_rot((uint64_t)(lane < 8 ? 32u * lane : 256u * (lane - 8u)) * seg_dp);
                if (lane < 12) rt[lane] = t;
                __syncwarp();
                rt_dp = seg_dp;
                rt_ok = true;
            }
            float2 r0 = nco_rot(seg_p0 + (uint64_t)(s0 + lane - seg_j0 + 1) * seg_dp);
            const float sc = c1024_fold_scale<FMT>();
            r0 = mul2(r0, make_float2(sc, sc));
            static_for<4>([&](auto AA) {
                constexpr int a = decltype(AA)::value;
                uint32_t raw[8];
                static_for<8>([&](auto BB) {
                    constexpr int bb = decltype(BB)::value;
                    raw[bb] = c1024_load_raw<FMT, LSB>(src, s0 + lane + 32u * (8 * a + bb), prm.lsb_shift);
                });
                const float2 ra = a == 0 ? r0 : cmul(r0, rt[8 + a]);
                static_for<8>([&](auto BB) {
                    constexpr int bb = decltype(BB)::value;
                    const float2 rot = bb == 0 ? ra : cmul(ra, rt[bb]);
                    v[8 * a + bb] = cmul(c1024_to_float<FMT>(raw[bb]), rot);
                });
            });
        } else {
            // The block straddles accumulator segments (stream start, binade edge, 2*pi wrap): mix
            // sample by sample.  Kept as a compact loop through shared memory so that this rarely
            // taken path does not bloat the hot instruction footprint.
            // (SPLIT: the table that follows the first pass carries e^{i lane dP_nom} and the scale)
            // and the first pass multiplies sample r by rt[r] = e^{i 32 r rt_dp}: both are taken out here)
            const float sc = SPLIT ? 1.0f : c1024_fold_scale<FMT>();
            if (SPLIT && !rt_ok) {
                rt[lane] = nco_rot((uint64_t)(32u * lane) * dp_nom);
                rt_dp = dp_nom;
                rt_ok = true;
            }
            const uint64_t back = SPLIT ? (uint64_t)lane * dp_nom : 0ull, back_r = SPLIT ? 32u * rt_dp : 0ull;
            blk = make_float2(1.0f, 0.0f);
            NcoCursor cur;
            __syncwarp();
#pragma unroll 1
            for (int r = 0; r < 32; ++r) {
                const uint32_t j = s0 + lane + 32u * r;
                if constexpr (BATCH)
                    cur.seek(*sd, j);
                else
                    cur.seek(nco, j);
                float2 rot = nco_rot(cur.phase(j) - back - (uint64_t)r * back_r);
                rot.x *= sc;
                rot.y *= sc;
                buf[lane + 33 * r] = cmul(c1024_to_float<FMT>(c1024_load_raw<FMT, LSB>(src, j, prm.lsb_shift)), rot);
            }
            __syncwarp();
            static_for<32>([&](auto RR) {
                constexpr int r = decltype(RR)::value;
                v[r] = buf[lane + 33 * r];
            });
        }
        }  // FMT != C64
        {   // the warp's next block, when it is in the same stream (its source pointer is at hand)
            const uint32_t nb = b + gb_step;
            if (nb < prm.nblocks && gb + gb_step < gb_end) prefetch_raw(src + (size_t)nb * kRawBlockBytes);
        }

        // ------------------------------------------------------------------ FFT, xH, IFFT
        // D even: only even-indexed z are ever kept (decimate blocks start on multiples of 1024), and
        //   z[2m] = IDFT_512(Y[k] + Y[k+512])[m]
        // so the inverse shrinks to a folded 512-point transform (16 x 16 x 2).
        const bool prune2 = SPLIT || (prm.D & 1u) == 0u;  // (SPLIT launches are pruned ones: the other path compiles away)
        const int npass = SPLIT ? 0 : (prune2 ? 2 : 4);
        if constexpr (SPLIT) {
            // two passes, each with its multipliers folded into the first butterfly stage (fft_reg_pre):
            // e^{i 32 r dP} from the warp's table, then W_1024^{r lane} e^{i r dP} scale from the launch's
            fft_reg_pre<32, FFT_FWD, 0, 32, true>(v, [&](auto RR) { return rt[decltype(RR)::value]; });
            __syncwarp();
            static_for<32>([&](auto QQ) {
                constexpr int q = decltype(QQ)::value;
                buf[lane * 33 + q] = v[bitrev(q, 5)];  // Ns = 1: index 32 lane + q, padded
            });
            __syncwarp();
            static_for<32>([&](auto RR) {
                constexpr int r = decltype(RR)::value;
                v[r] = buf[lane + 33 * r];  // element lane + 32 r
            });
            fft_reg_pre<32, FFT_FWD, 0, 32, false>(v, [&](auto RR) { return S.tw[decltype(RR)::value][lane]; });
        }
#pragma unroll 1
        for (int pass = 0; pass < npass; ++pass) {
            fft_reg<32, FFT_FWD, 0, 32>(v);
            // pruned path: after the second pass lane `lane` already holds X[lane + 32 q] (in
            // v[bitrev(q)]) -- exactly what the filter multiply and the fold need: no exchange
            if (prune2 && pass == 1) break;
            __syncwarp();  // every lane is done reading buf (previous gather / previous block's stage C)
            if (pass & 1) {
                static_for<32>([&](auto QQ) {
                    constexpr int q = decltype(QQ)::value;
                    buf[lane + 33 * q] = v[bitrev(q, 5)];  // Ns = 32: index lane + 32 q, padded
                });
            } else {
                static_for<32>([&](auto QQ) {
                    constexpr int q = decltype(QQ)::value;
                    buf[lane * 33 + q] = v[bitrev(q, 5)];  // Ns = 1: index 32 lane + q, padded
                });
            }
            __syncwarp();
            if (pass == 3) break;
            static_for<32>([&](auto RR) {
                constexpr int r = decltype(RR)::value;
                v[r] = buf[lane + 33 * r];  // element lane + 32 r
            });
            if (pass == 1) {
                // spectrum x filter (fft/convolution.go:187-189), then swap re/im: forward passes on
                // swapped data compute the unnormalised inverse transform (swapped)
                static_for<32>([&](auto RR) {
                    constexpr int r = decltype(RR)::value;
                    v[r] = cmul_swapped(v[r], S.H[r][lane]);
                });
            } else {
                static_for<31>([&](auto RR) {
                    constexpr int r = decltype(RR)::value + 1;
                    const float2 w = S.tw[r][lane];
                    v[r] = tw_mul<FFT_FWD>(v[r], w.x, w.y);
                });
            }
        }

        if (prune2) {
            // X[lane + 32 r] = v[bitrev(r)].  Filter, fold (Y[k] + Y[k+512]), swap re/im; then
            // 512 = 16 x 16 x 2 forward passes on the swapped data (Stockham, padding a + a/16).
            float2 w[16];
            static_for<16>([&](auto RR) {
                constexpr int r = decltype(RR)::value;
                // element lane + 32 r of the 512-point spectrum; the second product rides in the add (4 instructions, not 5)
                w[r] = cmac_swapped(cmul_swapped(v[bitrev(r, 5)], S.H[r][lane]), v[bitrev(r + 16, 5)], S.H[r + 16][lane]);
            });
            // passes A and B share their butterfly code (one 2-iteration loop):
            //   A: radix 16, Ns = 1,  item j = lane -> out[16 j + q]
            //   B: radix 16, Ns = 16, item j = lane: in[j + 32 r] * W_256^{r (j mod 16)} -> out[256 (j/16) + j%16 + 16 q]
            const int hi = lane >> 4, lo = lane & 15;
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                fft_reg<16, FFT_FWD, 0, 16>(w);
                __syncwarp();
                float2 *o = pass == 0 ? buf + 17 * lane : buf + 272 * hi + lo;
                if (pass == 0) {
                    static_for<16>([&](auto QQ) {
                        constexpr int q = decltype(QQ)::value;
                        o[q] = w[bitrev(q, 4)];
                    });
                } else {
                    static_for<16>([&](auto QQ) {
                        constexpr int q = decltype(QQ)::value;
                        o[17 * q] = w[bitrev(q, 4)];
                    });
                }
                __syncwarp();
                if (pass == 1) break;
                static_for<16>([&](auto RR) {
                    constexpr int r = decltype(RR)::value;
                    w[r] = buf[lane + hi + 34 * r];
                });
                static_for<15>([&](auto RR) {
                    constexpr int r = decltype(RR)::value + 1;
                    const float2 t = S.twB[r - 1][lane];
                    w[r] = tw_mul<FFT_FWD>(w[r], t.x, t.y);
                });
            }
            // pass C (radix 2, Ns = 256: out[j + 256 q] = in[j] + (-1)^q W_512^j in[j + 256]) is not run as a pass:
            // stage C evaluates it for the kept samples only -- one butterfly output per kept sample instead
            // of 512 per block, and one exchange (store, barrier, load) less.
        }

        // ------------------------------------------------------------------ stage C: decimate
        // buf holds swap(z) in natural order (pruned: swap(z[2m]) at m).  Keep z[g], (g mod DB) = D*i, i < M.
        const uint32_t g0 = prm.z0 + s0;
        const uint32_t p0 = g0 & db_mask;
        const uint32_t o0 = udiv_small(p0 + prm.D - 1u, prm.D, prm.inv_d);  // first kept output index in the block
        const uint32_t pos0 = o0 * prm.D - p0;                               // its position in this 1024-block
        uint32_t cnt = 0;
        if (pos0 < 1024u && o0 < prm.M) {
            cnt = udiv_small(1023u - pos0, prm.D, prm.inv_d) + 1u;
            if (cnt > prm.M - o0) cnt = prm.M - o0;
        }
        float2 *out = dst + (size_t)(g0 >> prm.db_log2) * prm.M + o0;
        if (prune2) {
            // buf holds the two interleaved 256-point halves of the folded transform (padded a + a/16, swapped
            // re/im): z[2m] = swap(in[j] + (-1)^q W_512^j in[j + 256]), j = m mod 256, q = m div 256
            for (uint32_t k = lane; k < cnt; k += 32u) {
                const uint32_t m = (pos0 + k * prm.D) >> 1;
                const uint32_t j = m & 255u;
                const uint32_t idx = j + (j >> 4);
                const float2 a = buf[idx], b = buf[idx + 272u];
                const float2 t = S.twC[j >> 5][j & 31u];
                const float2 bq = tw_mul<FFT_FWD>(b, t.x, t.y);
                const float2 z = (m & 256u) ? sub2(a, bq) : add2(a, bq);
                out[k] = SPLIT ? cmul(make_float2(z.y, z.x), blk) : make_float2(z.y, z.x);
            }
        } else {
            for (uint32_t k = lane; k < cnt; k += 32u) {
                const uint32_t pp = pos0 + k * prm.D;
                const float2 z = buf[pp + (pp >> 5)];
                out[k] = make_float2(z.y, z.x);
            }
        }
    }
    // a launch that was allowed to start early does not FINISH before its predecessors have (and
    // their writes are visible): later stream operations may rely on "this done => all earlier done"
    if (lane == 0 && warp == 0) overlap_join(prm.done);
}

template <int FMT, bool BATCH, bool LSB = false, bool SPLIT = false>
static int launch_one(hzsdr_ctx *ctx, const ChainParams &prm_in, const NcoTable &nco) {
    ChainParams prm = prm_in;
    static PerDevice attr_set;  // value = resident CTAs per SM on that device
    const size_t smem = sizeof(Chain1024Smem);
    int rc = attr_set.once(ctx->device, [&](int &occ_out) {
        HZ_CUDA(cudaFuncSetAttribute((const void *)k_chain1024<FMT, BATCH, LSB, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int o = 0;
        HZ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, (const void *)k_chain1024<FMT, BATCH, LSB, SPLIT>, kC1024Threads, smem));
        occ_out = o > 0 ? o : 1;
        return (int)HZSDR_OK;
    });
    if (rc) return rc;
    const int occ = attr_set.get(ctx->device);
    const size_t blocks = BATCH ? (size_t)prm.nblocks * prm.nstreams : prm.nblocks;
    size_t need = (blocks + kC1024Warps - 1) / kC1024Warps;
    size_t cap = (size_t)ctx->sm_count * occ;
    // Equal shares: with w = ceil(need / cap) rounds, ceil(need / w) CTAs do exactly w blocks per warp
    // (no CTA pays the table prologue for a single block); under programmatic dependent launch the
    // slots this leaves idle are taken by the next buffer's first CTAs.
    const size_t rounds = (need + cap - 1) / cap;
    const int grid = (int)((need + rounds - 1) / rounds);
    if (BATCH && SPLIT) {  // the CTA's contiguous range: `rounds` groups of kC1024Warps blocks
        if (prm.nblocks % kC1024Warps) return fail(HZSDR_ERR_INVALID, "chain1024 batch: blocks per stream must be a multiple of %d", kC1024Warps);
        prm.per_cta = (uint32_t)(rounds * kC1024Warps);
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kC1024Threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    // see griddepcontrol in the kernel.  A batched launch reads descriptors copied just before it and
    // writes through pointers the host side does not see here: it stays ordered.
    constexpr int sb = FMT == HZSDR_FORMAT_C64 ? 8 : (FMT == HZSDR_FORMAT_I16 ? 4 : 2);
    const bool may = chain_may_overlap(ctx, prm, 1024u, sb);  // batched: spans are meaningless, only the slot is used
    overlap_launch_config(cfg, attr, BATCH ? false : may);
    if (BATCH) {  // and nothing may overlap what it writes: the next overlappable launch goes out serialised
        ctx->overlap.n = 0;
        ctx->overlap_broken();
    }
    HZ_CUDA(cudaLaunchKernelEx(&cfg, k_chain1024<FMT, BATCH, LSB, SPLIT>, prm, nco));
    return HZSDR_OK;
}

// prm.tw must point at the table built by chain1024_twiddles(); prm.H at the 1024-bin filter.
int launch_chain1024(hzsdr_ctx *ctx, int fmt, const ChainParams &prm, const NcoTable &nco) {
    if (prm.split) {  // prm.tw carries the split table for prm.dp_nom (even decimation factor, raw input)
        switch (fmt) {
            case HZSDR_FORMAT_U8: return launch_one<HZSDR_FORMAT_U8, false, false, true>(ctx, prm, nco);
            case HZSDR_FORMAT_I8: return launch_one<HZSDR_FORMAT_I8, false, false, true>(ctx, prm, nco);
            default:
                return prm.lsb_shift ? launch_one<HZSDR_FORMAT_I16, false, true, true>(ctx, prm, nco)
                                     : launch_one<HZSDR_FORMAT_I16, false, false, true>(ctx, prm, nco);
        }
    }
    switch (fmt) {
        case HZSDR_FORMAT_C64: return launch_one<HZSDR_FORMAT_C64, false>(ctx, prm, nco);
        case HZSDR_FORMAT_U8: return launch_one<HZSDR_FORMAT_U8, false>(ctx, prm, nco);
        case HZSDR_FORMAT_I8: return launch_one<HZSDR_FORMAT_I8, false>(ctx, prm, nco);
        default:
            return prm.lsb_shift ? launch_one<HZSDR_FORMAT_I16, false, true>(ctx, prm, nco)
                                 : launch_one<HZSDR_FORMAT_I16, false>(ctx, prm, nco);
    }
}

int launch_chain1024_batch(hzsdr_ctx *ctx, int fmt, const ChainParams &prm) {
    static NcoTable empty{};  // unused in batch mode; a 2.7 KB parameter keeps one kernel signature
    if (prm.split) {  // every StreamDesc carries its split table and phase step
        switch (fmt) {
            case HZSDR_FORMAT_U8: return launch_one<HZSDR_FORMAT_U8, true, false, true>(ctx, prm, empty);
            case HZSDR_FORMAT_I8: return launch_one<HZSDR_FORMAT_I8, true, false, true>(ctx, prm, empty);
            default:
                return prm.lsb_shift ? launch_one<HZSDR_FORMAT_I16, true, true, true>(ctx, prm, empty)
                                     : launch_one<HZSDR_FORMAT_I16, true, false, true>(ctx, prm, empty);
        }
    }
    switch (fmt) {
        case HZSDR_FORMAT_U8: return launch_one<HZSDR_FORMAT_U8, true>(ctx, prm, empty);
        case HZSDR_FORMAT_I8: return launch_one<HZSDR_FORMAT_I8, true>(ctx, prm, empty);
        default:
            return prm.lsb_shift ? launch_one<HZSDR_FORMAT_I16, true, true>(ctx, prm, empty)
                                 : launch_one<HZSDR_FORMAT_I16, true>(ctx, prm, empty);
    }
}

void chain1024_twiddles(float2 *host_out /* kChain1024TableLen = 32*32 + 15*32 + 8*32 */) {
    auto w = [](double num, double den) {
        const double a = 2.0 * M_PI * num / den;
        return make_float2((float)cos(a), (float)sin(a));
    };
    float2 *tw = host_out, *twB = host_out + kC1024TwFull, *twC = twB + kC1024TwB;
    for (int r = 0; r < 32; r++)
        for (int lane = 0; lane < 32; lane++) tw[r * 32 + lane] = w(r * lane, 1024.0);
    for (int r = 1; r < 16; r++)
        for (int lane = 0; lane < 32; lane++) twB[(r - 1) * 32 + lane] = w(r * (lane & 15), 256.0);
    for (int i = 0; i < 8; i++)
        for (int lane = 0; lane < 32; lane++) twC[i * 32 + lane] = w(lane + 32 * i, 512.0);
}

// The same tables for every stream of a batched launch, built on the device just before it (one CTA
// per stream; fp64 sincospi of the exact fixed-point phase): 512 streams x 8 KB take a few microseconds,
// so the channelizer does not cache them.
__global__ void __launch_bounds__(256) k_split_tables(const StreamDesc *__restrict__ streams, float scale) {
    const StreamDesc &sd = streams[blockIdx.x];
    float2 *out = const_cast<float2 *>(sd.tw);
    const uint64_t dp = sd.dp_nom;
    for (int i = threadIdx.x; i < 1024; i += 256) {
        const int r = i >> 5, lane = i & 31;
        // multiplier W_1024^{r lane} A_r = e^{2 pi i (r dp / 2^64 - r lane / 1024)}
        const double turns = (double)((uint64_t)r * dp) * 5.421010862427522e-20 - (double)(r * lane) * (1.0 / 1024.0);
        double sn, c;
        sincospi(2.0 * turns, &sn, &c);
        out[i] = make_float2((float)(c * (double)scale), (float)(sn * (double)scale));
    }
}

int launch_split_tables(hzsdr_ctx *ctx, const StreamDesc *streams_dev, uint32_t nstreams, float scale) {
    if (nstreams == 0) return HZSDR_OK;
    k_split_tables<<<nstreams, 256, 0, ctx->stream>>>(streams_dev, scale);
    HZ_CHECK_LAUNCH();
    return HZSDR_OK;
}

// The first 32 x 32 entries of the table for a SPLIT launch.  After the exchange that follows the
// first pass, register r of a lane holds the partial transform of the samples l + 32 r', l = r: the
// row index is the sample's position inside its group of 32, so row r carries A_r = e^{i 2 pi r
// dp_nom / 2^64}.  Entries are the multipliers themselves: W_1024^{r lane} A_r scale, W = e^{-2 pi i / 1024}.
void chain1024_split_twiddles(float2 *host_out /* 32*32 */, uint64_t dp_nom, float scale) {
    for (int r = 0; r < 32; r++) {
        const double turns = ldexp((double)((uint64_t)r * dp_nom), -64);  // r * dp_nom wraps mod 2^64, as on the device
        const double ax = cos(2.0 * M_PI * turns), ay = sin(2.0 * M_PI * turns);
        for (int lane = 0; lane < 32; lane++) {
            const double a = 2.0 * M_PI * (double)(r * lane) / 1024.0;
            const double c = cos(a), sn = sin(a);
            // (c - i s)(ax + i ay) = (c ax + s ay) - i (s ax - c ay)
            host_out[r * 32 + lane] = make_float2((float)((c * ax + sn * ay) * scale), (float)((c * ay - sn * ax) * scale));
        }
    }
}

}  // namespace hz
