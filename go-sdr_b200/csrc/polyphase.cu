// polyphase.cu -- fused polyphase decimator on a RAW stream:
//     Convert -> Shift (NCO, stream/shifter.go:66-85 semantics) -> FIR with real taps -> keep every D-th sample
// in one kernel; only the kept outputs are computed.
//
// AN EXTENSION (BASELINE north_star's "fused polyphase FIR+decimate"): the reference has no
// FIR -- its ConvolutionReader is block-circular (stream/convolution.go:57-81).  Definition, as for hzsdr_fir_*:
//     z[n] = sum_{k < T} h[k] * y[n-k]   over the whole stream, y[n < 0] = 0,      out[i] = z[D*i]
// (continuous decimation phase; history and the NCO time carried between calls), checked against a complex128
// FIR of the oracle's Convert + Shift output.
//
// Polyphase form: tap k = D q + p, so  z[D i] = sum_p sum_q h[D q + p] * u_p[i - q],  u_p[j] = y[D j - p]:
// D short FIRs on D decimated sequences.  A PAIR of warps owns a tile of 256 outputs (8 per lane; the rows of a
// tile are ~26 KB of shared memory, so one warp per tile would leave an SM with 8 resident warps and every
// dependency stall exposed -- 128 against 200+ Gsamples/s):
//   stage A  the tile's D (256 + Q - 1) input samples are read once (coalesced raw loads), converted, mixed (one
//            exact rotation per 8 steps of a lane's stride-32 walk, a constant complex step in between) and
//            scattered into the warp's shared memory BY PHASE: row p holds u_p, padded x + x/8 so that lanes that
//            are 8 outputs apart hit different banks;
//   stage B  per phase, a lane slides a 15-sample register window down its row, 8 taps at a time: 8 new samples
//            (LDS.64) + two broadcast LDS.128 of taps feed 64 packed FMAs (acc[r] += tap * sample on (re, im)
//            pairs) -- every sample is loaded once per phase and used by up to 8 outputs;
//   store    the two warps take the even / odd phases, swap half of their partial sums through shared memory and
//            each stores 4 consecutive outputs per lane.
// A pair meets at two 64-thread named barriers per tile; pairs never synchronise with each other.  Work per input sample: T/D packed FMAs + ~6 for stage A, against the
// FFT chain's ~24 at N = 1024: the polyphase form wins below T/D ~ 16 and loses above (hzsdr_chain_* with
// overlap_save_taps is the long-filter form).
#include <vector>

#include "common.cuh"
#include "nco.cuh"
#include "poly_host.h"

namespace hz {

constexpr int kPolyMaxPairs = 4;       // warp pairs (tiles in flight) per CTA

struct PolyParams {
    const uint8_t *src;    // the call's raw samples: launch coordinates [hist_len, n_ext)
    const uint8_t *hist;   // carried raw history: launch coordinates [0, hist_len)
    float2 *dst;
    const float *taps;     // [D][qpad]: taps[p][q] = h[D q + p], zero beyond T
    const NcoSegment *segs;  // device table, or nullptr: the segments travel in the kernel's second parameter
    int nsegs;
    uint8_t *hist_out;       // the stream's last hist_len raw samples after this call (the other half of the ping-pong)
    uint32_t D, Q, qpad, npairs, half_last, joff, k0, row;
    uint32_t hist_len, n_ext, silent;  // launch coordinates < silent are silence (stream start)
    int64_t e_first;                   // launch coordinate of the sample behind output 0
    uint32_t cnt, ntiles;
    int lsb_shift;
};

template <int FMT, bool LSB>
__device__ __forceinline__ uint32_t poly_load(const uint8_t *base, uint32_t j, int lsb_shift) {
    if constexpr (FMT == HZSDR_FORMAT_I16) {
        uint32_t w = __ldg(reinterpret_cast<const uint32_t *>(base) + j);
        if constexpr (LSB) w = ((w & 0xffff0000u) << lsb_shift) | (((w & 0xffffu) << lsb_shift) & 0xffffu);
        return w;
    } else {
        return (uint32_t)__ldg(reinterpret_cast<const uint16_t *>(base) + j);
    }
}


template <int FMT, bool LSB>
__global__ void __launch_bounds__(64 * kPolyMaxPairs, 2) k_polyphase_chain(const __grid_constant__ PolyParams prm,
                                                                            const __grid_constant__ NcoTable tab) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 6, half_id = (threadIdx.x >> 5) & 1, nwarp = blockDim.x >> 6;  // warp = pair index
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + warp) : "memory"); };
    const uint32_t D = prm.D, row = prm.row, joff = prm.joff;
    const uint32_t wrap_step = (D - 1u) * row + 1u;  // from row 0 of one slot to row D - 1 of the next (plus 1 more over a padding slot)
    const float inv_d = __uint_as_float(__float_as_uint(1.0f / (float)D) - 1u);  // 1/D rounded down
    float *taps_s = reinterpret_cast<float *>(smem_raw);
    float2 *X = reinterpret_cast<float2 *>(smem_raw + (((size_t)D * prm.qpad * sizeof(float) + 15) & ~(size_t)15)) + (size_t)warp * 256;  // [2][32][4]
    float2 *U = reinterpret_cast<float2 *>(smem_raw + (((size_t)D * prm.qpad * sizeof(float) + 15) & ~(size_t)15)) + (size_t)nwarp * 256 +
                (size_t)warp * D * row;

    for (uint32_t i = threadIdx.x; i < D * prm.qpad; i += blockDim.x) taps_s[i] = __ldg(prm.taps + i);
    for (uint32_t i = lane + 32 * half_id; i < D * row; i += 64u) U[i] = make_float2(0.f, 0.f);  // (slots stage A never writes stay zero)
    __syncthreads();

    const SegView view{prm.segs ? prm.segs : tab.seg, prm.nsegs};
    // carry: the last CTA writes the history the NEXT call will read -- launch coordinates [n_ext - hist_len, n_ext) of this
    // call, into the half of the ping-pong this launch does not read (launches of one decimator are stream-ordered)
    if (blockIdx.x == gridDim.x - 1u && prm.hist_out) {
        constexpr int kSbC = RawTraits<FMT>::bytes;
        using Word = typename std::conditional<kSbC == 4, uint32_t, uint16_t>::type;
        Word *o = reinterpret_cast<Word *>(prm.hist_out);
        const uint32_t first = prm.n_ext - prm.hist_len;
        for (uint32_t i = threadIdx.x; i < prm.hist_len; i += blockDim.x) {
            const uint32_t e = first + i;
            o[i] = e < prm.hist_len ? reinterpret_cast<const Word *>(prm.hist)[e] : __ldg(reinterpret_cast<const Word *>(prm.src) + (e - prm.hist_len));
        }
    }
    const float sc = RawTraits<FMT>::scale();
    const uint32_t a_len = D * ((uint32_t)kPolyOT + prm.Q - 1u);
    constexpr int kSb = RawTraits<FMT>::bytes;
    // samples between the source pointer's last 16-byte boundary and its first sample (the API asks for sample alignment only)
    const uint32_t src_phase = (uint32_t)((reinterpret_cast<uintptr_t>(prm.src) & 15u) / kSb) & 7u;

    for (uint32_t tile = blockIdx.x * nwarp + warp; tile < prm.ntiles; tile += gridDim.x * nwarp) {
        // ------------------------------------------------------------------ stage A: raw -> mixed samples, by phase
        // a = e - e_lo walks the tile's inputs in stream order; sample a belongs to phase p = D - 1 - a % D, slot a / D.
        // A lane takes GROUPS of 8 consecutive samples that start on a 16-byte boundary of the source (one or two
        // 128-bit loads): one exact rotation for the group's first sample, the other seven by one complex multiply
        // with e^{i k dP}.  Groups that touch the history, the buffer's ends or two accumulator segments go sample
        // by sample.
        const int32_t e_lo = (int32_t)(prm.e_first + (int64_t)D * ((int64_t)kPolyOT * tile - (int64_t)(prm.Q - 1u)) - (int64_t)(D - 1u));
        {   // the next tile's raw bytes -> L2 while this one is computed
            const uint32_t nt = tile + gridDim.x * nwarp;
            if (nt < prm.ntiles) {
                const int64_t ne = (int64_t)e_lo + (int64_t)D * kPolyOT * (int64_t)(gridDim.x * nwarp) - (int64_t)prm.hist_len;
                const int64_t lo = ne > 0 ? ne : 0;
                const uint8_t *pb = prm.src + lo * kSb;
                const uint32_t bytes = a_len * kSb;
                for (uint32_t o = 128u * (lane + 32 * half_id); o < bytes + 128u && lo * kSb + o < (int64_t)(prm.n_ext - prm.hist_len) * kSb; o += 64u * 128u)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(pb + o));
            }
        }
        // first group start: the 16-byte boundary of the source at or below e_lo (in launch coordinates)
        const int32_t mis = (int32_t)((((int64_t)e_lo - (int64_t)prm.hist_len) + (int64_t)src_phase) & 7);  // samples past the boundary
        const int32_t e_al = e_lo - mis;
        const uint32_t ngroups = (a_len + (uint32_t)mis + 7u) >> 3;
        NcoCursor cur;
        float2 wk[7];
        uint64_t wk_dp = ~0ull;
        for (uint32_t g = lane + 32 * half_id; g < ngroups; g += 64u) {
            const int32_t e0 = e_al + (int32_t)(8u * g);
            const int32_t a = e0 - e_lo;                 // negative for a tile's first group: those samples belong to no slot
            const uint32_t a0 = a > 0 ? (uint32_t)a : 0u;  // the group's first sample inside the tile
            // a0 / D for a0 < 2^24: float estimate (never too large: inv_d is rounded down) + one fix-up
            uint32_t jj = __float2uint_rz(__uint2float_rz(a0) * inv_d);
            if (a0 - jj * D >= D) jj++;
            uint32_t rem = a0 - jj * D;
            const bool whole = a >= 0 && (uint32_t)a + 8u <= a_len;  // (a tile's first and last group may stick out of it)
            const bool inside = e0 >= (int32_t)prm.hist_len && e0 >= (int32_t)prm.silent && e0 + 8 <= (int32_t)prm.n_ext;
            bool fast = inside;
            if (fast) {
                cur.seek(view, (uint32_t)e0);
                fast = (uint32_t)e0 + 8u <= cur.end && cur.dp != 0ull;
            }
            float2 y[8];
            if (fast) {
                uint32_t w[8 * kSb / 4];
                const uint4 *gp = reinterpret_cast<const uint4 *>(prm.src + (size_t)(e0 - (int32_t)prm.hist_len) * kSb);
                const uint4 q0 = __ldg(gp);
                w[0] = q0.x, w[1] = q0.y, w[2] = q0.z, w[3] = q0.w;
                if constexpr (kSb == 4) {
                    const uint4 q1 = __ldg(gp + 1);
                    w[4] = q1.x, w[5] = q1.y, w[6] = q1.z, w[7] = q1.w;
                }
                if (cur.dp != wk_dp) {
#pragma unroll
                    for (int k = 0; k < 7; k++) wk[k] = nco_rot((uint64_t)(k + 1) * cur.dp);
                    wk_dp = cur.dp;
                }
                const float2 r0 = mul2(nco_rot(cur.phase((uint32_t)e0)), make_float2(sc, sc));
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    float2 x;
                    if constexpr (kSb == 4) {
                        uint32_t ww = w[k];
                        if constexpr (LSB) ww = ((ww & 0xffff0000u) << prm.lsb_shift) | (((ww & 0xffffu) << prm.lsb_shift) & 0xffffu);
                        x = RawTraits<FMT>::unscaled2(ww);
                    } else {
                        x = (k & 1) ? RawTraits<FMT>::unscaled_hi(w[k >> 1]) : RawTraits<FMT>::unscaled(w[k >> 1]);
                    }
                    y[k] = cmul(x, k == 0 ? r0 : cmul(r0, wk[k - 1]));
                }
            } else {
#pragma unroll 1
                for (int k = 0; k < 8; k++) {
                    const int32_t e = e0 + k;
                    float2 v = make_float2(0.f, 0.f);
                    if (e >= (int32_t)prm.silent && e < (int32_t)prm.n_ext && e >= e_lo) {
                        const uint32_t raw = e < (int32_t)prm.hist_len ? poly_load<FMT, LSB>(prm.hist, (uint32_t)e, prm.lsb_shift)
                                                                        : poly_load<FMT, LSB>(prm.src, (uint32_t)(e - (int32_t)prm.hist_len), prm.lsb_shift);
                        cur.seek(view, (uint32_t)e);
                        v = cmul(RawTraits<FMT>::unscaled2(raw), mul2(nco_rot(cur.phase((uint32_t)e)), make_float2(sc, sc)));
                    }
                    // (dynamic index into y: through the thread's own slots of row 0 would cost a second pass; store at once)
                    const int32_t ak = a + k;
                    if (ak >= 0 && (uint32_t)ak < a_len) {
                        const uint32_t j2 = (uint32_t)ak / D, r2 = (uint32_t)ak - j2 * D;
                        U[(D - 1u - r2) * row + (uint32_t)poly_pos((int)(j2 + prm.joff))] = v;
                    }
                }
                continue;
            }
            // scatter: the phase row steps down with every sample and wraps to the next slot after D of them
            const uint32_t slot = jj + joff;
            if (D >= 8u && whole) {  // at most one wrap inside the group: eight independent addresses off one base
                float2 *base = U + (D - 1u - rem) * row + slot + (slot >> 3);
                const uint32_t wrap_add = D * row + 1u + ((((slot + 1u) & 7u) == 0u) ? 1u : 0u);  // row 0 -> row D - 1 of the next slot
                const uint32_t first_wrapped = D - rem;                                             // samples k >= this are in the next slot
#pragma unroll
                for (int k = 0; k < 8; k++) base[((uint32_t)k >= first_wrapped ? wrap_add : 0u) - (uint32_t)k * row] = y[k];
            } else if (D >= 8u) {
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const uint32_t rk = rem + (uint32_t)(a + k) - a0;  // (a + k - a0 = k for a whole group)
                    const bool wrap = rk >= D;
                    const uint32_t sk = slot + (wrap ? 1u : 0u);
                    if (whole || (a + k >= 0 && (uint32_t)(a + k) < a_len)) U[(D - 1u - (wrap ? rk - D : rk)) * row + sk + (sk >> 3)] = y[k];
                }
            } else if (!whole) {  // (short rows, ragged group: sample by sample)
#pragma unroll 1
                for (int k = 0; k < 8; k++) {
                    const int32_t ak = a + k;
                    if (ak >= 0 && (uint32_t)ak < a_len) {
                        const uint32_t j2 = (uint32_t)ak / D, r2 = (uint32_t)ak - j2 * D;
                        U[(D - 1u - r2) * row + (uint32_t)poly_pos((int)(j2 + joff))] = y[k];
                    }
                }
            } else {
                uint32_t sk = slot;
                float2 *up = U + (D - 1u - rem) * row + (sk + (sk >> 3));
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    *up = y[k];
                    rem++;
                    const bool wrap = rem == D;
                    sk += wrap ? 1u : 0u;
                    up += wrap ? (ptrdiff_t)wrap_step + (((sk & 7u) == 0u) ? 1 : 0) : -(ptrdiff_t)row;
                    rem = wrap ? 0u : rem;
                }
            }
        }
        pair_sync();

        // ------------------------------------------------------------------ stage B: D short FIRs from registers (this warp: every other phase)
        float2 acc[kPolyR];
#pragma unroll
        for (int r = 0; r < kPolyR; r++) acc[r] = make_float2(0.f, 0.f);
#pragma unroll 1
        for (uint32_t p = half_id; p < D; p += 2u) {
            // pair m covers taps q = 8 m .. 8 m + 7; V[x] = u_p slot 8 lane + x + 8 (k0 - m) (+ joff), and output r, tap
            // 8 m + c meet at x = r + 7 - c
            const float2 *u = U + p * row + 9 * (lane + (int)prm.k0);
            const float4 *tp = reinterpret_cast<const float4 *>(taps_s + p * prm.qpad);
            // The window of pair m is [L | H]: L = slots x = 0..7 (loaded now), H = x = 8..14 = the L of pair m - 1.  Two register
            // arrays take turns as L and H, so nothing is moved.
            float2 P[8], Qv[8];
#pragma unroll
            for (int x = 0; x < 7; x++) Qv[x] = u[x + 9];  // x = 8..14 of pair 0 (padded offset x + 1)
            auto pair_step = [&](float2 (&L)[8], const float2 (&Hh)[8], bool half) {
                const float4 t0 = tp[0];
#pragma unroll
                for (int x = 4; x < 8; x++) L[x] = u[x];
                const float tv0[4] = {t0.x, t0.y, t0.z, t0.w};
                if (!half) {
#pragma unroll
                    for (int x = 0; x < 4; x++) L[x] = u[x];
                }
#pragma unroll
                for (int c = 0; c < 4; c++)
#pragma unroll
                    for (int r = 0; r < kPolyR; r++) {
                        const int x = r + 7 - c;
                        acc[r] = fma2(x < 8 ? L[x] : Hh[x - 8], make_float2(tv0[c], tv0[c]), acc[r]);
                    }
                if (!half) {
                    const float4 t1 = tp[1];
                    const float tv1[4] = {t1.x, t1.y, t1.z, t1.w};
#pragma unroll
                    for (int c = 0; c < 4; c++)
#pragma unroll
                        for (int r = 0; r < kPolyR; r++) {
                            const int x = r + 3 - c;
                            acc[r] = fma2(x < 8 ? L[x] : Hh[x - 8], make_float2(tv1[c], tv1[c]), acc[r]);
                        }
                }
                u -= 9;
                tp += 2;
            };
#pragma unroll 1
            for (uint32_t m = 0; m < prm.npairs; m += 2u) {
                pair_step(P, Qv, prm.half_last != 0u && m + 1u == prm.npairs);
                if (m + 1u < prm.npairs) pair_step(Qv, P, prm.half_last != 0u && m + 2u == prm.npairs);
            }
        }

        // ------------------------------------------------------------------ swap half of the partial sums, store
        // warp h of the pair finishes outputs r = 4 h .. 4 h + 3 of every lane: it hands the other four to its partner
        {
            float4 *xo = reinterpret_cast<float4 *>(X + (half_id * 32 + lane) * 4);
            if (half_id == 0) {
                xo[0] = make_float4(acc[4].x, acc[4].y, acc[5].x, acc[5].y);
                xo[1] = make_float4(acc[6].x, acc[6].y, acc[7].x, acc[7].y);
            } else {
                xo[0] = make_float4(acc[0].x, acc[0].y, acc[1].x, acc[1].y);
                xo[1] = make_float4(acc[2].x, acc[2].y, acc[3].x, acc[3].y);
            }
        }
        pair_sync();  // (both warps are done with the rows as well: the next tile's stage A may overwrite them)
        {
            const float4 *xi = reinterpret_cast<const float4 *>(X + ((half_id ^ 1) * 32 + lane) * 4);
            const float4 a = xi[0], b = xi[1];
            float2 o[4];
            if (half_id == 0) {
                o[0] = add2(acc[0], make_float2(a.x, a.y)), o[1] = add2(acc[1], make_float2(a.z, a.w));
                o[2] = add2(acc[2], make_float2(b.x, b.y)), o[3] = add2(acc[3], make_float2(b.z, b.w));
            } else {
                o[0] = add2(acc[4], make_float2(a.x, a.y)), o[1] = add2(acc[5], make_float2(a.z, a.w));
                o[2] = add2(acc[6], make_float2(b.x, b.y)), o[3] = add2(acc[7], make_float2(b.z, b.w));
            }
            const uint32_t o0 = (uint32_t)kPolyOT * tile + (uint32_t)kPolyR * lane + 4u * half_id;
            if (o0 + 4u <= prm.cnt && ((reinterpret_cast<uintptr_t>(prm.dst + o0) & 15u) == 0u)) {
                float4 *d4 = reinterpret_cast<float4 *>(prm.dst + o0);
                d4[0] = make_float4(o[0].x, o[0].y, o[1].x, o[1].y);
                d4[1] = make_float4(o[2].x, o[2].y, o[3].x, o[3].y);
            } else {
#pragma unroll
                for (int r = 0; r < 4; r++)
                    if (o0 + (uint32_t)r < prm.cnt) prm.dst[o0 + r] = o[r];
            }
        }
        // (the scratch is rewritten only after the next tile's first barrier, which the partner reaches after reading it)
    }
}

}  // namespace hz

using namespace hz;

struct hzsdr_polyphase {
    hzsdr_ctx *ctx = nullptr;
    int fmt = 0, lsb_bits = 0;
    size_t ntaps = 0, hist = 0;
    unsigned D = 1;
    double shift_hz = 0.0;
    hzsdr_nco nco{};
    uint32_t Q = 0, qpad = 0, npairs = 0, half_last = 0, joff = 0, k0 = 0, row = 0;
    int warps = 0;
    size_t smem = 0;
    float *taps = nullptr;          // device, [D][qpad]
    uint8_t *hist_raw[2] = {};      // device, ping-pong: the last `hist` raw samples of the stream
    int cur = 0;
    std::vector<HostSeg> tail;      // their accumulator segments, coordinates [0, hist)
    uint64_t pos = 0;               // samples consumed so far
    // segment tables: pinned staging (the host never rewrites a slot a pending copy reads) + one device image
    static constexpr int kStages = 3;
    NcoSegment *seg_host[kStages] = {};
    cudaEvent_t seg_done[kStages] = {};
    bool seg_used[kStages] = {};
    NcoSegment *seg_dev = nullptr;
    size_t seg_cap = 0;
    uint64_t calls = 0;
};

extern "C" int hzsdr_polyphase_destroy(hzsdr_polyphase *f) {
    if (!f) return HZSDR_OK;
    HZ_ENTER(f->ctx);
    cudaStreamSynchronize(f->ctx->stream);
    if (f->taps) cudaFree(f->taps);
    for (auto *p : f->hist_raw)
        if (p) cudaFree(p);
    for (int i = 0; i < hzsdr_polyphase::kStages; i++) {
        if (f->seg_host[i]) cudaFreeHost(f->seg_host[i]);
        if (f->seg_done[i]) cudaEventDestroy(f->seg_done[i]);
    }
    if (f->seg_dev) cudaFree(f->seg_dev);
    delete f;
    return HZSDR_OK;
}

template <int FMT, bool LSB>
static int poly_set_smem(size_t smem) {
    HZ_CUDA(cudaFuncSetAttribute((const void *)k_polyphase_chain<FMT, LSB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return HZSDR_OK;
}

extern "C" int hzsdr_polyphase_create(hzsdr_ctx *ctx, int src_format, uint32_t sample_rate, double shift_hz, const float *taps,
                                      size_t ntaps, unsigned decimate, int i16_lsb_bits, hzsdr_polyphase **out) {
    HZ_ENTER(ctx);
    if (!out || !taps || ntaps == 0 || decimate == 0 || sample_rate == 0) return fail(HZSDR_ERR_INVALID, "hzsdr_polyphase_create: bad arguments");
    *out = nullptr;
    if (src_format != HZSDR_FORMAT_U8 && src_format != HZSDR_FORMAT_I8 && src_format != HZSDR_FORMAT_I16)
        return fail(HZSDR_ERR_FORMAT_UNKNOWN, "hzsdr_polyphase_create: raw source format expected, got %d", src_format);
    if (i16_lsb_bits < 0 || i16_lsb_bits > 16 || (i16_lsb_bits && src_format != HZSDR_FORMAT_I16))
        return fail(HZSDR_ERR_INVALID, "hzsdr_polyphase_create: i16_lsb_bits = %d", i16_lsb_bits);
    if (decimate > (1u << 16) || ntaps > (1u << 20)) return fail(HZSDR_ERR_UNSUPPORTED, "hzsdr_polyphase_create: decimate / ntaps out of range");
    hzsdr_polyphase *f = new hzsdr_polyphase();
    f->ctx = ctx;
    f->fmt = src_format;
    f->lsb_bits = i16_lsb_bits;
    f->ntaps = ntaps;
    f->hist = ntaps - 1;
    f->D = decimate;
    f->shift_hz = shift_hz;
    f->nco.sample_rate = sample_rate;
    f->nco.ts = 0.0;
    const uint32_t D = decimate;
    const PolyPlan plan = poly_plan(ntaps, D);  // poly_host.h
    f->Q = plan.Q, f->npairs = plan.npairs, f->qpad = plan.qpad, f->half_last = plan.half_last;
    f->joff = plan.joff, f->k0 = plan.k0, f->row = plan.row;
    // tiles in flight (warp pairs) per CTA so that two CTAs fit an SM's shared memory
    const size_t taps_bytes = (((size_t)D * f->qpad * sizeof(float)) + 15) & ~(size_t)15;
    const size_t per_tile = (size_t)D * f->row * sizeof(float2);
    const size_t budget = (size_t)ctx->prop.sharedMemPerMultiprocessor / 2 - 2048;
    const size_t per_warp = per_tile + 256 * sizeof(float2);  // rows + the pair's exchange scratch
    int w = kPolyMaxPairs;
    while (w > 1 && taps_bytes + per_warp * w > budget) w >>= 1;
    if (taps_bytes + per_warp * w > (size_t)ctx->prop.sharedMemPerBlockOptin)
        return delete f, fail(HZSDR_ERR_UNSUPPORTED, "hzsdr_polyphase_create: %zu taps / decimate %u need %zu bytes of shared memory per tile "
                              "(use hzsdr_chain_* with overlap_save_taps for long filters)", ntaps, decimate, per_warp);
    f->warps = w;
    f->smem = taps_bytes + per_warp * w;
    auto bail = [&](int rc) {
        hzsdr_polyphase_destroy(f);
        return rc;
    };
    const std::vector<float> t = poly_taps_layout(taps, ntaps, plan);
    cudaError_t e = cudaMalloc((void **)&f->taps, sizeof(float) * t.size());
    if (e == cudaSuccess) e = cudaMemcpy(f->taps, t.data(), sizeof(float) * t.size(), cudaMemcpyHostToDevice);
    const size_t hb = (f->hist ? f->hist : 1) * (size_t)hzsdr_format_size(src_format);
    for (int i = 0; i < 2 && e == cudaSuccess; i++) {
        e = cudaMalloc((void **)&f->hist_raw[i], hb);
        if (e == cudaSuccess) e = cudaMemset(f->hist_raw[i], 0, hb);
    }
    if (e != cudaSuccess) return bail(fail(HZSDR_ERR_CUDA, "hzsdr_polyphase_create: %s", cudaGetErrorString(e)));
    int rc = HZSDR_OK;
    switch (src_format) {
        case HZSDR_FORMAT_U8: rc = poly_set_smem<HZSDR_FORMAT_U8, false>(f->smem); break;
        case HZSDR_FORMAT_I8: rc = poly_set_smem<HZSDR_FORMAT_I8, false>(f->smem); break;
        default: rc = i16_lsb_bits ? poly_set_smem<HZSDR_FORMAT_I16, true>(f->smem) : poly_set_smem<HZSDR_FORMAT_I16, false>(f->smem);
    }
    if (rc) return bail(rc);
    *out = f;
    return HZSDR_OK;
}

extern "C" int hzsdr_polyphase_get_ts(const hzsdr_polyphase *f, double *ts) {
    if (!f || !ts) return fail(HZSDR_ERR_INVALID, "hzsdr_polyphase_get_ts: null");
    *ts = f->nco.ts;
    return HZSDR_OK;
}
extern "C" int hzsdr_polyphase_set_ts(hzsdr_polyphase *f, double ts) {
    if (!f) return fail(HZSDR_ERR_INVALID, "hzsdr_polyphase_set_ts: null");
    f->nco.ts = ts;
    return HZSDR_OK;
}

// outputs the next n samples will produce: stream indices D*i inside [pos, pos + n)
extern "C" int hzsdr_polyphase_out_len(const hzsdr_polyphase *f, size_t n, size_t *n_out) {
    if (!f || !n_out) return fail(HZSDR_ERR_INVALID, "hzsdr_polyphase_out_len: null");
    const uint64_t g0 = (f->pos + f->D - 1) / f->D * f->D;
    *n_out = g0 < f->pos + n ? (size_t)((f->pos + n - 1 - g0) / f->D + 1) : 0;
    return HZSDR_OK;
}

extern "C" int hzsdr_polyphase_exec(hzsdr_polyphase *f, const void *src, size_t n, void *dst, size_t dst_len, size_t *n_out) {
    if (!f) return fail(HZSDR_ERR_INVALID, "hzsdr_polyphase_exec: null");
    HZ_ENTER(f->ctx);
    if (n_out) *n_out = 0;
    if (n == 0) return HZSDR_OK;
    if (n > 0x3fffffffull) return fail(HZSDR_ERR_INVALID, "hzsdr_polyphase_exec: at most 2^30-1 samples per call");
    size_t cnt = 0;
    hzsdr_polyphase_out_len(f, n, &cnt);
    if (dst_len < cnt) return fail(HZSDR_ERR_DST_TOO_SMALL, "hzsdr_polyphase_exec: %zu < %zu", dst_len, cnt);
    if (!src || (cnt && !dst)) return fail(HZSDR_ERR_INVALID, "hzsdr_polyphase_exec: null buffer");
    const int sb = hzsdr_format_size(f->fmt);
    if (((uintptr_t)src % sb) || ((uintptr_t)dst % 8)) return fail(HZSDR_ERR_INVALID, "hzsdr_polyphase_exec: misaligned buffer");
    hzsdr_ctx *ctx = f->ctx;
    cudaStream_t st = ctx->stream;
    const size_t hist = f->hist;

    // accumulator segments in launch coordinates: [0, hist) = carried history, [hist, hist + n) = this buffer
    std::vector<HostSeg> segs;
    double ts = f->nco.ts;
    build_segments(f->nco.sample_rate, n, &ts, segs);
    std::vector<HostSeg> ext;
    if (hist) {
        if (!f->tail.empty())
            ext = f->tail;
        else
            ext.push_back(HostSeg{0, hist, 0.0, 0.0});  // silence in front of the stream: any phase will do
    }
    for (HostSeg h : segs) {
        h.j0 += hist;
        ext.push_back(h);
    }
    {
        const bool in_params = ext.size() <= (size_t)kMaxSegsPerLaunch;
        NcoTable tab;
        if (in_params) {
            tab.count = (int)ext.size();
            for (size_t k = 0; k < ext.size(); k++) tab.seg[k] = to_device_segment(ext[k], 0, f->shift_hz);
        } else {
            if (ext.size() > f->seg_cap) {
                HZ_CUDA(cudaStreamSynchronize(st));
                const size_t cap = ext.size() * 2 + 64;
                for (int i = 0; i < hzsdr_polyphase::kStages; i++) {
                    if (f->seg_host[i]) cudaFreeHost(f->seg_host[i]);
                    f->seg_host[i] = nullptr;
                    f->seg_used[i] = false;
                    HZ_CUDA(cudaHostAlloc((void **)&f->seg_host[i], sizeof(NcoSegment) * cap, cudaHostAllocPortable));
                    if (!f->seg_done[i]) HZ_CUDA(cudaEventCreateWithFlags(&f->seg_done[i], cudaEventDisableTiming));
                }
                if (f->seg_dev) cudaFree(f->seg_dev);
                f->seg_dev = nullptr;
                f->seg_cap = 0;
                HZ_CUDA(cudaMalloc((void **)&f->seg_dev, sizeof(NcoSegment) * cap));
                f->seg_cap = cap;
            }
            const int stage = (int)(f->calls % hzsdr_polyphase::kStages);
            if (f->seg_used[stage]) HZ_CUDA(cudaEventSynchronize(f->seg_done[stage]));
            for (size_t k = 0; k < ext.size(); k++) f->seg_host[stage][k] = to_device_segment(ext[k], 0, f->shift_hz);
            HZ_CUDA(cudaMemcpyAsync(f->seg_dev, f->seg_host[stage], sizeof(NcoSegment) * ext.size(), cudaMemcpyHostToDevice, st));
            HZ_CUDA(cudaEventRecord(f->seg_done[stage], st));
            f->seg_used[stage] = true;
            f->calls++;
            tab.count = 0;
        }
        PolyParams prm{};
        prm.src = (const uint8_t *)src;
        prm.hist = f->hist_raw[f->cur];
        prm.dst = (float2 *)dst;
        prm.taps = f->taps;
        prm.segs = in_params ? nullptr : f->seg_dev;
        prm.nsegs = (int)ext.size();
        prm.hist_out = hist ? f->hist_raw[f->cur ^ 1] : nullptr;
        prm.D = f->D, prm.Q = f->Q, prm.qpad = f->qpad, prm.npairs = f->npairs, prm.half_last = f->half_last;
        prm.joff = f->joff, prm.k0 = f->k0, prm.row = f->row;
        prm.hist_len = (uint32_t)hist;
        prm.n_ext = (uint32_t)(hist + n);
        prm.silent = f->pos < hist ? (uint32_t)(hist - f->pos) : 0u;
        const uint64_t g0 = (f->pos + f->D - 1) / f->D * f->D;
        prm.e_first = (int64_t)(g0 - f->pos) + (int64_t)hist;
        prm.cnt = (uint32_t)cnt;
        prm.ntiles = (uint32_t)((cnt + kPolyOT - 1) / kPolyOT);  // (0 for a call that produces no output: the launch only carries the history)
        prm.lsb_shift = f->lsb_bits ? 16 - f->lsb_bits : 0;
        size_t ctas = (prm.ntiles + f->warps - 1) / f->warps;
        if (ctas < 1) ctas = 1;
        const size_t cap = (size_t)ctx->sm_count * 2;
        const int grid = (int)(ctas < cap ? ctas : cap);
        ctx->overlap.n = 0;  // (outside the overlap scheme: the next overlappable launch goes out serialised)
        ctx->overlap_broken();
        const dim3 block(64 * f->warps);  // f->warps pairs of warps
        switch (f->fmt) {
            case HZSDR_FORMAT_U8: k_polyphase_chain<HZSDR_FORMAT_U8, false><<<grid, block, f->smem, st>>>(prm, tab); break;
            case HZSDR_FORMAT_I8: k_polyphase_chain<HZSDR_FORMAT_I8, false><<<grid, block, f->smem, st>>>(prm, tab); break;
            default:
                if (f->lsb_bits)
                    k_polyphase_chain<HZSDR_FORMAT_I16, true><<<grid, block, f->smem, st>>>(prm, tab);
                else
                    k_polyphase_chain<HZSDR_FORMAT_I16, false><<<grid, block, f->smem, st>>>(prm, tab);
        }
        HZ_CHECK_LAUNCH();
    }
    // the kernel has saved the stream's last `hist` raw samples into the other half of the ping-pong; here the segments
    // that cover them
    if (hist) {
        f->cur ^= 1;
        f->tail.clear();
        for (HostSeg h : ext) {  // the part of [n, n + hist) in launch coordinates, re-based to 0
            const uint64_t h_end = h.j0 + h.count;
            if (h_end <= n) continue;
            if (h.j0 < n) {
                const uint64_t d = n - h.j0;
                if (h.step != 0.0) h.base += (double)d * h.step;
                h.j0 = n;
                h.count -= d;
            }
            h.j0 -= n;
            f->tail.push_back(h);
        }
    }
    f->pos += n;
    f->nco.ts = ts;
    if (n_out) *n_out = cnt;
    return HZSDR_OK;
}
