// beam.cuh -- the per-thread core of K8 (stream/beamform.go:148-171 -> multiply.go:46-70, add.go:115-185),
// shared by the single-GPU kernel (elementwise.cu) and the multi-GPU fused kernel (beamgroup.cu).
#pragma once
#include <type_traits>

#include "common.cuh"

namespace hz {

constexpr int kMaxBeamChans = 64;
struct BeamArgs {
    const uint8_t *chan[kMaxBeamChans];
    float2 w[kMaxBeamChans];  // weights, already multiplied by the format's conversion scale
    int nchan;
    int accumulate;  // continue a sum started by a previous launch (> kMaxBeamChans channels)
};

inline float beam_weight_scale(int src_format) {
    return src_format == HZSDR_FORMAT_U8 ? 1.0f / 127.5f : (src_format == HZSDR_FORMAT_I8 ? 0.0078125f : 1.0f / 32767.0f);
}

// acc[0..7] += sum_c w'_c * x_c[4i .. 4i+3]   (four consecutive output samples, channel order).
// The conversion scale is folded into the weight (exact for i8) and each term is accumulated with
// FMAs: per channel-sample 2 PRMT + 2 FADD + 4 FFMA.  One 8-byte (u8/i8) or 16-byte (i16) load per
// channel per thread, G channels in flight.
// chan / w: nchan raw-buffer pointers and scaled weights (kernel parameter space or device memory).
template <int FMT>
__device__ __forceinline__ void beam_quad(const uint8_t *const *chan, const float2 *w_, int nchan, size_t i, float (&acc)[8]) {
    using T = RawTraits<FMT>;
    // loads in flight per thread: 16 x 8 B (u8/i8) or 8 x 16 B (i16) = 128 B -- the kernel is bound
    // by outstanding HBM requests (62% of its stall samples were long-scoreboard with 64 B in flight)
    constexpr int G = T::bytes == 2 ? 16 : 8;
    using Raw = typename std::conditional<T::bytes == 2, uint2, uint4>::type;
    auto fma_sample = [&](float2 x, float2 w, int k) {
        acc[2 * k] = fmaf(x.x, w.x, acc[2 * k]);
        acc[2 * k] = fmaf(-x.y, w.y, acc[2 * k]);
        acc[2 * k + 1] = fmaf(x.x, w.y, acc[2 * k + 1]);
        acc[2 * k + 1] = fmaf(x.y, w.x, acc[2 * k + 1]);
    };
    auto accumulate = [&](const Raw &raw, float2 w) {
        if constexpr (T::bytes == 2) {  // raw.x, raw.y hold 4 samples
            fma_sample(T::unscaled(raw.x), w, 0);
            fma_sample(T::unscaled_hi(raw.x), w, 1);
            fma_sample(T::unscaled(raw.y), w, 2);
            fma_sample(T::unscaled_hi(raw.y), w, 3);
        } else {
            fma_sample(T::unscaled(raw.x), w, 0);
            fma_sample(T::unscaled(raw.y), w, 1);
            fma_sample(T::unscaled(raw.z), w, 2);
            fma_sample(T::unscaled(raw.w), w, 3);
        }
    };
    auto load = [&](int c) -> Raw {
        if constexpr (T::bytes == 2) {
            return ld_stream_u64(chan[c] + 8 * i);
        } else {
            return ld_stream_u128(chan[c] + 16 * i);
        }
    };
    int c = 0;
    for (; c + G <= nchan; c += G) {
        Raw v[G];
#pragma unroll
        for (int u = 0; u < G; u++) v[u] = load(c + u);
#pragma unroll
        for (int u = 0; u < G; u++) accumulate(v[u], w_[c + u]);
    }
    // the rest (a channel count that is not a multiple of G) one by one: rare, and kept this small on purpose -- a
    // cascade of G/2, G/4, ... groups here made ptxas cut the main loop's registers from 40 to 32, i.e. fewer loads in
    // flight, and cost the 64-channel kernel 10-20%.  Ranks with few channels use beam_quad2 below.
    for (; c < nchan; c++) accumulate(load(c), w_[c]);
}

// Two quads at once for a rank that holds only a few channels (nchan <= 8: a share of the 8-GPU split): all 2 x nchan
// loads are in flight before the first use, so a thread still has up to 128 B outstanding.
template <int FMT>
__device__ __forceinline__ void beam_quad2(const uint8_t *const *chan_a, const uint8_t *const *chan_b, const float2 *w_, int nchan, size_t ia,
                                           size_t ib, float (&acc_a)[8], float (&acc_b)[8]) {
    using T = RawTraits<FMT>;
    using Raw = typename std::conditional<T::bytes == 2, uint2, uint4>::type;
    auto load = [&](const uint8_t *const *chan, int c, size_t i) -> Raw {
        if constexpr (T::bytes == 2) {
            return ld_stream_u64(chan[c] + 8 * i);
        } else {
            return ld_stream_u128(chan[c] + 16 * i);
        }
    };
    auto accumulate = [&](float (&acc)[8], const Raw &raw, float2 w) {
        auto fma_sample = [&](float2 x, int k) {
            acc[2 * k] = fmaf(x.x, w.x, acc[2 * k]);
            acc[2 * k] = fmaf(-x.y, w.y, acc[2 * k]);
            acc[2 * k + 1] = fmaf(x.x, w.y, acc[2 * k + 1]);
            acc[2 * k + 1] = fmaf(x.y, w.x, acc[2 * k + 1]);
        };
        if constexpr (T::bytes == 2) {
            fma_sample(T::unscaled(raw.x), 0);
            fma_sample(T::unscaled_hi(raw.x), 1);
            fma_sample(T::unscaled(raw.y), 2);
            fma_sample(T::unscaled_hi(raw.y), 3);
        } else {
            fma_sample(T::unscaled(raw.x), 0);
            fma_sample(T::unscaled(raw.y), 1);
            fma_sample(T::unscaled(raw.z), 2);
            fma_sample(T::unscaled(raw.w), 3);
        }
    };
    Raw va[8], vb[8];
#pragma unroll
    for (int c = 0; c < 8; c++)
        if (c < nchan) va[c] = load(chan_a, c, ia), vb[c] = load(chan_b, c, ib);
#pragma unroll
    for (int c = 0; c < 8; c++)
        if (c < nchan) accumulate(acc_a, va[c], w_[c]), accumulate(acc_b, vb[c], w_[c]);
}

template <int FMT>
__device__ __forceinline__ void beam_quad(const BeamArgs &a, size_t i, float (&acc)[8]) {
    beam_quad<FMT>(a.chan, a.w, a.nchan, i, acc);
}

}  // namespace hz
