// elementwise.cu -- the HBM-bound kernels: convert (K1), lookup (K1L), shift (K2),
// rotate/scale/add (K3/K4/K5), decimate/downsample (K7), beamform (K8).
//
// All of these touch every byte once, so the design rules are the streaming ones: 16-byte
// coalesced stores, several independent loads in flight per thread, L1::no_allocate on data that
// is never re-read, grids sized as a multiple of the 148 SMs with a grid-stride loop.
#include "common.cuh"
#include "nco.cuh"
#include "batch_admit.h"
#include "nco_launch.h"
#include "beam.cuh"

namespace hz {

constexpr int kThreads = 256;
constexpr int kBlocksPerSM = 8;  // 2048 resident threads per SM

// =============================================================================================
// K1  integer -> complex64        (conv.go:55-93 -> iq_u8.go:111-121, iq_i8.go:107-119,
//                                  iq_i16.go:141-145).  Algorithmic HBM bytes: 2+8 (u8/i8),
//                                  4+8 (i16) per sample.
// Body: one thread converts a PAIR of samples per step (4 or 8 raw bytes in, one float4 out),
// UNROLL independent pairs in flight.  `head` (0/1) leading samples and an odd tail sample are
// converted by two designated threads so that sub-slices at any sample offset work
// (iq_u8_test.go:65-85).
// =============================================================================================
template <int FMT>
__device__ __forceinline__ float2 load_convert_one(const uint8_t *src, size_t j) {
    using T = RawTraits<FMT>;
    if constexpr (T::bytes == 2) {
        return T::conv((uint32_t) * reinterpret_cast<const uint16_t *>(src + 2 * j));
    } else {
        const uint16_t *p = reinterpret_cast<const uint16_t *>(src + 4 * j);
        return T::conv((uint32_t)p[0] | ((uint32_t)p[1] << 16));
    }
}

template <int FMT>
__device__ __forceinline__ float4 load_convert_pair(const uint8_t *src_body, size_t pair) {
    using T = RawTraits<FMT>;
    float2 a, b;
    if constexpr (T::bytes == 2) {
        const uint32_t w = ld_stream_u32(src_body + 4 * pair);
        a = T::conv(w);
        b = T::conv_hi(w);
    } else {
        const uint2 w = ld_stream_u64(src_body + 8 * pair);
        a = T::conv(w.x);
        b = T::conv(w.y);
    }
    return make_float4(a.x, a.y, b.x, b.y);
}

// the same pair, exact but not yet scaled (callers fold RawTraits::scale() into a later multiply)
template <int FMT>
__device__ __forceinline__ float4 load_unscaled_pair(const uint8_t *src_body, size_t pair) {
    using T = RawTraits<FMT>;
    float2 a, b;
    if constexpr (T::bytes == 2) {
        const uint32_t w = ld_stream_u32(src_body + 4 * pair);
        a = T::unscaled(w);
        b = T::unscaled_hi(w);
    } else {
        const uint2 w = ld_stream_u64(src_body + 8 * pair);
        a = T::unscaled(w.x);
        b = T::unscaled(w.y);
    }
    return make_float4(a.x, a.y, b.x, b.y);
}

template <int FMT, int UNROLL>
__global__ void __launch_bounds__(kThreads) k_convert(const uint8_t *__restrict__ src, float2 *__restrict__ dst,
                                                       size_t n, int head) {
    using T = RawTraits<FMT>;
    overlap_trigger();  // common.cuh, OverlapWindow: the next independent launch may start filling SMs
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t npairs = (n - head) / 2;
    const uint8_t *body = src + (size_t)head * T::bytes;
    float4 *out = reinterpret_cast<float4 *>(dst + head);

    // tile per CTA: UNROLL rows of kThreads consecutive pairs.  A CTA's reads and writes are one
    // contiguous span (HBM row locality); measured 6.4-6.5 TB/s for this 1:4 read:write mix against
    // 5.3 TB/s for a chip-wide grid-stride walk (tools/exp/exp_stream.cu).
    const size_t tile = (size_t)UNROLL * kThreads;
    for (size_t t0 = (size_t)blockIdx.x * tile; t0 < npairs; t0 += (size_t)gridDim.x * tile) {
        float4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const size_t i = t0 + (size_t)u * kThreads + threadIdx.x;
            if (i < npairs) v[u] = load_convert_pair<FMT>(body, i);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const size_t i = t0 + (size_t)u * kThreads + threadIdx.x;
            if (i < npairs) st_stream_f4(out + i, v[u]);
        }
    }

    if (tid == 0 && head) dst[0] = load_convert_one<FMT>(src, 0);
    if (tid == 1 && ((n - head) & 1)) dst[n - 1] = load_convert_one<FMT>(src, n - 1);
    overlap_join_all();
}

// fully general (any alignment): one sample per thread
template <int FMT>
__global__ void __launch_bounds__(kThreads) k_convert_scalar(const uint8_t *__restrict__ src, float2 *__restrict__ dst,
                                                              size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        dst[i] = load_convert_one<FMT>(src, i);
}

__global__ void __launch_bounds__(kThreads) k_i16_lsb_to_msb(uint16_t *buf, size_t ncomp, int shift) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < ncomp; i += stride)
        buf[i] = (uint16_t)(buf[i] << shift);
}

// K1L  dst[i] = table[src[i] as little-endian uint16]   (iq_lookup_table.go:198-251)
template <typename E>
__global__ void __launch_bounds__(kThreads) k_lookup(const uint16_t *__restrict__ src, const E *__restrict__ table,
                                                      E *__restrict__ dst, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = __ldg(table + src[i]);
}

// =============================================================================================
// K2  NCO mixer                    (stream/shifter.go:73-84).  In place: 8+8 B/sample; fused with
//                                  K1: 2+8 or 4+8 B/sample.
// One thread mixes a pair of samples per step.  The segment that contains the pair is cached in
// registers and re-looked-up only when the grid-stride walk leaves it.
// =============================================================================================
// Per-thread NCO state for a grid-stride walk over sample pairs.  Inside one accumulator segment
// the phase is linear in the sample index, so stepping the pair index by a fixed stride multiplies
// the rotation by a constant: rot(j + 2*stride) = rot(j) * E_stride, rot(j + 1) = rot(j) * E_1.  An
// exact polynomial sincos re-anchors the recurrence every kReanchor steps (error <= ~3 roundings)
// and whenever the walk crosses a segment boundary.
struct PairMixer {
    static constexpr int kReanchor = 4;
    NcoCursor cur;
    float2 rot = {1.f, 0.f}, e_step = {1.f, 0.f}, e_one = {1.f, 0.f};
    uint32_t next_j = 0xffffffffu;  // the pair index the recurrence can serve
    uint64_t seg_dp = ~0ull;
    int age = 0;
    float scale;
    uint32_t stride2;  // samples between consecutive pairs of this thread
    __device__ __forceinline__ PairMixer(float s, uint32_t pair_stride) : scale(s), stride2(2u * pair_stride) {}

    // rotations for samples j and j+1 (both scaled by `scale`)
    template <class Table>
    __device__ __forceinline__ void get(const Table &tab, uint32_t j, float2 &r0, float2 &r1) {
        const bool in_seg = j >= cur.j0 && j + 1 < cur.end;
        if (in_seg && j == next_j && age < kReanchor) {
            rot = cmul(rot, e_step);
            age++;
        } else {
            cur.seek(tab, j);
            if (cur.dp != seg_dp) {  // new segment: new step constants
                seg_dp = cur.dp;
                e_step = nco_rot((uint64_t)stride2 * cur.dp);
                e_one = nco_rot(cur.dp);
            }
            rot = nco_rot(cur.phase(j));
            rot.x *= scale;
            rot.y *= scale;
            age = 0;
        }
        next_j = j + stride2;
        r0 = rot;
        if (j + 1 < cur.end) {
            r1 = cmul(rot, e_one);
        } else {  // the pair straddles a segment boundary
            NcoCursor c2;
            c2.seek(tab, j + 1);
            r1 = nco_rot(c2.phase(j + 1));
            r1.x *= scale;
            r1.y *= scale;
            next_j = 0xffffffffu;
        }
    }
};

template <class Table>
__device__ __forceinline__ float2 mix_one(const Table &tab, uint32_t j, float2 v) {
    const int s = nco_find(tab, j);
    return cmul(v, nco_rot(nco_phase(tab.seg[s], j)));
}

// SRC_FMT == C64: in place (src ignored).  Otherwise fused convert+shift: the raw integers are
// converted exactly and the format's scale rides on the rotation.
// (the body is shared with the batched form below: `Table` is the launch's NcoTable or a buffer's view into a BatchTable)
template <int SRC_FMT, int UNROLL, class Table>
__device__ __forceinline__ void shift_body(const uint8_t *__restrict__ src, float2 *dst, uint32_t n, int head, const Table &tab) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t npairs = (n - head) / 2;
    float4 *out = reinterpret_cast<float4 *>(dst + head);
    float scale = 1.0f;
    if constexpr (SRC_FMT != HZSDR_FORMAT_C64) scale = RawTraits<SRC_FMT>::scale();
    // consecutive pairs of a thread: one row (kThreads pairs) apart in tile form, one grid apart otherwise
    PairMixer mix(scale, SRC_FMT == HZSDR_FORMAT_C64 ? (uint32_t)kThreads : gridDim.x * blockDim.x);

    auto load_pair = [&](uint32_t p) -> float4 {
        if constexpr (SRC_FMT == HZSDR_FORMAT_C64) {
            return ld_inplace_f4(out + p);
        } else {
            return load_unscaled_pair<SRC_FMT>(src + (size_t)head * RawTraits<SRC_FMT>::bytes, p);
        }
    };
    auto load_one = [&](uint32_t j) -> float2 {
        if constexpr (SRC_FMT == HZSDR_FORMAT_C64) {
            return dst[j];
        } else {
            return load_convert_one<SRC_FMT>(src, j);
        }
    };
    auto mix_pair = [&](uint32_t p, float4 v) -> float4 {
        float2 r0, r1;
        mix.get(tab, head + 2 * p, r0, r1);
        const float2 a = cmul(make_float2(v.x, v.y), r0);
        const float2 b = cmul(make_float2(v.z, v.w), r1);
        return make_float4(a.x, a.y, b.x, b.y);
    };

    if constexpr (SRC_FMT == HZSDR_FORMAT_C64) {
        // in place (8 B read + 8 B written per sample): one tile of UNROLL rows per CTA.  A CTA's
        // traffic is one contiguous span (HBM row locality): 103% of the copy peak, against 93% for
        // a chip-wide grid-stride walk.
        const uint32_t tile = UNROLL * kThreads;
        for (uint64_t t0 = (uint64_t)blockIdx.x * tile; t0 < npairs; t0 += (uint64_t)gridDim.x * tile) {
            float4 v[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                const uint32_t i = (uint32_t)t0 + u * kThreads + threadIdx.x;
                if (i < npairs) v[u] = load_pair(i);
            }
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                const uint32_t i = (uint32_t)t0 + u * kThreads + threadIdx.x;
                if (i < npairs) st_stream_f4(out + i, mix_pair(i, v[u]));
            }
        }
    } else {
        // fused convert+shift (2-4 B read + 8 B written): persistent grid-stride walk.  The reads are
        // too small to keep HBM busy from short-lived CTAs (tile form: 53-60%); a long-lived thread
        // amortises the NCO set-up over ~50 pairs and reaches 79-83%.
        const uint32_t stride = gridDim.x * blockDim.x;
        uint32_t i = tid;
        for (; (uint64_t)i + (uint64_t)(UNROLL - 1) * stride < npairs; i += UNROLL * stride) {
            float4 v[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; u++) v[u] = load_pair(i + u * stride);
#pragma unroll
            for (int u = 0; u < UNROLL; u++) st_stream_f4(out + i + u * stride, mix_pair(i + u * stride, v[u]));
        }
        for (; i < npairs; i += stride) st_stream_f4(out + i, mix_pair(i, load_pair(i)));
    }

    if (tid == 0 && head) dst[0] = mix_one(tab, 0, load_one(0));
    if (tid == 1 && ((n - head) & 1)) dst[n - 1] = mix_one(tab, n - 1, load_one(n - 1));
}

template <int SRC_FMT, int UNROLL>
__global__ void __launch_bounds__(kThreads) k_shift(const uint8_t *__restrict__ src, float2 *dst, uint32_t n, int head,
                                                     const __grid_constant__ NcoTable tab) {
    overlap_trigger();
    shift_body<SRC_FMT, UNROLL>(src, dst, n, head, tab);
    overlap_join_all();
}

// K consecutive buffers of one stream in ONE launch (hzsdr_convert_shift_batch): blockIdx.y = buffer, its source,
// destination and accumulator segments from the BatchTable in the kernel parameters (nco.cuh).  A 2^20-sample
// buffer is 10 MB of traffic = 1.6 us of HBM time: one launch per buffer is launch-bound at 35% of the roofline.
template <int SRC_FMT, int UNROLL>
__global__ void __launch_bounds__(kThreads) k_shift_batch(uint32_t n, const __grid_constant__ BatchTable tbl) {
    overlap_trigger();
    const StreamDesc &d = tbl.desc[blockIdx.y];
    shift_body<SRC_FMT, UNROLL>(d.src, reinterpret_cast<float2 *>(d.dst), n, 0, SegView{tbl.seg + d.seg_off, d.count});
    overlap_join_all();
}

// =============================================================================================
// K3 rotate / K4 scale             (internal/simd/mult.go:25-33, mult_simd_amd64.s:47-54)
// In place, 16 B/sample.  Rotate widens to fp64 like the Go compiler -> bit-equal results.
// =============================================================================================
template <bool ROTATE, int UNROLL>
__global__ void __launch_bounds__(kThreads) k_rotate_scale(float2 *buf, size_t n, int head, float mr, float mi) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t npairs = (n - head) / 2;
    float4 *body = reinterpret_cast<float4 *>(buf + head);
    const float2 m = make_float2(mr, mi);
    auto op1 = [&](float2 a) -> float2 {
        if constexpr (ROTATE) return go_cmul(a, m);
        return make_float2(a.x * mr, a.y * mr);
    };
    auto op2 = [&](float4 v) -> float4 {
        float2 a = op1(make_float2(v.x, v.y)), b = op1(make_float2(v.z, v.w));
        return make_float4(a.x, a.y, b.x, b.y);
    };
    const size_t tile = (size_t)UNROLL * kThreads;
    for (size_t t0 = (size_t)blockIdx.x * tile; t0 < npairs; t0 += (size_t)gridDim.x * tile) {
        float4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const size_t i = t0 + (size_t)u * kThreads + threadIdx.x;
            if (i < npairs) v[u] = ld_inplace_f4(body + i);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const size_t i = t0 + (size_t)u * kThreads + threadIdx.x;
            if (i < npairs) st_stream_f4(body + i, op2(v[u]));
        }
    }
    if (tid == 0 && head) buf[0] = op1(buf[0]);
    if (tid == 1 && ((n - head) & 1)) buf[n - 1] = op1(buf[n - 1]);
}

// =============================================================================================
// K5  ordered K-way add            (stream/add.go:115-119,165-168): out = ((0+s0)+s1)+...
// 8K+8 B/sample.  Works on fp32 components (a complex add is two independent fp32 adds), float4
// wide when everything is 16-byte aligned.
// =============================================================================================
constexpr int kMaxAddSrcs = 64;
struct AddSrcs {
    const float *p[kMaxAddSrcs];
};

template <typename V>
__global__ void __launch_bounds__(kThreads) k_add(V *__restrict__ dst, const __grid_constant__ AddSrcs srcs, int k,
                                                   size_t nvec, bool accumulate) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        V acc;
        if (accumulate) {
            acc = dst[i];
        } else {
            if constexpr (sizeof(V) == 16)
                acc = V{0.f, 0.f, 0.f, 0.f};
            else
                acc = V{0.f};
        }
        int c = 0;
        for (; c + 4 <= k; c += 4) {
            V v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) v[u] = reinterpret_cast<const V *>(srcs.p[c + u])[i];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if constexpr (sizeof(V) == 16) {
                    acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w;
                } else {
                    acc.x += v[u].x;
                }
            }
        }
        for (; c < k; c++) {
            V v = reinterpret_cast<const V *>(srcs.p[c])[i];
            if constexpr (sizeof(V) == 16) {
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            } else {
                acc.x += v.x;
            }
        }
        dst[i] = acc;
    }
}

// =============================================================================================
// K7  decimate / downsample        (stream/decimate.go:84-98, stream/downsample.go:97-124)
// =============================================================================================
template <typename E>
__global__ void __launch_bounds__(kThreads) k_decimate(const E *__restrict__ src, E *__restrict__ dst, size_t nblocks,
                                                        size_t block, uint32_t per_block, uint32_t factor) {
    const size_t total = nblocks * per_block;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += stride) {
        const size_t q = o / per_block;
        const size_t i = o - q * per_block;
        dst[o] = src[q * block + i * factor];
    }
}

template <int FMT>
__global__ void __launch_bounds__(kThreads) k_downsample(const uint8_t *__restrict__ src, float2 *__restrict__ dst,
                                                          size_t nblocks, size_t block, uint32_t per_block,
                                                          uint32_t factor) {
    const size_t total = nblocks * per_block;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const float ff = (float)factor;
    for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += stride) {
        const size_t q = o / per_block;
        const size_t i = o - q * per_block;
        const size_t base = q * block + i * factor;
        float sr = 0.f, si = 0.f;  // sequential complex64 sum, downsample.go:115-117
        for (uint32_t j = 0; j < factor; j++) {
            float2 v;
            if constexpr (FMT == HZSDR_FORMAT_C64)
                v = reinterpret_cast<const float2 *>(src)[base + j];
            else
                v = load_convert_one<FMT>(src, base + j);
            sr += v.x;
            si += v.y;
        }
        dst[o] = make_float2(__fdiv_rn(sr, ff), __fdiv_rn(si, ff));
    }
}

// =============================================================================================
// K8  beamform                     (stream/beamform.go:148-171 -> multiply.go:46-70, add.go:115-185)
// dst[n] = ((0 + w0*x0[n]) + w1*x1[n]) + ...   One pass: nchan*raw + 8 B per output sample.
// A thread owns a pair of output samples; channels are walked in order, 8 loads in flight.
// =============================================================================================
// K8 single GPU: a thread owns FOUR consecutive output samples (beam.cuh)
template <int FMT>
__global__ void __launch_bounds__(kThreads) k_beamform(float4 *__restrict__ dst, size_t nquads,
                                                        const __grid_constant__ BeamArgs a) {
    overlap_trigger();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nquads; i += stride) {
        float acc[8];
        if (a.accumulate) {
            const float4 p = dst[2 * i], q = dst[2 * i + 1];
            acc[0] = p.x; acc[1] = p.y; acc[2] = p.z; acc[3] = p.w; acc[4] = q.x; acc[5] = q.y; acc[6] = q.z; acc[7] = q.w;
        } else {
#pragma unroll
            for (int k = 0; k < 8; k++) acc[k] = 0.f;
        }
        beam_quad<FMT>(a, i, acc);
        st_stream_f4(dst + 2 * i, make_float4(acc[0], acc[1], acc[2], acc[3]));
        st_stream_f4(dst + 2 * i + 1, make_float4(acc[4], acc[5], acc[6], acc[7]));
    }
    overlap_join_all();
}

}  // namespace hz

using namespace hz;

// launch of an overlappable kernel (one that brackets its body with overlap_trigger / overlap_join_all):
// `r` / `w` are the byte spans it reads / writes
template <class... KArgs, class... Args>
static int launch_overlapping(hzsdr_ctx *ctx, void (*kernel)(KArgs...), int grid, OverlapWindow::Span r,
                              OverlapWindow::Span w, Args &&...args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    overlap_launch_config(cfg, attr, ctx->overlap.admit(r, w, ctx->overlap_pred_ok()));
    ctx->overlap_launched();
    HZ_CUDA(cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...));
    return HZSDR_OK;
}

// =============================================================================================
// C ABI
// =============================================================================================
// one CTA per tile of rows*kThreads work items (non-persistent; the kernels loop if the cap bites)
static inline int tile_grid(size_t items, int rows) {
    const size_t tile = (size_t)rows * kThreads;
    size_t g = (items + tile - 1) / tile;
    if (g < 1) g = 1;
    return (int)(g > (1u << 22) ? (1u << 22) : g);
}

static inline bool aligned(const void *p, size_t a) { return ((uintptr_t)p & (a - 1)) == 0; }

// returns head (0/1) if the pair-vectorised body can be used, -1 otherwise
static int vector_head(const void *src, int src_bytes, const void *dst) {
    for (int head = 0; head < 2; head++) {
        const bool s_ok = src == nullptr || aligned((const uint8_t *)src + (size_t)head * src_bytes, 2 * src_bytes);
        const bool d_ok = aligned((const uint8_t *)dst + (size_t)head * 8, 16);
        if (s_ok && d_ok) return head;
    }
    return -1;
}

extern "C" int hzsdr_convert_to_c64(hzsdr_ctx *ctx, int src_format, const void *src, size_t src_len, void *dst,
                                    size_t dst_len, size_t *n_out) {
    HZ_ENTER(ctx);
    if (n_out) *n_out = 0;
    if (src_format == HZSDR_FORMAT_C64) {  // conv.go:56-58 -> CopySamples copies min(len)
        const size_t n = src_len < dst_len ? src_len : dst_len;
        if (n) HZ_CUDA(cudaMemcpyAsync(dst, src, n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        if (n_out) *n_out = n;
        return HZSDR_OK;
    }
    if (src_format != HZSDR_FORMAT_U8 && src_format != HZSDR_FORMAT_I8 && src_format != HZSDR_FORMAT_I16)
        return fail(HZSDR_ERR_FORMAT_UNKNOWN, "hzsdr_convert_to_c64: unknown source format %d", src_format);
    if (src_len > dst_len) return fail(HZSDR_ERR_DST_TOO_SMALL, "hzsdr_convert_to_c64: %zu > %zu", src_len, dst_len);
    const size_t n = src_len;
    if (n == 0) return HZSDR_OK;
    if (!src || !dst) return fail(HZSDR_ERR_INVALID, "hzsdr_convert_to_c64: null buffer");
    const int sb = hzsdr_format_size(src_format);
    if (!aligned(src, sb) || !aligned(dst, 8))
        return fail(HZSDR_ERR_INVALID, "hzsdr_convert_to_c64: buffers must be sample-aligned");
    const int head = vector_head(src, sb, dst);
    const uint8_t *s = (const uint8_t *)src;
    float2 *d = (float2 *)dst;
    if (head >= 0 && n >= 2) {
        const int grid = tile_grid((n + 1) / 2, 4);
        const OverlapWindow::Span r = OverlapWindow::span(s, n * sb), w = OverlapWindow::span(d, n * 8);
        int rc;
        switch (src_format) {
            case HZSDR_FORMAT_U8: rc = launch_overlapping(ctx, k_convert<HZSDR_FORMAT_U8, 4>, grid, r, w, s, d, n, head); break;
            case HZSDR_FORMAT_I8: rc = launch_overlapping(ctx, k_convert<HZSDR_FORMAT_I8, 4>, grid, r, w, s, d, n, head); break;
            default: rc = launch_overlapping(ctx, k_convert<HZSDR_FORMAT_I16, 4>, grid, r, w, s, d, n, head); break;
        }
        if (rc) return rc;
    } else {
        const int grid = stream_grid(ctx, n, kThreads, kBlocksPerSM);
        switch (src_format) {
            case HZSDR_FORMAT_U8: k_convert_scalar<HZSDR_FORMAT_U8><<<grid, kThreads, 0, ctx->stream>>>(s, d, n); break;
            case HZSDR_FORMAT_I8: k_convert_scalar<HZSDR_FORMAT_I8><<<grid, kThreads, 0, ctx->stream>>>(s, d, n); break;
            default: k_convert_scalar<HZSDR_FORMAT_I16><<<grid, kThreads, 0, ctx->stream>>>(s, d, n); break;
        }
    }
    HZ_CHECK_LAUNCH();
    if (n_out) *n_out = n;
    return HZSDR_OK;
}

extern "C" int hzsdr_i16_shift_lsb_to_msb(hzsdr_ctx *ctx, void *buf, size_t n, int bits) {
    HZ_ENTER(ctx);
    if (bits < 1 || bits > 16) return fail(HZSDR_ERR_INVALID, "hzsdr_i16_shift_lsb_to_msb: bits=%d", bits);
    if (n == 0 || bits == 16) return HZSDR_OK;
    const int grid = stream_grid(ctx, 2 * n, kThreads, kBlocksPerSM);
    k_i16_lsb_to_msb<<<grid, kThreads, 0, ctx->stream>>>((uint16_t *)buf, 2 * n, 16 - bits);
    HZ_CHECK_LAUNCH();
    return HZSDR_OK;
}

extern "C" int hzsdr_lookup(hzsdr_ctx *ctx, int src_format, const void *src, size_t n, int table_format,
                            const void *table, void *dst, size_t dst_len) {
    HZ_ENTER(ctx);
    if (src_format != HZSDR_FORMAT_U8 && src_format != HZSDR_FORMAT_I8)
        return fail(HZSDR_ERR_FORMAT_UNKNOWN, "hzsdr_lookup: source must be U8 or I8");  // iq_lookup_table.go:111-116
    if (dst_len < n) return fail(HZSDR_ERR_DST_TOO_SMALL, "hzsdr_lookup: %zu < %zu", dst_len, n);
    if (n == 0) return HZSDR_OK;
    const int grid = stream_grid(ctx, n, kThreads, kBlocksPerSM);
    const uint16_t *s = (const uint16_t *)src;
    switch (hzsdr_format_size(table_format)) {
        case 2: k_lookup<uint16_t><<<grid, kThreads, 0, ctx->stream>>>(s, (const uint16_t *)table, (uint16_t *)dst, n); break;
        case 4: k_lookup<uint32_t><<<grid, kThreads, 0, ctx->stream>>>(s, (const uint32_t *)table, (uint32_t *)dst, n); break;
        case 8: k_lookup<uint2><<<grid, kThreads, 0, ctx->stream>>>(s, (const uint2 *)table, (uint2 *)dst, n); break;
        default: return fail(HZSDR_ERR_FORMAT_UNKNOWN, "hzsdr_lookup: unknown table format %d", table_format);
    }
    HZ_CHECK_LAUNCH();
    return HZSDR_OK;
}

// ---- shift -----------------------------------------------------------------------------------
namespace hz {
// Launch `launch(table, first_sample, count)` for consecutive sub-ranges of [0, n), each covered
// by at most kMaxSegsPerLaunch segments.  Steady state: one launch.
template <class Launch>
static int for_each_nco_launch(size_t n, double freq_hz, hzsdr_nco *state, Launch &&launch) {
    std::vector<HostSeg> segs;
    double ts = state->ts;
    build_segments(state->sample_rate, n, &ts, segs);
    std::vector<NcoLaunch> launches;
    int rc = plan_nco_launches(segs, n, 2, freq_hz, launches);
    if (rc) return rc;
    for (const NcoLaunch &L : launches) {
        rc = launch(L.table, L.first, L.count);
        if (rc != HZSDR_OK) return rc;
    }
    state->ts = ts;
    return HZSDR_OK;
}
}  // namespace hz

namespace hz {
template <int FMT>
__global__ void __launch_bounds__(kThreads) k_shift_scalar(const uint8_t *__restrict__ src, float2 *dst, uint32_t n,
                                                            const __grid_constant__ NcoTable tab) {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        float2 v;
        if constexpr (FMT == HZSDR_FORMAT_C64)
            v = dst[j];
        else
            v = load_convert_one<FMT>(src, j);
        dst[j] = mix_one(tab, j, v);
    }
}
}  // namespace hz

template <int FMT>
static int shift_impl(hzsdr_ctx *ctx, const void *src, void *dst, size_t n, double freq_hz, hzsdr_nco *state) {
    constexpr int sb = FMT == HZSDR_FORMAT_C64 ? 8 : RawTraits<FMT == HZSDR_FORMAT_C64 ? HZSDR_FORMAT_U8 : FMT>::bytes;
    return for_each_nco_launch(n, freq_hz, state, [&](const NcoTable &tab, size_t first, size_t count) -> int {
        const uint8_t *s = FMT == HZSDR_FORMAT_C64 ? nullptr : (const uint8_t *)src + first * sb;
        float2 *d = (float2 *)dst + first;
        const int head = vector_head(s, FMT == HZSDR_FORMAT_C64 ? 0 : sb, d);
        if (head >= 0 && count >= 2) {
            const int grid = FMT == HZSDR_FORMAT_C64 ? tile_grid((count + 1) / 2, 4) : stream_grid(ctx, (count + 1) / 2, kThreads, kBlocksPerSM);
            const OverlapWindow::Span w = OverlapWindow::span(d, count * 8);
            const OverlapWindow::Span r = FMT == HZSDR_FORMAT_C64 ? w : OverlapWindow::span(s, count * sb);
            return launch_overlapping(ctx, k_shift<FMT, 4>, grid, r, w, s, d, (uint32_t)count, head, tab);
        } else {
            const int grid = stream_grid(ctx, count, kThreads, kBlocksPerSM);
            k_shift_scalar<FMT><<<grid, kThreads, 0, ctx->stream>>>(s, d, (uint32_t)count, tab);
        }
        HZ_CHECK_LAUNCH();
        return HZSDR_OK;
    });
}

extern "C" int hzsdr_shift(hzsdr_ctx *ctx, void *buf, size_t n, double freq_hz, hzsdr_nco *state) {
    HZ_ENTER(ctx);
    if (!state || state->sample_rate == 0) return fail(HZSDR_ERR_INVALID, "hzsdr_shift: bad NCO state");
    if (n == 0) return HZSDR_OK;
    if (!buf || !aligned(buf, 8)) return fail(HZSDR_ERR_INVALID, "hzsdr_shift: bad buffer");
    return shift_impl<HZSDR_FORMAT_C64>(ctx, nullptr, buf, n, freq_hz, state);
}

extern "C" int hzsdr_convert_shift(hzsdr_ctx *ctx, int src_format, const void *src, size_t n, void *dst,
                                   size_t dst_len, double freq_hz, hzsdr_nco *state) {
    HZ_ENTER(ctx);
    if (!state || state->sample_rate == 0) return fail(HZSDR_ERR_INVALID, "hzsdr_convert_shift: bad NCO state");
    if (n > dst_len) return fail(HZSDR_ERR_DST_TOO_SMALL, "hzsdr_convert_shift: %zu > %zu", n, dst_len);
    if (n == 0) return HZSDR_OK;
    if (!src || !dst || !aligned(dst, 8) || !aligned(src, hzsdr_format_size(src_format) ? hzsdr_format_size(src_format) : 1))
        return fail(HZSDR_ERR_INVALID, "hzsdr_convert_shift: bad buffer");
    switch (src_format) {
        case HZSDR_FORMAT_U8: return shift_impl<HZSDR_FORMAT_U8>(ctx, src, dst, n, freq_hz, state);
        case HZSDR_FORMAT_I8: return shift_impl<HZSDR_FORMAT_I8>(ctx, src, dst, n, freq_hz, state);
        case HZSDR_FORMAT_I16: return shift_impl<HZSDR_FORMAT_I16>(ctx, src, dst, n, freq_hz, state);
        case HZSDR_FORMAT_C64:
            if (src != dst) HZ_CUDA(cudaMemcpyAsync(dst, src, n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
            return shift_impl<HZSDR_FORMAT_C64>(ctx, nullptr, dst, n, freq_hz, state);
        default: return fail(HZSDR_ERR_FORMAT_UNKNOWN, "hzsdr_convert_shift: unknown format %d", src_format);
    }
}

template <int FMT>
static int shift_batch_launch(hzsdr_ctx *ctx, const BatchTable &tbl, uint32_t nb, size_t n, bool may) {
    // about kBlocksPerSM CTAs per SM over the whole batch, every thread a long walk inside its buffer
    size_t per = ((size_t)ctx->sm_count * kBlocksPerSM + nb - 1) / nb;
    const size_t need = ((n + 1) / 2 + kThreads - 1) / kThreads;
    if (per > need) per = need;
    if (per < 1) per = 1;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)per, nb);
    cfg.blockDim = dim3(kThreads);
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    overlap_launch_config(cfg, attr, may);
    HZ_CUDA(cudaLaunchKernelEx(&cfg, k_shift_batch<FMT, 4>, (uint32_t)n, tbl));
    return HZSDR_OK;
}

// `count` consecutive buffers of one stream (n_each samples each) through the fused Convert + Shift: one launch per
// <= 64 buffers, descriptors and accumulator segments in the kernel parameters, spans hazard-checked like single
// launches.  Buffers that are not 16-byte aligned or need more segments than the table holds go one by one.
extern "C" int hzsdr_convert_shift_batch(hzsdr_ctx *ctx, int src_format, const void *const *srcs, size_t n_each, void *const *dsts,
                                         size_t dst_len_each, size_t count, double freq_hz, hzsdr_nco *state) {
    HZ_ENTER(ctx);
    if (!state || state->sample_rate == 0) return fail(HZSDR_ERR_INVALID, "hzsdr_convert_shift_batch: bad NCO state");
    if (count && (!srcs || !dsts)) return fail(HZSDR_ERR_INVALID, "hzsdr_convert_shift_batch: null buffer table");
    if (n_each > dst_len_each) return fail(HZSDR_ERR_DST_TOO_SMALL, "hzsdr_convert_shift_batch: %zu > %zu", n_each, dst_len_each);
    if (n_each == 0 || count == 0) return HZSDR_OK;
    const int sb = hzsdr_format_size(src_format);
    const bool raw = src_format == HZSDR_FORMAT_U8 || src_format == HZSDR_FORMAT_I8 || src_format == HZSDR_FORMAT_I16;
    if (!raw || n_each < 2 || n_each > 0x7fffffffull) {
        for (size_t k = 0; k < count; k++) {
            int rc = hzsdr_convert_shift(ctx, src_format, srcs[k], n_each, dsts[k], dst_len_each, freq_hz, state);
            if (rc) return rc;
        }
        return HZSDR_OK;
    }
    {   // inside one kernel nothing is ordered: a call whose buffers overlap each other keeps its order one buffer at a time
        std::vector<BufSpan> all;
        all.reserve(2 * count);
        for (size_t k = 0; k < count; k++) {
            if (!srcs[k] || !dsts[k]) return fail(HZSDR_ERR_INVALID, "hzsdr_convert_shift_batch: null buffer %zu", k);
            all.push_back({(uintptr_t)srcs[k], (uintptr_t)srcs[k] + n_each * (size_t)sb, false});
            all.push_back({(uintptr_t)dsts[k], (uintptr_t)dsts[k] + n_each * 8, true});
        }
        if (write_conflict(std::move(all))) {
            for (size_t k = 0; k < count; k++) {
                int rc = hzsdr_convert_shift(ctx, src_format, srcs[k], n_each, dsts[k], dst_len_each, freq_hz, state);
                if (rc) return rc;
            }
            return HZSDR_OK;
        }
    }
    BatchTable tbl;
    uint32_t nb = 0, ns = 0;
    std::vector<BufSpan> pending;
    auto flush = [&]() -> int {
        if (!nb) return HZSDR_OK;
        const bool may = admit_spans(ctx, std::move(pending));
        pending.clear();
        int rc = HZSDR_OK;
        switch (src_format) {
            case HZSDR_FORMAT_U8: rc = shift_batch_launch<HZSDR_FORMAT_U8>(ctx, tbl, nb, n_each, may); break;
            case HZSDR_FORMAT_I8: rc = shift_batch_launch<HZSDR_FORMAT_I8>(ctx, tbl, nb, n_each, may); break;
            default: rc = shift_batch_launch<HZSDR_FORMAT_I16>(ctx, tbl, nb, n_each, may); break;
        }
        nb = ns = 0;
        return rc;
    };
    std::vector<HostSeg> segs;
    std::vector<NcoLaunch> launches;
    for (size_t k = 0; k < count; k++) {
        if (!srcs[k] || !dsts[k]) return fail(HZSDR_ERR_INVALID, "hzsdr_convert_shift_batch: null buffer %zu", k);
        double ts = state->ts;
        build_segments(state->sample_rate, n_each, &ts, segs);
        int rc = plan_nco_launches(segs, n_each, 2, freq_hz, launches);
        if (rc) return rc;
        const bool fits = launches.size() == 1 && launches[0].table.count <= kParamSegs && aligned(srcs[k], 16) && aligned(dsts[k], 16);
        if (!fits) {  // in order; advances state->ts itself
            rc = flush();
            if (rc) return rc;
            rc = hzsdr_convert_shift(ctx, src_format, srcs[k], n_each, dsts[k], dst_len_each, freq_hz, state);
            if (rc) return rc;
            continue;
        }
        const NcoTable &t = launches[0].table;
        if (nb == (uint32_t)kParamStreams || ns + (uint32_t)t.count > (uint32_t)kParamSegs) {
            rc = flush();
            if (rc) return rc;
        }
        pending.push_back({(uintptr_t)srcs[k], (uintptr_t)srcs[k] + n_each * (size_t)sb, false});
        pending.push_back({(uintptr_t)dsts[k], (uintptr_t)dsts[k] + n_each * 8, true});
        StreamDesc &d = tbl.desc[nb];
        d.src = (const uint8_t *)srcs[k];
        d.dst = dsts[k];
        d.seg_off = ns;
        d.count = t.count;
        d.dp_nom = 0;
        for (int q = 0; q < t.count; q++) tbl.seg[ns + q] = t.seg[q];
        nb++;
        ns += (uint32_t)t.count;
        state->ts = ts;
    }
    return flush();
}

// ---- rotate / scale / add --------------------------------------------------------------------
template <bool ROTATE>
static int rotate_scale(hzsdr_ctx *ctx, void *buf, size_t n, float a, float b) {
    if (n == 0) return HZSDR_OK;
    if (!buf || !aligned(buf, 8)) return fail(HZSDR_ERR_INVALID, "rotate/scale: bad buffer");
    const int head = aligned(buf, 16) ? 0 : 1;
    const int grid = tile_grid((n + 1) / 2, 4);
    k_rotate_scale<ROTATE, 4><<<grid, kThreads, 0, ctx->stream>>>((float2 *)buf, n, head, a, b);
    HZ_CHECK_LAUNCH();
    return HZSDR_OK;
}

extern "C" int hzsdr_rotate(hzsdr_ctx *ctx, void *buf, size_t n, float m_re, float m_im) {
    HZ_ENTER(ctx);
    return rotate_scale<true>(ctx, buf, n, m_re, m_im);
}

extern "C" int hzsdr_scale(hzsdr_ctx *ctx, void *buf, size_t n, float r) {
    HZ_ENTER(ctx);
    return rotate_scale<false>(ctx, buf, n, r, 0.f);
}

extern "C" int hzsdr_add(hzsdr_ctx *ctx, void *dst, const void *const *srcs, int k, size_t n) {
    HZ_ENTER(ctx);
    if (k < 1 || !srcs) return fail(HZSDR_ERR_INVALID, "hzsdr_add: no sources");  // stream/add.go:44-45
    if (n == 0) return HZSDR_OK;
    bool vec = aligned(dst, 16) && (n % 2 == 0);
    for (int c = 0; c < k; c++) {
        if (!srcs[c] || !aligned(srcs[c], 8)) return fail(HZSDR_ERR_INVALID, "hzsdr_add: bad source %d", c);
        vec = vec && aligned(srcs[c], 16);
    }
    for (int c0 = 0; c0 < k; c0 += kMaxAddSrcs) {
        AddSrcs a;
        const int kk = (k - c0) < kMaxAddSrcs ? (k - c0) : kMaxAddSrcs;
        for (int c = 0; c < kk; c++) a.p[c] = (const float *)srcs[c0 + c];
        if (vec) {
            const int grid = tile_grid(n / 2, 1);
            k_add<float4><<<grid, kThreads, 0, ctx->stream>>>((float4 *)dst, a, kk, n / 2, c0 > 0);
        } else {
            const int grid = stream_grid(ctx, 2 * n, kThreads, kBlocksPerSM);
            k_add<float1><<<grid, kThreads, 0, ctx->stream>>>((float1 *)dst, a, kk, 2 * n, c0 > 0);
        }
        HZ_CHECK_LAUNCH();
    }
    return HZSDR_OK;
}

// ---- decimate / downsample -------------------------------------------------------------------
extern "C" int hzsdr_decimate(hzsdr_ctx *ctx, int format, const void *src, size_t n, void *dst, size_t dst_len,
                              unsigned factor, size_t block, size_t *n_out) {
    HZ_ENTER(ctx);
    if (n_out) *n_out = 0;
    if (factor == 0) return fail(HZSDR_ERR_INVALID, "hzsdr_decimate: factor 0");
    if (format != HZSDR_FORMAT_U8 && format != HZSDR_FORMAT_I16 && format != HZSDR_FORMAT_C64)
        return fail(HZSDR_ERR_FORMAT_UNKNOWN, "hzsdr_decimate: format %d not supported (stream/decimate.go:85-97)", format);
    const size_t blk = block ? block : n;
    const size_t nblocks = blk ? n / blk : 0;
    const size_t per = blk / factor;
    const size_t total = nblocks * per;
    if (dst_len < total) return fail(HZSDR_ERR_DST_TOO_SMALL, "hzsdr_decimate: %zu < %zu", dst_len, total);
    if (total == 0) return HZSDR_OK;
    const int grid = stream_grid(ctx, total, kThreads, kBlocksPerSM);
    switch (format) {
        case HZSDR_FORMAT_U8: k_decimate<uint16_t><<<grid, kThreads, 0, ctx->stream>>>((const uint16_t *)src, (uint16_t *)dst, nblocks, blk, (uint32_t)per, factor); break;
        case HZSDR_FORMAT_I16: k_decimate<uint32_t><<<grid, kThreads, 0, ctx->stream>>>((const uint32_t *)src, (uint32_t *)dst, nblocks, blk, (uint32_t)per, factor); break;
        default: k_decimate<float2><<<grid, kThreads, 0, ctx->stream>>>((const float2 *)src, (float2 *)dst, nblocks, blk, (uint32_t)per, factor); break;
    }
    HZ_CHECK_LAUNCH();
    if (n_out) *n_out = total;
    return HZSDR_OK;
}

extern "C" int hzsdr_downsample(hzsdr_ctx *ctx, int src_format, const void *src, size_t n, void *dst, size_t dst_len,
                                unsigned factor, size_t block, size_t *n_out) {
    HZ_ENTER(ctx);
    if (n_out) *n_out = 0;
    if (factor == 0) return fail(HZSDR_ERR_INVALID, "hzsdr_downsample: factor 0");
    if (src_format != HZSDR_FORMAT_U8 && src_format != HZSDR_FORMAT_I16 && src_format != HZSDR_FORMAT_C64)
        return fail(HZSDR_ERR_FORMAT_UNKNOWN, "hzsdr_downsample: format %d not supported (stream/downsample.go:104-113)", src_format);
    const size_t blk = block ? block : n;
    const size_t nblocks = blk ? n / blk : 0;
    const size_t per = blk / factor;
    const size_t total = nblocks * per;
    if (dst_len < total) return fail(HZSDR_ERR_DST_TOO_SMALL, "hzsdr_downsample: %zu < %zu", dst_len, total);
    if (total == 0) return HZSDR_OK;
    const int grid = stream_grid(ctx, total, kThreads, kBlocksPerSM);
    const uint8_t *s = (const uint8_t *)src;
    switch (src_format) {
        case HZSDR_FORMAT_U8: k_downsample<HZSDR_FORMAT_U8><<<grid, kThreads, 0, ctx->stream>>>(s, (float2 *)dst, nblocks, blk, (uint32_t)per, factor); break;
        case HZSDR_FORMAT_I16: k_downsample<HZSDR_FORMAT_I16><<<grid, kThreads, 0, ctx->stream>>>(s, (float2 *)dst, nblocks, blk, (uint32_t)per, factor); break;
        default: k_downsample<HZSDR_FORMAT_C64><<<grid, kThreads, 0, ctx->stream>>>(s, (float2 *)dst, nblocks, blk, (uint32_t)per, factor); break;
    }
    HZ_CHECK_LAUNCH();
    if (n_out) *n_out = total;
    return HZSDR_OK;
}

// ---- beamform --------------------------------------------------------------------------------
extern "C" int hzsdr_beamform(hzsdr_ctx *ctx, int src_format, const void *const *chans, int nchan,
                              const float *weights, size_t n, void *dst) {
    HZ_ENTER(ctx);
    if (nchan < 1 || !chans || !weights) return fail(HZSDR_ERR_INVALID, "hzsdr_beamform: no channels");
    if (src_format != HZSDR_FORMAT_U8 && src_format != HZSDR_FORMAT_I8 && src_format != HZSDR_FORMAT_I16)
        return fail(HZSDR_ERR_FORMAT_UNKNOWN, "hzsdr_beamform: raw source format expected, got %d", src_format);
    if (n == 0) return HZSDR_OK;
    if (n % 4 || !aligned(dst, 16)) return fail(HZSDR_ERR_INVALID, "hzsdr_beamform: n must be a multiple of 4 and dst 16-byte aligned");
    const int sb = hzsdr_format_size(src_format);
    for (int c = 0; c < nchan; c++)
        if (!chans[c] || !aligned(chans[c], 4 * sb)) return fail(HZSDR_ERR_INVALID, "hzsdr_beamform: channel %d misaligned", c);
    const int grid = tile_grid(n / 4, 1);
    const float wscale = beam_weight_scale(src_format);
    for (int c0 = 0; c0 < nchan; c0 += kMaxBeamChans) {
        BeamArgs a;
        a.nchan = (nchan - c0) < kMaxBeamChans ? (nchan - c0) : kMaxBeamChans;
        a.accumulate = c0 > 0;
        for (int c = 0; c < a.nchan; c++) {
            a.chan[c] = (const uint8_t *)chans[c0 + c];
            a.w[c] = make_float2(weights[2 * (c0 + c)] * wscale, weights[2 * (c0 + c) + 1] * wscale);
        }
        // read span: the hull of this launch's channel buffers (conservative)
        uintptr_t lo = (uintptr_t)a.chan[0], hi = lo;
        for (int c = 0; c < a.nchan; c++) {
            const uintptr_t p = (uintptr_t)a.chan[c];
            lo = p < lo ? p : lo;
            hi = p > hi ? p : hi;
        }
        const OverlapWindow::Span r{lo, hi + n * (size_t)sb}, w = OverlapWindow::span(dst, n * 8);
        int rc;
        switch (src_format) {
            case HZSDR_FORMAT_U8: rc = launch_overlapping(ctx, k_beamform<HZSDR_FORMAT_U8>, grid, r, w, (float4 *)dst, n / 4, a); break;
            case HZSDR_FORMAT_I8: rc = launch_overlapping(ctx, k_beamform<HZSDR_FORMAT_I8>, grid, r, w, (float4 *)dst, n / 4, a); break;
            default: rc = launch_overlapping(ctx, k_beamform<HZSDR_FORMAT_I16>, grid, r, w, (float4 *)dst, n / 4, a); break;
        }
        if (rc) return rc;
    }
    return HZSDR_OK;
}

// End to end: chans_host are HOST buffers (pinned for the copies to overlap), dst_host a host
// buffer.  The time axis is cut into pieces that go through the context's staging pipe: the raw
// channels of piece k+1 cross PCIe while piece k is summed and piece k-1's beam travels back.
extern "C" int hzsdr_beamform_submit_host(hzsdr_ctx *ctx, int src_format, const void *const *chans_host, int nchan,
                                          const float *weights, size_t n, void *dst_host) {
    HZ_ENTER(ctx);
    if (nchan < 1 || !chans_host || !weights) return fail(HZSDR_ERR_INVALID, "hzsdr_beamform_submit_host: no channels");
    if (src_format != HZSDR_FORMAT_U8 && src_format != HZSDR_FORMAT_I8 && src_format != HZSDR_FORMAT_I16)
        return fail(HZSDR_ERR_FORMAT_UNKNOWN, "hzsdr_beamform_submit_host: raw source format expected, got %d", src_format);
    if (n == 0) return HZSDR_OK;
    if (n % 4 || !dst_host) return fail(HZSDR_ERR_INVALID, "hzsdr_beamform_submit_host: n must be a multiple of 4, dst non-null");
    for (int c = 0; c < nchan; c++)
        if (!chans_host[c]) return fail(HZSDR_ERR_INVALID, "hzsdr_beamform_submit_host: channel %d is null", c);
    const size_t sb = (size_t)hzsdr_format_size(src_format);
    HostPipe &hp = ctx->host_pipe;
    HZ_CUDA(hp.init());
    // piece: about 32 MiB of raw samples over all channels, a multiple of 4 samples
    size_t piece = (((size_t)32 << 20) / (sb * (size_t)nchan)) & ~(size_t)3;
    if (piece < 4096) piece = 4096;
    if (piece > n) piece = n;
    // channels laid out at one pitch in host memory (one pinned block) travel as ONE 2-D copy per piece
    bool pitched = nchan > 1;
    const ptrdiff_t pitch = nchan > 1 ? (const uint8_t *)chans_host[1] - (const uint8_t *)chans_host[0] : 0;
    for (int c = 1; c < nchan && pitched; c++)
        pitched = (const uint8_t *)chans_host[c] - (const uint8_t *)chans_host[c - 1] == pitch;
    pitched = pitched && pitch >= (ptrdiff_t)(n * sb);
    std::vector<const void *> dev(nchan);
    for (size_t off = 0; off < n; off += piece) {
        const size_t len = n - off < piece ? n - off : piece;
        HostPipe::Slot *sl = nullptr;
        HZ_CUDA(hp.next(ctx->stream, (size_t)nchan * piece * sb, piece * 8, &sl));
        for (int c = 0; c < nchan; c++) dev[c] = (const uint8_t *)sl->in + (size_t)c * piece * sb;
        if (pitched) {
            HZ_CUDA(cudaMemcpy2DAsync(sl->in, piece * sb, (const uint8_t *)chans_host[0] + off * sb, (size_t)pitch, len * sb,
                                      (size_t)nchan, cudaMemcpyHostToDevice, hp.copy_in));
        } else {
            for (int c = 0; c < nchan; c++)
                HZ_CUDA(cudaMemcpyAsync((void *)dev[c], (const uint8_t *)chans_host[c] + off * sb, len * sb, cudaMemcpyHostToDevice,
                                        hp.copy_in));
        }
        HZ_CUDA(hp.before_kernel(ctx->stream, *sl));
        int rc = hzsdr_beamform(ctx, src_format, dev.data(), nchan, weights, len, sl->out);
        if (rc) return rc;
        HZ_CUDA(hp.after_kernel(ctx->stream, *sl));
        HZ_CUDA(cudaMemcpyAsync((uint8_t *)dst_host + off * 8, sl->out, len * 8, cudaMemcpyDeviceToHost, hp.copy_out));
        HZ_CUDA(hp.done(*sl));
    }
    return HZSDR_OK;
}
