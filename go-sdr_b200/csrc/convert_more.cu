// convert_more.cu -- SURVEY.md 8(f) rank 3: the rest of ConvertBuffer's 4x4 format matrix
// (conv.go:36-46) and the wrapping integer adds of stream.Add.  Bit-exact integer work.
//   c64 -> u8 / i16 / i8   iq_c64.go:77-117  (fp32 multiply, separate add, truncate toward zero)
//   u8 <-> i8 <-> i16      iq_u8.go:73-101, iq_i8.go:73-97, iq_i16.go:116-134,150-162
//   Add on i8 / i16        stream/add.go:95-113
#include "common.cuh"

namespace hz {

constexpr int kCmThreads = 256;

// Go's float32 -> narrow integer conversion as gc emits it on amd64: CVTTSS2SL (truncate toward
// zero to int32), then keep the low bits.  In-range values are simply truncated.
__device__ __forceinline__ uint32_t trunc_low(float x) { return (uint32_t)__float2int_rz(x); }

template <int DST>
__device__ __forceinline__ uint32_t from_c64_one(float2 v) {
    if constexpr (DST == HZSDR_FORMAT_U8) {  // uint8(real*127.5 + 127.5): two roundings, not an FMA
        const uint32_t a = trunc_low(__fadd_rn(__fmul_rn(v.x, 127.5f), 127.5f)) & 0xffu;
        const uint32_t b = trunc_low(__fadd_rn(__fmul_rn(v.y, 127.5f), 127.5f)) & 0xffu;
        return a | (b << 8);
    } else if constexpr (DST == HZSDR_FORMAT_I8) {  // int8(real * math.MaxInt8)
        return (trunc_low(__fmul_rn(v.x, 127.0f)) & 0xffu) | ((trunc_low(__fmul_rn(v.y, 127.0f)) & 0xffu) << 8);
    } else {  // int16(real * math.MaxInt16)
        return (trunc_low(__fmul_rn(v.x, 32767.0f)) & 0xffffu) | ((trunc_low(__fmul_rn(v.y, 32767.0f)) & 0xffffu) << 16);
    }
}

template <int DST>
__global__ void __launch_bounds__(kCmThreads) k_from_c64(const float2 *__restrict__ src, void *__restrict__ dst, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t w = from_c64_one<DST>(ld_stream_f2(src + i));
        if constexpr (DST == HZSDR_FORMAT_I16)
            reinterpret_cast<uint32_t *>(dst)[i] = w;
        else
            reinterpret_cast<uint16_t *>(dst)[i] = (uint16_t)w;
    }
}

// one IQ sample, integer to integer
template <int SRC, int DST>
__device__ __forceinline__ uint32_t int_convert_one(uint32_t w) {
    if constexpr ((SRC == HZSDR_FORMAT_U8 && DST == HZSDR_FORMAT_I8) || (SRC == HZSDR_FORMAT_I8 && DST == HZSDR_FORMAT_U8)) {
        return w ^ 0x8080u;  // int8(int16(b) - 128)  /  uint8(int16(b) + 128)
    } else if constexpr (SRC == HZSDR_FORMAT_U8 && DST == HZSDR_FORMAT_I16) {
        const uint32_t f = w ^ 0x8080u;  // int16((int32(b) << 8) - 32768)
        return ((f & 0xffu) << 8) | ((f & 0xff00u) << 16);
    } else if constexpr (SRC == HZSDR_FORMAT_I8 && DST == HZSDR_FORMAT_I16) {
        return ((w & 0xffu) << 8) | ((w & 0xff00u) << 16);  // int16(b) << 8
    } else if constexpr (SRC == HZSDR_FORMAT_I16 && DST == HZSDR_FORMAT_U8) {
        const uint32_t f = w ^ 0x80008000u;  // uint8(uint16(int32(v) + 32768) >> 8)
        return ((f >> 8) & 0xffu) | ((f >> 16) & 0xff00u);
    } else {  // I16 -> I8: int8(v >> 8)
        return ((w >> 8) & 0xffu) | ((w >> 16) & 0xff00u);
    }
}

template <int SRC, int DST>
__global__ void __launch_bounds__(kCmThreads) k_int_convert(const void *__restrict__ src, void *__restrict__ dst, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t w;
        if constexpr (SRC == HZSDR_FORMAT_I16)
            w = reinterpret_cast<const uint32_t *>(src)[i];
        else
            w = reinterpret_cast<const uint16_t *>(src)[i];
        const uint32_t o = int_convert_one<SRC, DST>(w);
        if constexpr (DST == HZSDR_FORMAT_I16)
            reinterpret_cast<uint32_t *>(dst)[i] = o;
        else
            reinterpret_cast<uint16_t *>(dst)[i] = (uint16_t)o;
    }
}

constexpr int kMaxIntAddSrcs = 32;
struct IntAddSrcs {
    const void *p[kMaxIntAddSrcs];
};
// out = ((0 + b0) + b1) + ... with wrapping integer adds, component-wise (stream/add.go:95-113)
template <typename T>
__global__ void __launch_bounds__(kCmThreads) k_add_int(T *__restrict__ dst, const __grid_constant__ IntAddSrcs srcs, int k,
                                                         size_t ncomp, bool accumulate) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < ncomp; i += stride) {
        T acc = accumulate ? dst[i] : (T)0;
        for (int c = 0; c < k; c++) acc = (T)(acc + reinterpret_cast<const T *>(srcs.p[c])[i]);
        dst[i] = acc;
    }
}

static inline int cm_grid(const hzsdr_ctx *ctx, size_t items) {
    size_t g = (items + kCmThreads - 1) / kCmThreads;
    const size_t cap = (size_t)ctx->sm_count * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace hz

using namespace hz;

static bool is_raw(int f) { return f == HZSDR_FORMAT_U8 || f == HZSDR_FORMAT_I8 || f == HZSDR_FORMAT_I16; }

// sdr.ConvertBuffer, the full matrix (conv.go:55-93)
extern "C" int hzsdr_convert(hzsdr_ctx *ctx, int src_format, const void *src, size_t src_len, int dst_format, void *dst,
                             size_t dst_len, size_t *n_out) {
    HZ_ENTER(ctx);
    if (n_out) *n_out = 0;
    if (!hzsdr_format_size(src_format) || !hzsdr_format_size(dst_format))
        return fail(HZSDR_ERR_FORMAT_UNKNOWN, "hzsdr_convert: unknown format %d -> %d", src_format, dst_format);
    if (dst_format == HZSDR_FORMAT_C64) return hzsdr_convert_to_c64(ctx, src_format, src, src_len, dst, dst_len, n_out);
    if (src_format == dst_format) {  // CopySamples: min of the two lengths (copy.go:31-52)
        const size_t n = src_len < dst_len ? src_len : dst_len;
        if (n) HZ_CUDA(cudaMemcpyAsync(dst, src, n * hzsdr_format_size(src_format), cudaMemcpyDeviceToDevice, ctx->stream));
        if (n_out) *n_out = n;
        return HZSDR_OK;
    }
    if (src_len > dst_len) return fail(HZSDR_ERR_DST_TOO_SMALL, "hzsdr_convert: %zu > %zu", src_len, dst_len);
    const size_t n = src_len;
    if (n == 0) return HZSDR_OK;
    if (!src || !dst || ((uintptr_t)src % hzsdr_format_size(src_format)) || ((uintptr_t)dst % hzsdr_format_size(dst_format)))
        return fail(HZSDR_ERR_INVALID, "hzsdr_convert: null or misaligned buffer");
    const int grid = cm_grid(ctx, n);
    cudaStream_t st = ctx->stream;
#define HZ_CASE(S, D)                                                                             \
    if (src_format == S && dst_format == D) {                                                     \
        k_int_convert<S, D><<<grid, kCmThreads, 0, st>>>(src, dst, n);                            \
    } else
    if (src_format == HZSDR_FORMAT_C64) {
        switch (dst_format) {
            case HZSDR_FORMAT_U8: k_from_c64<HZSDR_FORMAT_U8><<<grid, kCmThreads, 0, st>>>((const float2 *)src, dst, n); break;
            case HZSDR_FORMAT_I8: k_from_c64<HZSDR_FORMAT_I8><<<grid, kCmThreads, 0, st>>>((const float2 *)src, dst, n); break;
            default: k_from_c64<HZSDR_FORMAT_I16><<<grid, kCmThreads, 0, st>>>((const float2 *)src, dst, n); break;
        }
    } else if (is_raw(src_format) && is_raw(dst_format)) {
        HZ_CASE(HZSDR_FORMAT_U8, HZSDR_FORMAT_I8)
        HZ_CASE(HZSDR_FORMAT_U8, HZSDR_FORMAT_I16)
        HZ_CASE(HZSDR_FORMAT_I8, HZSDR_FORMAT_U8)
        HZ_CASE(HZSDR_FORMAT_I8, HZSDR_FORMAT_I16)
        HZ_CASE(HZSDR_FORMAT_I16, HZSDR_FORMAT_U8)
        HZ_CASE(HZSDR_FORMAT_I16, HZSDR_FORMAT_I8) { return fail(HZSDR_ERR_CONVERSION_NOT_IMPLEMENTED, "hzsdr_convert: %d -> %d", src_format, dst_format); }
    } else {
        return fail(HZSDR_ERR_CONVERSION_NOT_IMPLEMENTED, "hzsdr_convert: %d -> %d", src_format, dst_format);
    }
#undef HZ_CASE
    HZ_CHECK_LAUNCH();
    if (n_out) *n_out = n;
    return HZSDR_OK;
}

// stream.Add on I8 / I16 readers (stream/add.go:95-113,169-182)
extern "C" int hzsdr_add_int(hzsdr_ctx *ctx, int format, void *dst, const void *const *srcs, int k, size_t n) {
    HZ_ENTER(ctx);
    if (format != HZSDR_FORMAT_I8 && format != HZSDR_FORMAT_I16)
        return fail(HZSDR_ERR_FORMAT_UNKNOWN, "hzsdr_add_int: I8 or I16 expected (stream/add.go:56-61), got %d", format);
    if (k < 1 || !srcs) return fail(HZSDR_ERR_INVALID, "hzsdr_add_int: no sources");
    if (n == 0) return HZSDR_OK;
    const size_t ncomp = 2 * n;
    const int grid = cm_grid(ctx, ncomp);
    for (int c0 = 0; c0 < k; c0 += kMaxIntAddSrcs) {
        IntAddSrcs a;
        const int kk = (k - c0) < kMaxIntAddSrcs ? (k - c0) : kMaxIntAddSrcs;
        for (int c = 0; c < kk; c++) {
            if (!srcs[c0 + c]) return fail(HZSDR_ERR_INVALID, "hzsdr_add_int: null source %d", c0 + c);
            a.p[c] = srcs[c0 + c];
        }
        if (format == HZSDR_FORMAT_I8)
            k_add_int<int8_t><<<grid, kCmThreads, 0, ctx->stream>>>((int8_t *)dst, a, kk, ncomp, c0 > 0);
        else
            k_add_int<int16_t><<<grid, kCmThreads, 0, ctx->stream>>>((int16_t *)dst, a, kk, ncomp, c0 > 0);
        HZ_CHECK_LAUNCH();
    }
    return HZSDR_OK;
}
