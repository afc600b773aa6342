// Experiment harness (not product code): variants of the write-heavy streaming kernels, to find
// what limits K1 (u8 -> c64, 2 B read + 8 B written per sample) at 58% of the copy peak.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../go-sdr_b200/csrc/common.cuh"
using namespace hz;

template <int UNROLL, int MODE>
__global__ void __launch_bounds__(256) conv_pairs(const uint8_t *__restrict__ src, float4 *__restrict__ out, size_t npairs) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = tid; i < npairs; i += UNROLL * stride) {
        uint32_t w[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) if (i + u * stride < npairs) w[u] = ld_stream_u32(src + 4 * (i + u * stride));
#pragma unroll
        for (int u = 0; u < UNROLL; u++) if (i + u * stride < npairs) {
            float2 a = RawTraits<HZSDR_FORMAT_U8>::conv(w[u]), b = RawTraits<HZSDR_FORMAT_U8>::conv_hi(w[u]);
            float4 v = make_float4(a.x, a.y, b.x, b.y);
            if (MODE == 0) st_stream_f4(out + i + u * stride, v);
            else if (MODE == 1) out[i + u * stride] = v;
            else __stcs(out + i + u * stride, v);
        }
    }
}
// contiguous per-block tiles: block b handles pairs [b*T, (b+1)*T) in a loop of blockDim-wide rows
template <int ROWS>
__global__ void __launch_bounds__(256) conv_tiles(const uint8_t *__restrict__ src, float4 *__restrict__ out, size_t npairs) {
    const size_t tile = (size_t)ROWS * blockDim.x;
    for (size_t t0 = (size_t)blockIdx.x * tile; t0 < npairs; t0 += (size_t)gridDim.x * tile) {
        uint32_t w[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; r++) { size_t i = t0 + r * blockDim.x + threadIdx.x; if (i < npairs) w[r] = ld_stream_u32(src + 4 * i); }
#pragma unroll
        for (int r = 0; r < ROWS; r++) { size_t i = t0 + r * blockDim.x + threadIdx.x; if (i < npairs) {
            float2 a = RawTraits<HZSDR_FORMAT_U8>::conv(w[r]), b = RawTraits<HZSDR_FORMAT_U8>::conv_hi(w[r]);
            st_stream_f4(out + i, make_float4(a.x, a.y, b.x, b.y)); } }
    }
}
// 16-byte loads (8 samples) and 4 float4 stores per thread, lanes interleaved through shuffles so stores stay coalesced
__global__ void __launch_bounds__(256) conv_wide(const uint4 *__restrict__ src, float4 *__restrict__ out, size_t nvec) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        uint4 w = ld_stream_u128(src + i);
        uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            float2 a = RawTraits<HZSDR_FORMAT_U8>::conv(ww[k]), b = RawTraits<HZSDR_FORMAT_U8>::conv_hi(ww[k]);
            st_stream_f4(out + 4 * i + k, make_float4(a.x, a.y, b.x, b.y));
        }
    }
}
__global__ void __launch_bounds__(256) fill_only(float4 *__restrict__ out, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) st_stream_f4(out + i, make_float4(1, 2, 3, 4));
}
__global__ void __launch_bounds__(256) copy_f4(const float4 *__restrict__ in, float4 *__restrict__ out, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) st_stream_f4(out + i, ld_stream_f4(in + i));
}

template <class F> float timeit(F f, int reps = 20) {
    for (int i = 0; i < 3; i++) f();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a); for (int i = 0; i < reps; i++) f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / reps * 1e3f;
}
int main() {
    const size_t N = 1ull << 26, npairs = N / 2;
    uint8_t *src; float4 *out, *in2;
    cudaMalloc(&src, 2 * N); cudaMalloc(&out, 8 * N); cudaMalloc(&in2, 8 * N); cudaMemset(src, 0x5a, 2 * N); cudaMemset(in2, 0, 8 * N);
    auto rep = [&](const char *name, float us, double bytes) { printf("%-44s %8.1f us  %7.0f GB/s\n", name, us, bytes / us / 1e3); };
    const double B = 10.0 * N;
    for (int bps : {4, 8, 16, 32}) {
        int g = 148 * bps; char nm[96];
        snprintf(nm, 96, "pairs U4 no_alloc  grid=148x%d", bps); rep(nm, timeit([&] { conv_pairs<4, 0><<<g, 256>>>(src, out, npairs); }), B);
    }
    rep("pairs U1 no_alloc  grid=148x8", timeit([&] { conv_pairs<1, 0><<<148 * 8, 256>>>(src, out, npairs); }), B);
    rep("pairs U2 no_alloc  grid=148x8", timeit([&] { conv_pairs<2, 0><<<148 * 8, 256>>>(src, out, npairs); }), B);
    rep("pairs U8 no_alloc  grid=148x8", timeit([&] { conv_pairs<8, 0><<<148 * 8, 256>>>(src, out, npairs); }), B);
    rep("pairs U4 plain st  grid=148x8", timeit([&] { conv_pairs<4, 1><<<148 * 8, 256>>>(src, out, npairs); }), B);
    rep("pairs U4 st.cs     grid=148x8", timeit([&] { conv_pairs<4, 2><<<148 * 8, 256>>>(src, out, npairs); }), B);
    rep("pairs U4 no_alloc  one CTA per 1024 pairs", timeit([&] { conv_pairs<4, 0><<<(unsigned)(npairs / 1024), 256>>>(src, out, npairs); }), B);
    rep("tiles R4 grid=148x8", timeit([&] { conv_tiles<4><<<148 * 8, 256>>>(src, out, npairs); }), B);
    rep("tiles R8 grid=148x8", timeit([&] { conv_tiles<8><<<148 * 8, 256>>>(src, out, npairs); }), B);
    rep("tiles R4 one tile per CTA", timeit([&] { conv_tiles<4><<<(unsigned)(npairs / 1024), 256>>>(src, out, npairs); }), B);
    rep("wide 16B loads grid=148x8", timeit([&] { conv_wide<<<148 * 8, 256>>>((const uint4 *)src, out, N / 8); }), B);
    rep("fill only (8 B/sample written)", timeit([&] { fill_only<<<148 * 8, 256>>>(out, N / 2); }), 8.0 * N);
    rep("copy f4 (8 B read + 8 B written)", timeit([&] { copy_f4<<<148 * 8, 256>>>(in2, out, N / 2); }), 16.0 * N);
    rep("cudaMemset 8 B/sample", timeit([&] { cudaMemsetAsync(out, 0, 8 * N); }), 8.0 * N);
    rep("cudaMemcpy d2d 8+8", timeit([&] { cudaMemcpyAsync(out, in2, 8 * N, cudaMemcpyDeviceToDevice); }), 16.0 * N);
    return 0;
}
