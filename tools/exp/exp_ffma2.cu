// Microbenchmark: issue behaviour of packed fp32 (FFMA2) on sm_100a: alone, and interleaved 1:1 with
// integer ALU ops, scalar FFMA and shared-memory loads.  Answers: does a packed op hold the issue
// port for 2 cycles, or only its pipe?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exp_ffma2 exp_ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ unsigned lop(unsigned a, unsigned b) { unsigned r; asm volatile("add.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ unsigned prmt(unsigned a, unsigned b) { unsigned r; asm volatile("prmt.b32 %0, %1, %2, 0x2103;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ unsigned imad(unsigned a, unsigned b, unsigned c) { unsigned r; asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }

// NP packed ops + NA "other" ops per iteration, all independent chains of length 8
template <int OTHER, int NP, int NA> __global__ void __launch_bounds__(256) k(float* out, int iters, float s, float t, unsigned key) {
    __shared__ float sm[256 * 9];
    u64 a[8]; unsigned x[8], y[8]; float f[8];
    u64 ss = pk(s, s), tt = pk(t, t * 0.5f + threadIdx.x);
    for (int i = 0; i < 8; i++) { a[i] = pk(threadIdx.x + i, threadIdx.x - i); x[i] = threadIdx.x * 17 + i; y[i] = threadIdx.x * 3 + i; f[i] = i + threadIdx.x; }
    for (int i = threadIdx.x; i < 256 * 9; i += 256) sm[i] = i;
    __syncthreads();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (i < NP) a[i] = fma2(a[i], ss, tt);
            if (i < NA) {
                if (OTHER == 0) x[i] = lop(x[i], key);
                if (OTHER == 1) f[i] = ffma(f[i], s, t);
                if (OTHER == 2) x[i] = imad(x[i], key, key);
                if (OTHER == 5) x[i] = prmt(x[i], key + i);
                if (OTHER == 6) { x[i] = lop(x[i], key); y[i] = prmt(y[i], key + i); }
                if (OTHER == 3) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"((unsigned)__cvta_generic_to_shared(sm + threadIdx.x + 256 * i))); f[i] += v; }
                if (OTHER == 4) { float2 v; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"((unsigned)__cvta_generic_to_shared(sm + 2 * threadIdx.x + 256 * i))); f[i] += v.x; }
            }
        }
    }
    float r = 0;
    for (int i = 0; i < 8; i++) { float p, q; upk(a[i], p, q); r += p + q + x[i] + f[i] + y[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int OTHER, int NP, int NA> void run(const char* name, float* d) {
    int iters = 4096, blocks = 148 * 8, thr = 256;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OTHER, NP, NA><<<blocks, thr>>>(d, 16, 1.0001f, 0.5f, 0x9e3779b9u);
    cudaEventRecord(e0);
    k<OTHER, NP, NA><<<blocks, thr>>>(d, iters, 1.0001f, 0.5f, 0x9e3779b9u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // cycles per iteration per SMSP: 16 warps per SMSP (8 blocks x 8 warps / 4), clock 1.965 GHz
    double cyc = ms * 1e-3 * 1.965e9 / iters / 16.0;
    printf("%-34s %8.3f ms  %6.2f cycles per warp-iteration (%d packed + %d other)\n", name, ms, cyc, NP, NA);
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    run<0, 8, 0>("FFMA2 x8", d);
    run<0, 0, 8>("IADD x8", d);
    run<0, 8, 8>("FFMA2 x8 + IADD x8", d);
    run<0, 4, 8>("FFMA2 x4 + IADD x8", d);
    run<5, 0, 8>("PRMT x8", d);
    run<5, 8, 8>("FFMA2 x8 + PRMT x8", d);
    run<6, 0, 8>("IADD x8 + PRMT x8", d);
    run<6, 8, 8>("FFMA2 x8 + IADD x8 + PRMT x8", d);
    run<1, 0, 8>("FFMA x8", d);
    run<1, 8, 8>("FFMA2 x8 + FFMA x8", d);
    run<1, 4, 8>("FFMA2 x4 + FFMA x8", d);
    run<2, 0, 8>("IMAD x8", d);
    run<2, 8, 8>("FFMA2 x8 + IMAD x8", d);
    run<3, 0, 8>("LDS.32 x8 (+FADD)", d);
    run<3, 8, 8>("FFMA2 x8 + LDS.32 x8 (+FADD)", d);
    run<4, 0, 8>("LDS.64 x8 (+FADD)", d);
    run<4, 8, 8>("FFMA2 x8 + LDS.64 x8 (+FADD)", d);
    cudaError_t e = cudaDeviceSynchronize(); printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
