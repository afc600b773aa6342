// Microbenchmark: scalar FFMA/FADD vs packed fma.rn.f32x2 / add.rn.f32x2 issue rate on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exp_ffma2 exp_ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

template <int MODE> __global__ void __launch_bounds__(256) k(float* out, int iters, float s, float t) {
    const int NA = 8;
    if (MODE == 0) {          // scalar FFMA, 16 independent chains (same flops as MODE 1)
        float a[2 * NA];
        for (int i = 0; i < 2 * NA; i++) a[i] = threadIdx.x + i;
        for (int it = 0; it < iters; it++)
#pragma unroll
            for (int i = 0; i < 2 * NA; i++) a[i] = fmaf(a[i], s, t);
        float r = 0; for (int i = 0; i < 2 * NA; i++) r += a[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    } else if (MODE == 1) {   // packed FFMA2, 8 independent chains
        u64 a[NA]; u64 ss = pk(s, s), tt = pk(t, t);
        for (int i = 0; i < NA; i++) a[i] = pk(threadIdx.x + i, threadIdx.x - i);
        for (int it = 0; it < iters; it++)
#pragma unroll
            for (int i = 0; i < NA; i++) a[i] = fma2(a[i], ss, tt);
        float r = 0; for (int i = 0; i < NA; i++) { float x, y; upk(a[i], x, y); r += x + y; }
        out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    } else if (MODE == 2) {   // scalar FADD
        float a[2 * NA];
        for (int i = 0; i < 2 * NA; i++) a[i] = threadIdx.x + i;
        for (int it = 0; it < iters; it++)
#pragma unroll
            for (int i = 0; i < 2 * NA; i++) a[i] = a[i] + t;
        float r = 0; for (int i = 0; i < 2 * NA; i++) r += a[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    } else if (MODE == 3) {   // packed FADD2
        u64 a[NA]; u64 tt = pk(t, s);
        for (int i = 0; i < NA; i++) a[i] = pk(threadIdx.x + i, threadIdx.x - i);
        for (int it = 0; it < iters; it++)
#pragma unroll
            for (int i = 0; i < NA; i++) a[i] = add2(a[i], tt);
        float r = 0; for (int i = 0; i < NA; i++) { float x, y; upk(a[i], x, y); r += x + y; }
        out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    } else if (MODE == 4) {   // scalar FFMA all-register operands (3 distinct regs)
        float a[2 * NA], b[2 * NA];
        for (int i = 0; i < 2 * NA; i++) { a[i] = threadIdx.x + i; b[i] = s + i; }
        for (int it = 0; it < iters; it++)
#pragma unroll
            for (int i = 0; i < 2 * NA; i++) a[i] = fmaf(a[i], b[i], b[(i + 1) % (2 * NA)]);
        float r = 0; for (int i = 0; i < 2 * NA; i++) r += a[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    } else if (MODE == 5) {   // packed mul2 + add2 mix
        u64 a[NA]; u64 tt = pk(t, s), ss = pk(s, s);
        for (int i = 0; i < NA; i++) a[i] = pk(threadIdx.x + i, threadIdx.x - i);
        for (int it = 0; it < iters; it++)
#pragma unroll
            for (int i = 0; i < NA; i++) a[i] = (i & 1) ? add2(a[i], tt) : mul2(a[i], ss);
        float r = 0; for (int i = 0; i < NA; i++) { float x, y; upk(a[i], x, y); r += x + y; }
        out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    }
}
template <int MODE> void run(const char* name, float* d, int flops_per_iter_thread) {
    int iters = 4096, blocks = 148 * 8, thr = 256;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, thr>>>(d, 16, 1.0001f, 0.5f);
    cudaEventRecord(e0);
    k<MODE><<<blocks, thr>>>(d, iters, 1.0001f, 0.5f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = double(blocks) * thr * iters * flops_per_iter_thread;   // lane-ops (fma counts 1)
    printf("%-28s %8.3f ms  %8.2f Tlane-op/s  (%6.1f lane-ops/clk/SM at 1.9GHz)\n", name, ms, ops / ms * 1e-9,
           ops / (ms * 1e-3) / 148 / 1.9e9);
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    run<0>("scalar FFMA (imm/const ops)", d, 16);
    run<4>("scalar FFMA (3 regs)", d, 16);
    run<1>("packed FFMA2", d, 16);
    run<2>("scalar FADD", d, 16);
    run<3>("packed FADD2", d, 16);
    run<5>("packed FMUL2+FADD2", d, 16);
    cudaError_t e = cudaDeviceSynchronize(); printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
