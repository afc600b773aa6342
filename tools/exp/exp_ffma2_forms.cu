// Microbenchmark: pipe cycles per packed fp32 instruction by operand form (sm_100a).
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

template <int FORM> __global__ void __launch_bounds__(256) k(float* out, int iters, float s, float t) {
    u64 a[8], b[8];
    float sc[8];
    for (int i = 0; i < 8; i++) { a[i] = pk(threadIdx.x + i, threadIdx.x - i); b[i] = pk(s + i + threadIdx.x, t - i - threadIdx.x); sc[i] = s * (i + 1 + threadIdx.x); }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            float bx, by; upk(b[i], bx, by);
            if (FORM == 0) a[i] = add2(a[i], b[i]);                              // pair + pair
            if (FORM == 1) a[i] = add2(a[i], pk(by, -bx));                       // pair + swapped/negated pair
            if (FORM == 2) a[i] = mul2(a[i], pk(sc[i], sc[i]));                  // pair * scalar reg
            if (FORM == 3) a[i] = fma2(a[i], pk(0.7071067f, 0.7071067f), b[i]);  // pair * imm + pair
            if (FORM == 4) a[i] = fma2(a[i], pk(sc[i], sc[i]), b[i]);            // pair * scalar reg + pair
            if (FORM == 5) a[i] = fma2(a[i], b[i], b[(i + 1) & 7]);              // pair * pair + pair
            if (FORM == 6) a[i] = fma2(pk(by, -bx), pk(sc[i], sc[i]), a[i]);     // swapped/neg pair * scalar + pair (acc)
            if (FORM == 7) a[i] = fma2(b[i], pk(2.0f, 2.0f), a[i]);              // pair * imm + acc
            if (FORM == 8) a[i] = mul2(pk(-by, bx), pk(sc[i], sc[i]));           // no dependence on a: pure throughput
        }
    }
    float r = 0;
    for (int i = 0; i < 8; i++) { float p, q; upk(a[i], p, q); r += p + q; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int FORM> void run(const char* name, float* d) {
    int iters = 4096, blocks = 148 * 8, thr = 256;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<FORM><<<blocks, thr>>>(d, 16, 1.0001f, 0.5f);
    cudaEventRecord(e0);
    k<FORM><<<blocks, thr>>>(d, iters, 1.0001f, 0.5f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double cyc = ms * 1e-3 * 1.965e9 / iters / 16.0 / 8.0;
    printf("%-44s %8.3f ms  %5.2f cycles per instruction per SMSP\n", name, ms, cyc);
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    run<0>("FADD2 pair + pair", d);
    run<1>("FADD2 pair + swapped/neg pair", d);
    run<2>("FMUL2 pair * scalar", d);
    run<3>("FFMA2 pair * imm + pair", d);
    run<4>("FFMA2 pair * scalar + pair", d);
    run<5>("FFMA2 pair * pair + pair", d);
    run<6>("FFMA2 swapped/neg pair * scalar + pair", d);
    run<7>("FFMA2 pair * imm + pair (acc is c)", d);
    run<8>("FMUL2 swapped/neg pair * scalar (no dep)", d);
    cudaError_t e = cudaDeviceSynchronize(); printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
