// Microbenchmark (sm_100a): can per-lane tables live in Tensor Memory instead of shared memory?
//   A  32 x LDS.64 per iteration (conflict-free, the chain kernel's table/exchange pattern)
//   B  tcgen05.ld.32x32b.x16 x 4 per iteration (64 registers = 32 float2 per lane: one whole table row set)
//   C  A + B in the same iteration: independent data paths => time ~ max(A, B)
//   D  B + 64 FFMA2 per iteration: does LDTM disturb the FMA pipe?
// 4 CTAs of 4 warps per SM (the chain1024 shape), 128 TMEM columns per CTA.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exp_tmem exp_tmem.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_alloc(uint32_t *smem_slot, uint32_t cols) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_slot);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
                 "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

template <int MODE>
__global__ void __launch_bounds__(128, 4) k(float *out, int iters, int check) {
    __shared__ uint32_t slot;
    __shared__ float2 tab[4][32 * 33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 4 * 32 * 33; i += 128) (&tab[0][0])[i] = make_float2(i * 0.001f, 1.0f - i * 0.002f);
    if (warp == 0) tmem_alloc(&slot, 128);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot;
    const uint32_t mine = base + ((uint32_t)(32 * warp) << 16);  // lanes 32 warp .. 32 warp + 31
    // fill: column c of lane l = float(1000 warp + 32 c + l)  (so a read-back check is possible)
    for (int c0 = 0; c0 < 128; c0 += 16) {
        uint32_t r[16];
        for (int i = 0; i < 16; i++) r[i] = __float_as_uint((float)(1000 * warp + 32 * (c0 + i) + lane));
        tmem_st16(mine + c0, r);
    }
    tmem_wait_st();
    float acc = 0.f;
    u64 a[8];
    for (int i = 0; i < 8; i++) a[i] = pk(lane + i, lane - i);
    const float2 *t = tab[warp];
    if (check) {
        uint32_t r[16];
        int bad = 0;
        for (int c0 = 0; c0 < 128; c0 += 16) {
            tmem_ld16(mine + c0, r);
            tmem_wait_ld();
            for (int i = 0; i < 16; i++) bad += __uint_as_float(r[i]) != (float)(1000 * warp + 32 * (c0 + i) + lane);
        }
        out[blockIdx.x * 128 + threadIdx.x] = (float)bad;
    } else {
        for (int it = 0; it < iters; it++) {
            if (MODE == 0 || MODE == 2) {
#pragma unroll
                for (int r = 0; r < 32; r++) {
                    const float2 v = t[lane + 33 * r];
                    acc += v.x * v.y;
                }
            }
            if (MODE == 1 || MODE == 2 || MODE == 3) {
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    uint32_t r[16];
                    tmem_ld16(mine + 16 * q + (it & 1) * 64, r);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 16; i += 2) acc += __uint_as_float(r[i]) * __uint_as_float(r[i + 1]);
                }
            }
            if (MODE == 3 || MODE == 4) {
#pragma unroll
                for (int q = 0; q < 8; q++)
#pragma unroll
                    for (int i = 0; i < 8; i++) a[i] = fma2(a[i], pk(0.7071067f, 0.7071067f), a[(i + 1) & 7]);
            }
        }
        float r = acc;
        for (int i = 0; i < 8; i++) r += (float)(a[i] & 0xff);
        out[blockIdx.x * 128 + threadIdx.x] = r;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) tmem_dealloc(base, 128);
}

template <int MODE>
void run(const char *name, float *d) {
    const int iters = 2048, blocks = 148 * 4;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE><<<blocks, 128>>>(d, 16, 0);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 128>>>(d, iters, 0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double cyc = ms * 1e-3 * 1.965e9 / iters;  // SM cycles per iteration (16 warps per SM each doing one iteration)
    printf("%-52s %8.3f ms  %8.1f SM-cycles per iteration of all 16 warps (%.1f per warp-iteration)\n", name, ms, cyc, cyc / 16.0);
}

int main() {
    float *d;
    cudaMalloc(&d, 148 * 4 * 128 * 4);
    k<1><<<148 * 4, 128>>>(d, 0, 1);
    float *h = (float *)malloc(148 * 4 * 128 * 4);
    cudaMemcpy(h, d, 148 * 4 * 128 * 4, cudaMemcpyDeviceToHost);
    double bad = 0;
    for (int i = 0; i < 148 * 4 * 128; i++) bad += h[i];
    printf("read-back mismatches: %.0f (%s)\n", bad, cudaGetErrorString(cudaGetLastError()));
    run<0>("A: 32 x LDS.64 (64 wavefronts per warp-iteration)", d);
    run<1>("B: 4 x LDTM.32x32b.x16 (8 KB per warp-iteration)", d);
    run<2>("C: A + B", d);
    run<4>("E: 64 FFMA2", d);
    run<3>("D: B + 64 FFMA2", d);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
