// How many CTAs of one SM can hold Tensor Memory columns at the same time?  Each CTA allocates `cols`
// columns, notes the time, burns ~20 us of FFMA, notes the time again, deallocates.  The host counts,
// per SM, the largest number of [after-alloc, before-dealloc] intervals that overlap.
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t gtime() { uint64_t t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ uint32_t smid() { uint32_t s; asm volatile("mov.u32 %0, %smid;" : "=r"(s)); return s; }
struct Rec { uint64_t t0, t1, t2; uint32_t sm, pad; };
template <int COLS>
__global__ void __launch_bounds__(128, 4) k(Rec *rec, int iters, float *out) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    const uint64_t tstart = gtime();
    if (COLS > 0 && warp == 0) {
        const uint32_t s = (uint32_t)__cvta_generic_to_shared(&slot);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s), "r"(COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint64_t t0 = gtime();
    float a = threadIdx.x, b = 1.0001f;
    for (int i = 0; i < iters; i++) a = fmaf(a, b, 0.5f);
    out[blockIdx.x * 128 + threadIdx.x] = a;
    const uint64_t t1 = gtime();
    __syncthreads();
    if (COLS > 0 && warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(COLS) : "memory");
    if (threadIdx.x == 0) rec[blockIdx.x] = Rec{tstart, t0, t1, smid(), 0};
}
template <int COLS> void run(Rec *d, float *out, int blocks) {
    k<COLS><<<blocks, 128>>>(d, 1000, out);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<COLS><<<blocks, 128>>>(d, 10000, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<Rec> h(blocks);
    cudaMemcpy(h.data(), d, sizeof(Rec) * blocks, cudaMemcpyDeviceToHost);
    int worst = 0, worst_res = 0; double wait_ns = 0;
    for (int sm = 0; sm < 160; sm++) {
        std::vector<std::pair<uint64_t,int>> ev, ev2;
        for (auto &r : h) if ((int)r.sm == sm) { ev.push_back({r.t1, +1}); ev.push_back({r.t2, -1}); ev2.push_back({r.t0, +1}); ev2.push_back({r.t2, -1}); }
        std::sort(ev.begin(), ev.end()); std::sort(ev2.begin(), ev2.end());
        int c = 0; for (auto &e : ev) { c += e.second; worst = std::max(worst, c); }
        c = 0; for (auto &e : ev2) { c += e.second; worst_res = std::max(worst_res, c); }
    }
    for (auto &r : h) wait_ns += (double)(r.t1 - r.t0);
    printf("cols %3d: %4d CTAs, %.3f ms; max CTAs resident on one SM %d, max holding TMEM at once %d; mean alloc wait %.1f us (%s)\n", COLS, blocks, ms,
           worst_res, worst, wait_ns / blocks / 1e3, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    Rec *d; float *out;
    const int blocks = 148 * 8;
    cudaMalloc(&d, sizeof(Rec) * blocks); cudaMalloc(&out, blocks * 128 * 4);
    run<0>(d, out, blocks);
    run<32>(d, out, blocks);
    run<64>(d, out, blocks);
    run<128>(d, out, blocks);
    run<256>(d, out, blocks);
    run<512>(d, out, blocks);
    printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
