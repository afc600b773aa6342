// NVLink write bandwidth from SM code, one process, GPU 0 -> GPU 1 (cudaDeviceEnablePeerAccess):
//   A  per-lane float4 stores, a warp writes 512 contiguous bytes per instruction (the k_beamform_rs pattern)
//   B  the same through shared memory + cp.async.bulk (TMA bulk copy shared -> peer global), 4 KB per copy
// at several grid sizes; plus local stores and cudaMemcpyPeer as yardsticks.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) k_st(float4 *dst, size_t nvec, int iters) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int it = 0; it < iters; it++)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride)
            dst[i] = make_float4(i, it, 1.f, 2.f);
}

__global__ void __launch_bounds__(256) k_bulk(float4 *dst, size_t nvec, int iters) {
    // each warp: fill 4 KB of shared memory, fence, one lane issues cp.async.bulk to global, waits on the bulk group
    __shared__ __align__(128) float4 stage[8][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t nchunks = nvec / 256, wstride = (size_t)gridDim.x * 8;
    for (int it = 0; it < iters; it++)
        for (size_t c = (size_t)blockIdx.x * 8 + warp; c < nchunks; c += wstride) {
#pragma unroll
            for (int k = 0; k < 8; k++) stage[warp][lane + 32 * k] = make_float4(c, it, lane, k);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                const uint32_t s = (uint32_t)__cvta_generic_to_shared(&stage[warp][0]);
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 4096;" ::"l"(dst + c * 256), "r"(s) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the source may be rewritten
            }
            __syncwarp();
        }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <class F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
    int n; cudaGetDeviceCount(&n);
    if (n < 2) { printf("need 2 GPUs\n"); return 0; }
    const size_t bytes = 256u << 20, nvec = bytes / 16;
    float4 *local, *peer;
    cudaSetDevice(1); cudaMalloc(&peer, bytes); cudaMemset(peer, 0, bytes);
    cudaSetDevice(0); cudaMalloc(&local, bytes);
    cudaError_t e = cudaDeviceEnablePeerAccess(1, 0);
    printf("peer access: %s\n", cudaGetErrorString(e));
    const int iters = 4;
    for (int grid : {148, 296, 592, 1184}) {
        float a = timeit([&] { k_st<<<grid, 256>>>(peer, nvec, iters); });
        float b = timeit([&] { k_bulk<<<grid, 256>>>(peer, nvec, iters); });
        float l = timeit([&] { k_st<<<grid, 256>>>(local, nvec, iters); });
        printf("grid %4d: peer float4 stores %7.1f GB/s   peer cp.async.bulk %7.1f GB/s   local stores %7.1f GB/s\n", grid,
               bytes * iters / a / 1e6, bytes * iters / b / 1e6, bytes * iters / l / 1e6);
    }
    float m = timeit([&] { for (int i = 0; i < iters; i++) cudaMemcpyPeerAsync(peer, 1, local, 0, bytes); });
    printf("cudaMemcpyPeer %7.1f GB/s   (%s)\n", bytes * iters / m / 1e6, cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
