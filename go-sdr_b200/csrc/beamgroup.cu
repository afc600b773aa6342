// beamgroup.cu -- multi-GPU Beamform with the reduction fused into the kernel over NVLink peer
// memory (one process per GPU, CUDA IPC).
//
// Channels are sharded across G ranks (hzsdr_shard.channel_shard).  Every rank computes the partial
// beam of ITS channels for ALL samples, but it never stores that partial beam locally: output
// sample slice s (n/G samples) belongs to rank s, and the kernel writes its partial sums for slice s
// straight into rank s's staging slot over NVLink, tile by tile, while the next tiles are still being
// computed.  That is a reduce-scatter whose transfer overlaps the math: each GPU sends and receives
// (G-1)/G * 8 B per output sample instead of funnelling (G-1) * 8 B into one root.
// When the kernel's last CTA has fenced its stores (system scope) it raises a flag on every peer;
// a small finishing kernel on each rank waits for the G flags and sums the G slots of its slice in
// rank order (deterministic; inside a shard the channel order is the reference's left-to-right).
//
// Staging is double-buffered by step parity.  Rank A may start step s+2 (which reuses the set of
// step s) only after its step-s+1 finishing kernel, which waited for every peer's step-s+1 flag,
// which every peer raised after ITS step-s finishing kernel in stream order -- so no peer is still
// reading the set.
#include <vector>

#include "beam.cuh"
#include "common.cuh"

namespace hz {

constexpr int kMaxRanks = 16;

struct GroupArgs {
    float4 *slot[kMaxRanks];      // slot[s]: this rank's staging slot on rank s (peer pointer), current parity
    uint32_t *flag[kMaxRanks];    // flag[s]: this rank's flag word on rank s
    unsigned int *done_counter;   // local CTA counter
    uint32_t quads_per_slice;
    uint32_t step;
    int nranks;
};

template <int FMT>
__global__ void __launch_bounds__(256) k_beamform_rs(size_t nquads, const __grid_constant__ BeamArgs a,
                                                      const __grid_constant__ GroupArgs g) {
    // Peer stores must be full lines: a warp's 32 quads (64 float4 = 1 KB, contiguous in the owner's
    // slot because slices are multiples of 128 samples) are transposed through shared memory so that
    // each of the two store instructions writes 512 contiguous bytes over NVLink.
    __shared__ float4 stage[8][64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t nround = (nquads + 31) / 32 * 32;  // whole warps iterate together (shared-memory staging)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += stride) {
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; k++) acc[k] = 0.f;
        if (i < nquads) beam_quad<FMT>(a, i, acc);
        stage[warp][2 * lane] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        stage[warp][2 * lane + 1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
        __syncwarp();
        const size_t q0 = i - lane;  // first quad of this warp (multiple of 32)
        const uint32_t owner = (uint32_t)(q0 / g.quads_per_slice);
        float4 *dst = g.slot[owner] + 2 * (q0 - (size_t)owner * g.quads_per_slice);  // peer memory unless owner == this rank
        const size_t valid = 2 * (nquads - q0 < 32 ? nquads - q0 : 32);
        if ((size_t)lane < valid) dst[lane] = stage[warp][lane];
        if ((size_t)(32 + lane) < valid) dst[32 + lane] = stage[warp][32 + lane];
        __syncwarp();
    }
    // publish: all of this CTA's stores first, then (last CTA only) the flags on every rank
    __threadfence_system();
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) last = atomicAdd(g.done_counter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (last) {
        __threadfence_system();
        if (threadIdx.x < g.nranks) {
            volatile uint32_t *f = g.flag[threadIdx.x];
            *f = g.step;
        }
        if (threadIdx.x == 0) *g.done_counter = 0;
    }
}

// out[i] = sum over ranks (in rank order) of slot_r[i], after every rank's flag shows `step`
__global__ void __launch_bounds__(256) k_beam_finish(const float4 *__restrict__ slots, size_t slot_stride_vec, const uint32_t *flags,
                                                      float4 *__restrict__ out, size_t nvec, int nranks, uint32_t step) {
    if (threadIdx.x < nranks) {
        const volatile uint32_t *f = flags + threadIdx.x * 32;  // one flag per 128-byte line
        while ((int32_t)(*f - step) < 0) __nanosleep(100);
    }
    __syncthreads();
    __threadfence_system();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        float4 acc = __ldcv(slots + i);  // peers wrote these lines: bypass any stale cached copy
        for (int r = 1; r < nranks; r++) {
            const float4 v = __ldcv(slots + (size_t)r * slot_stride_vec + i);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        st_stream_f4(out + i, acc);
    }
}

}  // namespace hz

using namespace hz;

struct hzsdr_beam_group {
    hzsdr_ctx *ctx = nullptr;
    int nranks = 0, rank = 0;
    size_t n = 0, slice = 0;     // samples per buffer, per slice
    size_t slot_bytes = 0;       // one staging slot (slice * 8 B, 256-aligned)
    size_t set_bytes = 0;        // nranks slots
    uint8_t *base = nullptr;     // local allocation: [flags 4 KB | set 0 | set 1]
    unsigned int *done_counter = nullptr;
    uint8_t *peer[kMaxRanks] = {};  // every rank's base as seen from here (peer[rank] == base)
    bool connected = false;
    uint32_t step = 0;
    // the finishing kernel spins on the peers' flags: it runs on a side stream so that the next
    // buffer's compute does not queue behind the wait
    cudaStream_t fin_stream = nullptr;
    cudaEvent_t computed = nullptr;      // this step's k_beamform_rs is done
    cudaEvent_t finished[2] = {};        // finishing kernel of the step with this parity is done
    bool fin_used[2] = {};
};

static constexpr size_t kFlagBytes = 4096;

extern "C" int hzsdr_beam_group_destroy(hzsdr_beam_group *g) {
    if (!g) return HZSDR_OK;
    HZ_ENTER(g->ctx);
    cudaStreamSynchronize(g->ctx->stream);
    for (int r = 0; r < g->nranks; r++)
        if (g->connected && r != g->rank && g->peer[r]) cudaIpcCloseMemHandle(g->peer[r]);
    if (g->fin_stream) {
        cudaStreamSynchronize(g->fin_stream);
        cudaStreamDestroy(g->fin_stream);
    }
    if (g->computed) cudaEventDestroy(g->computed);
    for (auto ev : g->finished)
        if (ev) cudaEventDestroy(ev);
    if (g->base) cudaFree(g->base);
    if (g->done_counter) cudaFree(g->done_counter);
    delete g;
    return HZSDR_OK;
}

extern "C" int hzsdr_beam_group_create(hzsdr_ctx *ctx, int nranks, int rank, size_t n, void *handle_out, hzsdr_beam_group **out) {
    HZ_ENTER(ctx);
    if (!out || !handle_out || nranks < 1 || nranks > kMaxRanks || rank < 0 || rank >= nranks)
        return fail(HZSDR_ERR_INVALID, "hzsdr_beam_group_create: bad arguments");
    *out = nullptr;
    if (n == 0 || n % ((size_t)nranks * 128)) return fail(HZSDR_ERR_INVALID, "hzsdr_beam_group_create: n must be a multiple of 128 * nranks");
    static_assert(sizeof(cudaIpcMemHandle_t) == HZSDR_IPC_HANDLE_BYTES, "IPC handle size");
    hzsdr_beam_group *g = new hzsdr_beam_group();
    g->ctx = ctx;
    g->nranks = nranks;
    g->rank = rank;
    g->n = n;
    g->slice = n / nranks;
    g->slot_bytes = (g->slice * 8 + 255) / 256 * 256;
    g->set_bytes = g->slot_bytes * nranks;
    const size_t total = kFlagBytes + 2 * g->set_bytes;
    cudaError_t e = cudaMalloc((void **)&g->base, total);
    if (e == cudaSuccess) e = cudaMemset(g->base, 0, total);
    if (e == cudaSuccess) e = cudaMalloc((void **)&g->done_counter, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemset(g->done_counter, 0, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&g->fin_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g->computed, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g->finished[0], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g->finished[1], cudaEventDisableTiming);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, g->base);
    if (e != cudaSuccess) {
        hzsdr_beam_group_destroy(g);
        return fail(HZSDR_ERR_CUDA, "hzsdr_beam_group_create: %s", cudaGetErrorString(e));
    }
    memcpy(handle_out, &h, sizeof(h));
    g->peer[rank] = g->base;
    *out = g;
    return HZSDR_OK;
}

extern "C" int hzsdr_beam_group_connect(hzsdr_beam_group *g, const void *all_handles) {
    if (!g || !all_handles) return fail(HZSDR_ERR_INVALID, "hzsdr_beam_group_connect: null");
    HZ_ENTER(g->ctx);
    for (int r = 0; r < g->nranks; r++) {
        if (r == g->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const uint8_t *)all_handles + (size_t)r * HZSDR_IPC_HANDLE_BYTES, sizeof(h));
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return fail(HZSDR_ERR_CUDA, "hzsdr_beam_group_connect: rank %d: %s", r, cudaGetErrorString(e));
        g->peer[r] = (uint8_t *)p;
    }
    g->connected = true;
    return HZSDR_OK;
}

extern "C" int hzsdr_beam_group_exec(hzsdr_beam_group *g, int src_format, const void *const *chans, int nchan,
                                     const float *weights, void *dst_slice) {
    if (!g) return fail(HZSDR_ERR_INVALID, "hzsdr_beam_group_exec: null");
    HZ_ENTER(g->ctx);
    if (!g->connected && g->nranks > 1) return fail(HZSDR_ERR_INVALID, "hzsdr_beam_group_exec: call hzsdr_beam_group_connect first");
    if (nchan < 0 || nchan > kMaxBeamChans || (nchan && (!chans || !weights)))
        return fail(HZSDR_ERR_INVALID, "hzsdr_beam_group_exec: 0..%d channels per rank", kMaxBeamChans);
    if (src_format != HZSDR_FORMAT_U8 && src_format != HZSDR_FORMAT_I8 && src_format != HZSDR_FORMAT_I16)
        return fail(HZSDR_ERR_FORMAT_UNKNOWN, "hzsdr_beam_group_exec: raw source format expected, got %d", src_format);
    if (!dst_slice || ((uintptr_t)dst_slice & 15)) return fail(HZSDR_ERR_INVALID, "hzsdr_beam_group_exec: dst must be 16-byte aligned");
    const int sb = hzsdr_format_size(src_format);
    BeamArgs a;
    a.nchan = nchan;
    a.accumulate = 0;
    const float ws = beam_weight_scale(src_format);
    for (int c = 0; c < nchan; c++) {
        if (!chans[c] || ((uintptr_t)chans[c] % (4 * sb))) return fail(HZSDR_ERR_INVALID, "hzsdr_beam_group_exec: channel %d misaligned", c);
        a.chan[c] = (const uint8_t *)chans[c];
        a.w[c] = make_float2(weights[2 * c] * ws, weights[2 * c + 1] * ws);
    }
    g->step++;
    const size_t set_off = kFlagBytes + (size_t)(g->step & 1u) * g->set_bytes;
    GroupArgs ga;
    ga.nranks = g->nranks;
    ga.step = g->step;
    ga.done_counter = g->done_counter;
    ga.quads_per_slice = (uint32_t)(g->slice / 4);
    for (int s = 0; s < g->nranks; s++) {
        ga.slot[s] = (float4 *)(g->peer[s] + set_off + (size_t)g->rank * g->slot_bytes);
        ga.flag[s] = (uint32_t *)(g->peer[s]) + (size_t)g->rank * 32;
    }
    const size_t nquads = g->n / 4;
    const int grid = (int)std::min<size_t>((nquads + 255) / 256, (size_t)g->ctx->sm_count * 8);
    cudaStream_t st = g->ctx->stream;
    const int par = (int)(g->step & 1u);
    // this step overwrites the staging set last used two steps ago: its finishing kernel must be done
    if (g->fin_used[par]) HZ_CUDA(cudaStreamWaitEvent(st, g->finished[par], 0));
    switch (src_format) {
        case HZSDR_FORMAT_U8: k_beamform_rs<HZSDR_FORMAT_U8><<<grid, 256, 0, st>>>(nquads, a, ga); break;
        case HZSDR_FORMAT_I8: k_beamform_rs<HZSDR_FORMAT_I8><<<grid, 256, 0, st>>>(nquads, a, ga); break;
        default: k_beamform_rs<HZSDR_FORMAT_I16><<<grid, 256, 0, st>>>(nquads, a, ga); break;
    }
    HZ_CHECK_LAUNCH();
    HZ_CUDA(cudaEventRecord(g->computed, st));
    HZ_CUDA(cudaStreamWaitEvent(g->fin_stream, g->computed, 0));
    const size_t nvec = g->slice / 2;
    const int fgrid = (int)std::min<size_t>((nvec + 255) / 256, (size_t)g->ctx->sm_count * 2);
    k_beam_finish<<<fgrid, 256, 0, g->fin_stream>>>((const float4 *)(g->base + set_off), g->slot_bytes / 16, (const uint32_t *)g->base,
                                                     (float4 *)dst_slice, nvec, g->nranks, g->step);
    HZ_CHECK_LAUNCH();
    HZ_CUDA(cudaEventRecord(g->finished[par], g->fin_stream));
    g->fin_used[par] = true;
    return HZSDR_OK;
}

// Make the context's stream wait (on the device, not the host) for every finishing kernel enqueued
// so far: after this, work enqueued on the context stream -- or hzsdr_ctx_sync -- sees the slices.
extern "C" int hzsdr_beam_group_join(hzsdr_beam_group *g) {
    if (!g) return fail(HZSDR_ERR_INVALID, "hzsdr_beam_group_join: null");
    HZ_ENTER(g->ctx);
    for (int p = 0; p < 2; p++)
        if (g->fin_used[p]) HZ_CUDA(cudaStreamWaitEvent(g->ctx->stream, g->finished[p], 0));
    return HZSDR_OK;
}
