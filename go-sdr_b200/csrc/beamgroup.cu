// beamgroup.cu -- multi-GPU Beamform with the reduction fused into the kernel over NVLink peer
// memory (one process per GPU, CUDA IPC).
//
// Channels are sharded across G ranks (hzsdr_shard.channel_shard).  Every rank computes the partial
// beam of ITS channels for ALL samples, but it never stores that partial beam locally: output
// sample slice s (n/G samples of every buffer) belongs to rank s, and the kernel writes its partial
// sums for slice s straight into rank s's staging slot over NVLink, tile by tile, while the next
// tiles are still being computed.  That is a reduce-scatter whose transfer overlaps the math: each
// GPU sends and receives (G-1)/G * 8 B per output sample instead of funnelling (G-1) * 8 B into one
// root.
//
//   batch    One exchange covers `nbuf` buffers (hzsdr_beam_group_exec_batch): one launch, one flag
//            round and one finishing launch per nbuf * n samples.  At 8 GPUs a single 2^20-sample
//            buffer is ~4 us of HBM reads and ~10 us of NVLink against ~35 us of launch + flag latency.
//   rotation Rank r walks the slices in the order r+1, r+2, ..., r (mod G), every slice across the whole
//            batch before the next.  At any moment the G ranks therefore write to G DIFFERENT owners:
//            every NVLink port receives from one peer and sends to one peer.  (In natural order all
//            ranks write owner 0 first -- a 7:1 incast on one port while the other ports idle.)
//   publish  When the kernel's last CTA has fenced its stores (system scope) it raises this rank's flag
//            (= the step number) on every peer.
//   finish   A small kernel on a side stream waits for the G flags of the step and sums the G slots of
//            this rank's slice in rank order (deterministic; inside a shard the channel order is the
//            reference's left-to-right, stream/add.go:115-119).  Its last CTA then writes an ACK
//            (= the step number) to every peer: "I have finished reading my staging set of this step".
//   reuse    Staging is a ring of kSets = 3 sets.  Step s+3 overwrites the set of step s on every
//            peer, so the compute kernel of step s+3 spins -- before its first store -- until every
//            peer's ack shows >= s.  (With two sets the wait sat on the critical path: a step could not start
//            storing before every peer's finishing kernel of the step before last had run, and that kernel shares
//            its GPU with the next compute kernel -- 70 us of fixed cost per exchange at 8 GPUs.  The finishing
//            kernel also runs on a high-priority stream, and the compute grid leaves it SM slots.)  (The earlier version inferred that from local events, which only
//            prove that the peers have COMPUTED step s, not that their finishing kernels -- on side
//            streams, possibly delayed by the next compute kernel -- have read it.)
#include <cstdlib>
#include <vector>

#include "beam.cuh"
#include "common.cuh"

namespace hz {

constexpr int kMaxRanks = 16;
constexpr int kSets = 3;              // staging sets (steps in flight)
constexpr int kMaxGroupBatch = 64;    // buffers per exchange
constexpr int kMaxGroupPtrs = 1024;   // nbuf * (channels of this rank) raw-buffer pointers per launch
constexpr size_t kFlagBytes = 4096;   // [flags: kMaxRanks x 128 B | acks: kMaxRanks x 128 B]
constexpr size_t kAckOffset = 2048;

struct GroupArgs {
    const uint8_t *chan[kMaxGroupPtrs];  // [k * nchan + c]: channel c of buffer k
    float2 w[kMaxBeamChans];             // weights x the format's conversion scale
    float4 *slot[kMaxRanks];             // slot[s]: this rank's staging region on rank s, current parity
    uint32_t *flag[kMaxRanks];           // flag[s]: this rank's flag word on rank s
    const uint32_t *ack;                 // local: ack[32 r] = last step whose set rank r has finished reading
    unsigned int *done_counter;          // local CTA counter
    uint32_t quads_per_slice;            // n / G / 4
    uint32_t nbuf, nchan;
    uint32_t step, need_ack;             // need_ack: every ack must show >= this before the first store (0: none)
    int nranks, rank;
};

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Q2: two quads per thread in flight, for ranks that hold <= 8 channels (74 registers, 3 CTAs per SM); otherwise one
// quad with 16 channel loads in flight (44 registers, 5 CTAs per SM).  Resident warps matter here: a warp stalls on its
// NVLink stores, and only other warps keep the HBM reads going (2 / 3 CTAs per SM at 2 GPUs: 3.0 / 4.0 T channel-samples/s).
template <int FMT, bool Q2>
__global__ void __launch_bounds__(256, Q2 ? 3 : 5) k_beamform_rs(const __grid_constant__ GroupArgs g) {
    // Peer stores must be full lines: a warp's 32 quads (64 float4 = 1 KB, contiguous in the owner's
    // slot because slices are multiples of 128 samples) are transposed through shared memory so that
    // each of the two store instructions writes 512 contiguous bytes over NVLink.
    __shared__ float4 stage[8][64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (g.need_ack) {  // the staging set this step overwrites has been read by every owner
        if ((int)threadIdx.x < g.nranks)
            while ((int32_t)(ld_volatile_u32(g.ack + threadIdx.x * 32) - g.need_ack) < 0) __nanosleep(64);
        __syncthreads();
        __threadfence_system();
    }
    // (all indices fit 32 bits: hzsdr_beam_group_create bounds n / 4 * max_batch)
    const uint32_t per_slice = g.nbuf * g.quads_per_slice;  // quads of one owner's slice over the whole batch
    const uint32_t total = per_slice * (uint32_t)g.nranks;
    const uint32_t stride = gridDim.x * blockDim.x;
    auto locate = [&](uint32_t i, uint32_t &owner, uint32_t &k, uint32_t &q) {
        const uint32_t j = i / per_slice;           // position in this rank's rotated slice order
        const uint32_t rem = i - j * per_slice;
        k = rem / g.quads_per_slice;                // buffer
        q = rem - k * g.quads_per_slice;
        owner = (uint32_t)g.rank + 1u + j;
        if (owner >= (uint32_t)g.nranks) owner -= (uint32_t)g.nranks;
    };
    // a warp's 32 quads -> the owner's region (buffer-major, (k, q - lane)): two 512-byte stores
    auto ship = [&](const float (&acc)[8], uint32_t owner, uint32_t k, uint32_t q) {
        stage[warp][2 * lane] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        stage[warp][2 * lane + 1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
        __syncwarp();
        float4 *dst = g.slot[owner] + 2 * ((size_t)k * g.quads_per_slice + (q - lane));  // peer memory unless owner == this rank
        dst[lane] = stage[warp][lane];
        dst[32 + lane] = stage[warp][32 + lane];
        __syncwarp();
    };
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    // (total is a multiple of 32: whole warps iterate together)
    if constexpr (Q2) {  // few channels per rank: two quads per thread in flight (beam.cuh, beam_quad2)
        for (; i + stride < total; i += 2u * stride) {
            uint32_t oa, ka, qa, ob, kb, qb;
            locate(i, oa, ka, qa);
            locate(i + stride, ob, kb, qb);
            float acc_a[8], acc_b[8];
#pragma unroll
            for (int u = 0; u < 8; u++) acc_a[u] = 0.f, acc_b[u] = 0.f;
            beam_quad2<FMT>(g.chan + (size_t)ka * g.nchan, g.chan + (size_t)kb * g.nchan, g.w, (int)g.nchan,
                            (size_t)oa * g.quads_per_slice + qa, (size_t)ob * g.quads_per_slice + qb, acc_a, acc_b);
            ship(acc_a, oa, ka, qa);
            ship(acc_b, ob, kb, qb);
        }
    }
    for (; i < total; i += stride) {
        uint32_t owner, k, q;
        locate(i, owner, k, q);
        float acc[8];
#pragma unroll
        for (int u = 0; u < 8; u++) acc[u] = 0.f;
        beam_quad<FMT>(g.chan + (size_t)k * g.nchan, g.w, (int)g.nchan, (size_t)owner * g.quads_per_slice + q, acc);
        ship(acc, owner, k, q);
    }
    // publish: all of this CTA's stores first, then (last CTA only) the flags on every rank.  One system-scope
    // fence per CTA, by the thread that then counts the CTA in -- behind the CTA barrier it covers every thread's
    // stores (fences are cumulative).  A fence in every thread was a third of this kernel's stall samples.
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) {
        __threadfence_system();
        last = atomicAdd(g.done_counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last) {
        __threadfence_system();
        if ((int)threadIdx.x < g.nranks) {
            volatile uint32_t *f = g.flag[threadIdx.x];
            *f = g.step;
        }
        if (threadIdx.x == 0) *g.done_counter = 0;
    }
}

struct FinishArgs {
    float4 *out[kMaxGroupBatch];   // out[k]: this rank's slice of buffer k
    uint32_t *ack[kMaxRanks];      // ack[s]: this rank's ack word on rank s
    const float4 *slots;           // local staging set of the step: [source rank][buffer][slice]
    const uint32_t *flags;         // local: flags[32 r] = last step rank r has published
    unsigned int *done_counter;
    size_t slot_stride_vec;        // float4 per source rank region
    uint32_t vec_per_slice;        // n / G / 2
    uint32_t nbuf, step;
    int nranks;
};

// out[k][i] = sum over ranks (in rank order) of slot_r[k][i], after every rank's flag shows `step`
__global__ void __launch_bounds__(256) k_beam_finish(const __grid_constant__ FinishArgs f) {
    if ((int)threadIdx.x < f.nranks)
        while ((int32_t)(ld_volatile_u32(f.flags + threadIdx.x * 32) - f.step) < 0) __nanosleep(100);
    __syncthreads();
    __threadfence_system();
    const size_t nvec = (size_t)f.nbuf * f.vec_per_slice;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        float4 acc = __ldcv(f.slots + i);  // peers wrote these lines: bypass any stale cached copy
        for (int r = 1; r < f.nranks; r++) {
            const float4 v = __ldcv(f.slots + (size_t)r * f.slot_stride_vec + i);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        const uint32_t k = (uint32_t)(i / f.vec_per_slice);
        st_stream_f4(f.out[k] + (i - (size_t)k * f.vec_per_slice), acc);
    }
    // acknowledge: every CTA's loads have returned (their values were consumed above); the last CTA tells
    // every rank that this staging set may be overwritten
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) last = atomicAdd(f.done_counter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (last) {
        __threadfence_system();
        if ((int)threadIdx.x < f.nranks) {
            volatile uint32_t *a = f.ack[threadIdx.x];
            *a = f.step;
        }
        if (threadIdx.x == 0) *f.done_counter = 0;
    }
}

}  // namespace hz

using namespace hz;

struct hzsdr_beam_group {
    hzsdr_ctx *ctx = nullptr;
    int nranks = 0, rank = 0;
    size_t n = 0, slice = 0;     // samples per buffer, per slice
    size_t max_batch = 1;        // buffers per exchange the staging is sized for
    size_t slot_bytes = 0;       // one source rank's region of a set: max_batch * slice * 8 B, 256-aligned
    size_t set_bytes = 0;        // nranks regions
    uint8_t *base = nullptr;     // local allocation: [flags + acks 4 KB | kSets staging sets]
    unsigned int *counters = nullptr;  // [0]: compute kernel, [1]: finishing kernel
    uint8_t *peer[kMaxRanks] = {};     // every rank's base as seen from here (peer[rank] == base)
    bool connected = false;
    uint32_t step = 0;
    // the finishing kernel spins on the peers' flags: it runs on a side stream so that the next
    // batch's compute does not queue behind the wait
    cudaStream_t fin_stream = nullptr;
    cudaEvent_t computed = nullptr;      // this step's k_beamform_rs is done
    cudaEvent_t finished[kSets] = {};    // finishing kernel of the latest step on this staging set is done
    bool fin_used[kSets] = {};
};

extern "C" int hzsdr_beam_group_destroy(hzsdr_beam_group *g) {
    if (!g) return HZSDR_OK;
    HZ_ENTER(g->ctx);
    cudaStreamSynchronize(g->ctx->stream);
    if (g->fin_stream) {
        cudaStreamSynchronize(g->fin_stream);
        cudaStreamDestroy(g->fin_stream);
    }
    for (int r = 0; r < g->nranks; r++)
        if (g->connected && r != g->rank && g->peer[r]) cudaIpcCloseMemHandle(g->peer[r]);
    if (g->computed) cudaEventDestroy(g->computed);
    for (auto ev : g->finished)
        if (ev) cudaEventDestroy(ev);
    if (g->base) cudaFree(g->base);
    if (g->counters) cudaFree(g->counters);
    delete g;
    return HZSDR_OK;
}

extern "C" int hzsdr_beam_group_create(hzsdr_ctx *ctx, int nranks, int rank, size_t n, size_t max_batch, void *handle_out,
                                       hzsdr_beam_group **out) {
    HZ_ENTER(ctx);
    if (!out || !handle_out || nranks < 1 || nranks > kMaxRanks || rank < 0 || rank >= nranks)
        return fail(HZSDR_ERR_INVALID, "hzsdr_beam_group_create: bad arguments");
    *out = nullptr;
    if (n == 0 || n % ((size_t)nranks * 128)) return fail(HZSDR_ERR_INVALID, "hzsdr_beam_group_create: n must be a multiple of 128 * nranks");
    if (max_batch < 1 || max_batch > (size_t)kMaxGroupBatch)
        return fail(HZSDR_ERR_INVALID, "hzsdr_beam_group_create: max_batch must be 1..%d", kMaxGroupBatch);
    if (n / 4 * max_batch > 0x7fffffffull) return fail(HZSDR_ERR_INVALID, "hzsdr_beam_group_create: batch too large");
    static_assert(sizeof(cudaIpcMemHandle_t) == HZSDR_IPC_HANDLE_BYTES, "IPC handle size");
    hzsdr_beam_group *g = new hzsdr_beam_group();
    g->ctx = ctx;
    g->nranks = nranks;
    g->rank = rank;
    g->n = n;
    g->slice = n / nranks;
    g->max_batch = max_batch;
    g->slot_bytes = (g->slice * 8 * max_batch + 255) / 256 * 256;
    g->set_bytes = g->slot_bytes * nranks;
    const size_t total = kFlagBytes + (size_t)kSets * g->set_bytes;
    cudaError_t e = cudaMalloc((void **)&g->base, total);
    if (e == cudaSuccess) e = cudaMemset(g->base, 0, total);
    if (e == cudaSuccess) e = cudaMalloc((void **)&g->counters, 2 * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemset(g->counters, 0, 2 * sizeof(unsigned int));
    if (e == cudaSuccess) {
        int lo = 0, hi = 0;  // (numerically lowest = greatest priority)
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        e = cudaStreamCreateWithPriority(&g->fin_stream, cudaStreamNonBlocking, hi);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g->computed, cudaEventDisableTiming);
    for (int k = 0; k < kSets && e == cudaSuccess; k++) e = cudaEventCreateWithFlags(&g->finished[k], cudaEventDisableTiming);
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

extern "C" int hzsdr_beam_group_exec_batch(hzsdr_beam_group *g, int src_format, const void *const *chans, int nchan,
                                           const float *weights, size_t nbuf, void *const *dst_slices) {
    if (!g) return fail(HZSDR_ERR_INVALID, "hzsdr_beam_group_exec_batch: null");
    HZ_ENTER(g->ctx);
    if (!g->connected && g->nranks > 1) return fail(HZSDR_ERR_INVALID, "hzsdr_beam_group_exec_batch: call hzsdr_beam_group_connect first");
    if (nchan < 0 || nchan > kMaxBeamChans || (nchan && (!chans || !weights)))
        return fail(HZSDR_ERR_INVALID, "hzsdr_beam_group_exec_batch: 0..%d channels per rank", kMaxBeamChans);
    if (nbuf < 1 || nbuf > g->max_batch || !dst_slices)
        return fail(HZSDR_ERR_INVALID, "hzsdr_beam_group_exec_batch: 1..%zu buffers per exchange (max_batch of the group)", g->max_batch);
    if (nbuf * (size_t)nchan > (size_t)kMaxGroupPtrs)
        return fail(HZSDR_ERR_INVALID, "hzsdr_beam_group_exec_batch: nbuf * nchan = %zu exceeds %d", nbuf * (size_t)nchan, kMaxGroupPtrs);
    if (src_format != HZSDR_FORMAT_U8 && src_format != HZSDR_FORMAT_I8 && src_format != HZSDR_FORMAT_I16)
        return fail(HZSDR_ERR_FORMAT_UNKNOWN, "hzsdr_beam_group_exec_batch: raw source format expected, got %d", src_format);
    const int sb = hzsdr_format_size(src_format);
    GroupArgs ga;
    FinishArgs fa;
    const float ws = beam_weight_scale(src_format);
    for (int c = 0; c < nchan; c++) ga.w[c] = make_float2(weights[2 * c] * ws, weights[2 * c + 1] * ws);
    for (size_t k = 0; k < nbuf; k++) {
        if (!dst_slices[k] || ((uintptr_t)dst_slices[k] & 15)) return fail(HZSDR_ERR_INVALID, "hzsdr_beam_group_exec_batch: dst %zu must be 16-byte aligned", k);
        fa.out[k] = (float4 *)dst_slices[k];
        for (int c = 0; c < nchan; c++) {
            const void *p = chans[k * (size_t)nchan + c];
            if (!p || ((uintptr_t)p % (4 * sb))) return fail(HZSDR_ERR_INVALID, "hzsdr_beam_group_exec_batch: buffer %zu channel %d misaligned", k, c);
            ga.chan[k * (size_t)nchan + c] = (const uint8_t *)p;
        }
    }
    g->step++;
    const int par = (int)(g->step % (uint32_t)kSets);
    const size_t set_off = kFlagBytes + (size_t)par * g->set_bytes;
    ga.nranks = g->nranks;
    ga.rank = g->rank;
    ga.step = g->step;
    ga.need_ack = g->step > (uint32_t)kSets ? g->step - (uint32_t)kSets : 0;
    ga.nbuf = (uint32_t)nbuf;
    ga.nchan = (uint32_t)nchan;
    ga.done_counter = g->counters;
    ga.ack = (const uint32_t *)(g->base + kAckOffset);
    ga.quads_per_slice = (uint32_t)(g->slice / 4);
    for (int s = 0; s < g->nranks; s++) {
        ga.slot[s] = (float4 *)(g->peer[s] + set_off + (size_t)g->rank * g->slot_bytes);
        ga.flag[s] = (uint32_t *)(g->peer[s]) + (size_t)g->rank * 32;
        fa.ack[s] = (uint32_t *)(g->peer[s] + kAckOffset) + (size_t)g->rank * 32;
    }
    const size_t nquads = g->n / 4 * nbuf;
    // one wave of every CTA the SM can hold (the finishing kernel of the previous step, on its high-priority stream,
    // takes slots as they come free; three staging sets give it the slack)
    static const int q2_env = [] { const char *e = getenv("HZSDR_BEAM_Q2"); return e ? atoi(e) : -1; }();  // (experiments: force either form)
    const bool q2 = q2_env >= 0 ? (q2_env != 0 && nchan <= 8) : nchan <= 8;
    const int grid = (int)std::min<size_t>((nquads + 255) / 256, (size_t)g->ctx->sm_count * (q2 ? 3 : 5));
    cudaStream_t st = g->ctx->stream;
    g->ctx->overlap_broken();  // a kernel outside the overlap scheme
    switch (src_format) {
        case HZSDR_FORMAT_U8:
            if (q2) k_beamform_rs<HZSDR_FORMAT_U8, true><<<grid, 256, 0, st>>>(ga); else k_beamform_rs<HZSDR_FORMAT_U8, false><<<grid, 256, 0, st>>>(ga);
            break;
        case HZSDR_FORMAT_I8:
            if (q2) k_beamform_rs<HZSDR_FORMAT_I8, true><<<grid, 256, 0, st>>>(ga); else k_beamform_rs<HZSDR_FORMAT_I8, false><<<grid, 256, 0, st>>>(ga);
            break;
        default:
            if (q2) k_beamform_rs<HZSDR_FORMAT_I16, true><<<grid, 256, 0, st>>>(ga); else k_beamform_rs<HZSDR_FORMAT_I16, false><<<grid, 256, 0, st>>>(ga);
            break;
    }
    HZ_CHECK_LAUNCH();
    HZ_CUDA(cudaEventRecord(g->computed, st));
    HZ_CUDA(cudaStreamWaitEvent(g->fin_stream, g->computed, 0));
    fa.slots = (const float4 *)(g->base + set_off);
    fa.flags = (const uint32_t *)g->base;
    fa.done_counter = g->counters + 1;
    fa.slot_stride_vec = g->slot_bytes / 16;
    fa.vec_per_slice = (uint32_t)(g->slice / 2);
    fa.nbuf = (uint32_t)nbuf;
    fa.step = g->step;
    fa.nranks = g->nranks;
    const size_t nvec = g->slice / 2 * nbuf;
    const int fgrid = (int)std::min<size_t>((nvec + 255) / 256, (size_t)g->ctx->sm_count);
    k_beam_finish<<<fgrid, 256, 0, g->fin_stream>>>(fa);
    HZ_CHECK_LAUNCH();
    HZ_CUDA(cudaEventRecord(g->finished[par], g->fin_stream));
    g->fin_used[par] = true;
    return HZSDR_OK;
}

extern "C" int hzsdr_beam_group_exec(hzsdr_beam_group *g, int src_format, const void *const *chans, int nchan,
                                     const float *weights, void *dst_slice) {
    void *const dst[1] = {dst_slice};
    return hzsdr_beam_group_exec_batch(g, src_format, chans, nchan, weights, 1, dst);
}

// Make the context's stream wait (on the device, not the host) for every finishing kernel enqueued
// so far: after this, work enqueued on the context stream -- or hzsdr_ctx_sync -- sees the slices.
extern "C" int hzsdr_beam_group_join(hzsdr_beam_group *g) {
    if (!g) return fail(HZSDR_ERR_INVALID, "hzsdr_beam_group_join: null");
    HZ_ENTER(g->ctx);
    for (int p = 0; p < kSets; p++)
        if (g->fin_used[p]) HZ_CUDA(cudaStreamWaitEvent(g->ctx->stream, g->finished[p], 0));
    return HZSDR_OK;
}
