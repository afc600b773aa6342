// api.cu -- context, memory, pinned ring, NCCL communicator and host-side helpers of the C ABI.
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>

#include <mutex>
#include <vector>

#include "common.cuh"

namespace hz {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int fail(int status, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return status;
}

}  // namespace hz

using namespace hz;

extern "C" const char *hzsdr_last_error(void) { return g_err; }

extern "C" const char *hzsdr_version(void) { return "hzsdrcuda 0.1.0 (sm_100a)"; }

extern "C" int hzsdr_format_size(int format) {
    switch (format) {  // iq.go:99-110
        case HZSDR_FORMAT_U8:
        case HZSDR_FORMAT_I8: return 2;
        case HZSDR_FORMAT_I16: return 4;
        case HZSDR_FORMAT_C64: return 8;
        default: return 0;
    }
}

extern "C" int hzsdr_device_count(int *count) {
    if (!count) return fail(HZSDR_ERR_INVALID, "hzsdr_device_count: null");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        cudaGetLastError();
        return fail(HZSDR_ERR_NO_DEVICE, "hzsdr_device_count: %s", cudaGetErrorString(e));
    }
    *count = n;
    return HZSDR_OK;
}

extern "C" int hzsdr_ctx_create(int device, hzsdr_ctx **out) {
    if (!out) return fail(HZSDR_ERR_INVALID, "hzsdr_ctx_create: null out");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(HZSDR_ERR_NO_DEVICE,
                    "hzsdr_ctx_create: no CUDA device (%s); libhzsdrcuda has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (device < 0 || device >= n) return fail(HZSDR_ERR_INVALID, "hzsdr_ctx_create: device %d of %d", device, n);
    hzsdr_ctx *ctx = new hzsdr_ctx();
    ctx->device = device;
    if ((e = cudaGetDeviceProperties(&ctx->prop, device)) != cudaSuccess) {
        delete ctx;
        return fail(HZSDR_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    }
    if (ctx->prop.major != 10) {
        int maj = ctx->prop.major, min = ctx->prop.minor;
        delete ctx;
        return fail(HZSDR_ERR_NO_DEVICE, "hzsdr_ctx_create: device %d is sm_%d%d; this library carries sm_100a code only",
                    device, maj, min);
    }
    ctx->sm_count = ctx->prop.multiProcessorCount;
    if ((e = cudaSetDevice(device)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        delete ctx;
        return fail(HZSDR_ERR_CUDA, "hzsdr_ctx_create: %s", cudaGetErrorString(e));
    }
    if ((e = cudaMalloc(&ctx->overlap_done, OverlapWindow::kSlots * sizeof(uint32_t))) != cudaSuccess ||
        (e = cudaMemset(ctx->overlap_done, 0, OverlapWindow::kSlots * sizeof(uint32_t))) != cudaSuccess) {
        cudaStreamDestroy(ctx->stream);
        delete ctx;
        return fail(HZSDR_ERR_CUDA, "hzsdr_ctx_create: %s", cudaGetErrorString(e));
    }
    *out = ctx;
    return HZSDR_OK;
}

namespace hz {
int ctx_workspace(hzsdr_ctx *ctx, size_t bytes, void **out) {
    if (bytes > ctx->workspace_bytes) {
        if (ctx->workspace) HZ_CUDA(cudaFree(ctx->workspace));  // cudaFree waits for the kernels still using it
        ctx->workspace = nullptr;
        ctx->workspace_bytes = 0;
        const size_t want = (bytes + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);
        HZ_CUDA(cudaMalloc(&ctx->workspace, want));
        ctx->workspace_bytes = want;
    }
    *out = ctx->workspace;
    return HZSDR_OK;
}
}  // namespace hz

extern "C" int hzsdr_ctx_destroy(hzsdr_ctx *ctx) {
    if (!ctx) return HZSDR_OK;
    HZ_ENTER(ctx);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->workspace) cudaFree(ctx->workspace);
    if (ctx->overlap_done) cudaFree(ctx->overlap_done);
    ctx->host_pipe.destroy();
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return HZSDR_OK;
}

extern "C" int hzsdr_ctx_sync(hzsdr_ctx *ctx) {
    HZ_ENTER(ctx);
    HZ_CUDA(cudaStreamSynchronize(ctx->stream));
    return HZSDR_OK;
}

// completes everything the *_submit_host entry points of this context have enqueued (their copy
// streams included); after it the destination host buffers hold the results
extern "C" int hzsdr_ctx_wait_host(hzsdr_ctx *ctx) {
    HZ_ENTER(ctx);
    HZ_CUDA(ctx->host_pipe.drain(ctx->stream));
    return HZSDR_OK;
}

extern "C" int hzsdr_ctx_stream(hzsdr_ctx *ctx, void **cuda_stream) {
    if (!ctx || !cuda_stream) return fail(HZSDR_ERR_INVALID, "hzsdr_ctx_stream: null");
    *cuda_stream = (void *)ctx->stream;
    return HZSDR_OK;
}

extern "C" int hzsdr_ctx_info(hzsdr_ctx *ctx, char *name, size_t name_len, int *sm_major, int *sm_minor,
                              int *sm_count, size_t *hbm_bytes) {
    if (!ctx) return fail(HZSDR_ERR_INVALID, "hzsdr_ctx_info: null context");
    if (name && name_len) {
        strncpy(name, ctx->prop.name, name_len - 1);
        name[name_len - 1] = 0;
    }
    if (sm_major) *sm_major = ctx->prop.major;
    if (sm_minor) *sm_minor = ctx->prop.minor;
    if (sm_count) *sm_count = ctx->sm_count;
    if (hbm_bytes) *hbm_bytes = ctx->prop.totalGlobalMem;
    return HZSDR_OK;
}

// ---- memory ------------------------------------------------------------------------------------
extern "C" int hzsdr_dev_alloc(hzsdr_ctx *ctx, size_t bytes, void **out) {
    HZ_ENTER(ctx);
    if (!out) return fail(HZSDR_ERR_INVALID, "hzsdr_dev_alloc: null out");
    *out = nullptr;
    cudaError_t e = cudaMalloc(out, bytes ? bytes : 1);
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        return fail(HZSDR_ERR_NOMEM, "hzsdr_dev_alloc: %zu bytes: out of device memory", bytes);
    }
    HZ_CUDA(e);
    return HZSDR_OK;
}

extern "C" int hzsdr_dev_free(hzsdr_ctx *ctx, void *dev) {
    HZ_ENTER(ctx);
    if (dev) HZ_CUDA(cudaFree(dev));
    return HZSDR_OK;
}

extern "C" int hzsdr_dev_memset(hzsdr_ctx *ctx, void *dev, int value, size_t bytes) {
    HZ_ENTER(ctx);
    if (bytes) HZ_CUDA(cudaMemsetAsync(dev, value, bytes, ctx->stream));
    return HZSDR_OK;
}

extern "C" int hzsdr_pinned_alloc(size_t bytes, void **out) {
    if (!out) return fail(HZSDR_ERR_INVALID, "hzsdr_pinned_alloc: null out");
    *out = nullptr;
    cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(e == cudaErrorMemoryAllocation ? HZSDR_ERR_NOMEM : HZSDR_ERR_CUDA, "hzsdr_pinned_alloc: %s",
                    cudaGetErrorString(e));
    }
    return HZSDR_OK;
}

extern "C" int hzsdr_pinned_free(void *host) {
    if (host) HZ_CUDA(cudaFreeHost(host));
    return HZSDR_OK;
}

extern "C" int hzsdr_upload(hzsdr_ctx *ctx, void *dst, const void *src, size_t bytes) {
    HZ_ENTER(ctx);
    if (bytes) HZ_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return HZSDR_OK;
}

extern "C" int hzsdr_download(hzsdr_ctx *ctx, void *dst, const void *src, size_t bytes) {
    HZ_ENTER(ctx);
    if (bytes) HZ_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    HZ_CUDA(cudaStreamSynchronize(ctx->stream));
    return HZSDR_OK;
}

extern "C" int hzsdr_copy(hzsdr_ctx *ctx, void *dst, const void *src, size_t bytes) {
    HZ_ENTER(ctx);
    if (bytes) HZ_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return HZSDR_OK;
}

// ---- stream.BeamformAngles2D, stream/beamform.go:57-107 (fp64 host math) ----------------------
extern "C" int hzsdr_beamform_angles_2d(double frequency_hz, double angle_deg, const double center[2],
                                        const double *antennas, int n, float *out) {
    if (n < 0 || (n > 0 && (!antennas || !out || !center)))
        return fail(HZSDR_ERR_INVALID, "hzsdr_beamform_angles_2d: bad arguments");
    const double wavelength = 299792458.0 / frequency_hz;  // rf.Hz.Wavelength (hz.tools/rf v0.0.7)
    for (int i = 0; i < n; i++) {
        const double xd = antennas[2 * i] - center[0];
        const double yd = antennas[2 * i + 1] - center[1];
        const double dist = sqrt(xd * xd + yd * yd);
        if (dist == 0) {  // beamform.go:80-83
            out[2 * i] = 1.f;
            out[2 * i + 1] = 0.f;
            continue;
        }
        const double angle_r = angle_deg * (M_PI / 180);
        const double n_theta = asin(yd / dist);
        const double p_opposite = sin(n_theta + angle_r) * dist;
        const double phase_shift = (p_opposite / wavelength) * 360;
        const double phase_r = phase_shift * (M_PI / 180);
        out[2 * i] = (float)cos(phase_r);   // conj(cos + i sin), beamform.go:100-103
        out[2 * i + 1] = (float)(-sin(phase_r));
    }
    return HZSDR_OK;
}

// =================================================================================================
// Pinned-host slot ring with async H2D (stream/ring.go:48-392; the allocator hook at :60-64)
// =================================================================================================
struct hzsdr_ring {
    hzsdr_ctx *ctx = nullptr;
    int format = 0;
    size_t slots = 0, slot_len = 0, slot_bytes = 0;
    uint8_t *host = nullptr;  // slots * slot_bytes, cudaHostAlloc
    uint8_t *dev = nullptr;   // slots * slot_bytes
    cudaStream_t copy_stream = nullptr;
    std::vector<cudaEvent_t> landed;   // H2D of slot i complete
    std::vector<cudaEvent_t> consumed; // compute on slot i's device copy complete
    std::vector<size_t> fill;          // samples written per slot
    std::vector<char> consumed_valid;
    std::vector<char> landed_valid;    // an H2D copy out of host slot i has been enqueued
    size_t widx = 0, ridx = 0, pending = 0;
    bool reading = false;
    std::mutex mu;
};

extern "C" int hzsdr_ring_create(hzsdr_ctx *ctx, int format, size_t slots, size_t slot_len, hzsdr_ring **out) {
    HZ_ENTER(ctx);
    if (!out) return fail(HZSDR_ERR_INVALID, "hzsdr_ring_create: null out");
    *out = nullptr;
    const int sb = hzsdr_format_size(format);
    if (!sb) return fail(HZSDR_ERR_FORMAT_UNKNOWN, "hzsdr_ring_create: unknown format %d", format);
    if (slots < 2 || slot_len == 0) return fail(HZSDR_ERR_INVALID, "hzsdr_ring_create: need >= 2 slots");
    hzsdr_ring *r = new hzsdr_ring();
    r->ctx = ctx;
    r->format = format;
    r->slots = slots;
    r->slot_len = slot_len;
    r->slot_bytes = ((slot_len * sb + 255) / 256) * 256;  // keep every slot 256-byte aligned
    auto bail = [&](cudaError_t e, const char *what) {
        fail(HZSDR_ERR_CUDA, "hzsdr_ring_create: %s: %s", what, cudaGetErrorString(e));
        hzsdr_ring_destroy(r);
        return HZSDR_ERR_CUDA;
    };
    cudaError_t e;
    if ((e = cudaHostAlloc((void **)&r->host, slots * r->slot_bytes, cudaHostAllocPortable)) != cudaSuccess)
        return bail(e, "cudaHostAlloc");
    if ((e = cudaMalloc((void **)&r->dev, slots * r->slot_bytes)) != cudaSuccess) return bail(e, "cudaMalloc");
    if ((e = cudaStreamCreateWithFlags(&r->copy_stream, cudaStreamNonBlocking)) != cudaSuccess)
        return bail(e, "cudaStreamCreate");
    r->landed.resize(slots);
    r->consumed.resize(slots);
    r->fill.assign(slots, 0);
    r->consumed_valid.assign(slots, 0);
    r->landed_valid.assign(slots, 0);
    for (size_t i = 0; i < slots; i++) {
        if ((e = cudaEventCreateWithFlags(&r->landed[i], cudaEventDisableTiming)) != cudaSuccess) return bail(e, "event");
        if ((e = cudaEventCreateWithFlags(&r->consumed[i], cudaEventDisableTiming)) != cudaSuccess) return bail(e, "event");
    }
    *out = r;
    return HZSDR_OK;
}

extern "C" int hzsdr_ring_destroy(hzsdr_ring *r) {
    if (!r) return HZSDR_OK;
    if (r->ctx) cudaSetDevice(r->ctx->device);
    if (r->copy_stream) {
        cudaStreamSynchronize(r->copy_stream);
        cudaStreamDestroy(r->copy_stream);
    }
    for (auto ev : r->landed)
        if (ev) cudaEventDestroy(ev);
    for (auto ev : r->consumed)
        if (ev) cudaEventDestroy(ev);
    if (r->host) cudaFreeHost(r->host);
    if (r->dev) cudaFree(r->dev);
    delete r;
    return HZSDR_OK;
}

extern "C" int hzsdr_ring_write_peek(hzsdr_ring *r, void **slot_host) {
    if (!r || !slot_host) return fail(HZSDR_ERR_INVALID, "hzsdr_ring_write_peek: null");
    HZ_ENTER_PRODUCER(r->ctx);
    cudaEvent_t pending_copy = nullptr;
    size_t i;
    {
        std::lock_guard<std::mutex> lk(r->mu);
        i = r->widx;
        if (r->landed_valid[i]) pending_copy = r->landed[i];
    }
    // The async H2D enqueued out of this pinned slot `slots` pokes ago may still be parked behind a slow
    // consumer (it waits for consumed[i] on the copy stream).  The producer must not scribble over memory a
    // pending cudaMemcpyAsync will still read, so the slot is handed out only once that copy has executed
    // (outside the lock: readers keep going).  This is where a producer that laps the consumer blocks; the
    // reference's host ring overwrites instead (ring.go:170-186), which DMA in flight does not allow.
    if (pending_copy) HZ_CUDA(cudaEventSynchronize(pending_copy));
    *slot_host = r->host + i * r->slot_bytes;
    return HZSDR_OK;
}

extern "C" int hzsdr_ring_write_poke(hzsdr_ring *r, size_t n_samples) {
    if (!r) return fail(HZSDR_ERR_INVALID, "hzsdr_ring_write_poke: null");
    HZ_ENTER_PRODUCER(r->ctx);
    std::lock_guard<std::mutex> lk(r->mu);
    if (n_samples > r->slot_len) return fail(HZSDR_ERR_DST_TOO_SMALL, "hzsdr_ring_write_poke: %zu > slot %zu", n_samples, r->slot_len);
    const size_t i = r->widx;
    if (r->pending == r->slots) {
        // overrun: the oldest unread slot is overwritten, like the reference (ring.go:170-186)
        if (r->reading && r->ridx == i) return fail(HZSDR_ERR_INVALID, "hzsdr_ring_write_poke: overrun onto the slot being read");
        r->ridx = (r->ridx + 1) % r->slots;
        r->pending--;
    }
    // the device copy of this slot may still be in use by compute enqueued earlier
    if (r->consumed_valid[i]) HZ_CUDA(cudaStreamWaitEvent(r->copy_stream, r->consumed[i], 0));
    const size_t bytes = n_samples * (size_t)hzsdr_format_size(r->format);
    if (bytes)
        HZ_CUDA(cudaMemcpyAsync(r->dev + i * r->slot_bytes, r->host + i * r->slot_bytes, bytes, cudaMemcpyHostToDevice,
                                r->copy_stream));
    HZ_CUDA(cudaEventRecord(r->landed[i], r->copy_stream));
    r->landed_valid[i] = 1;
    r->fill[i] = n_samples;
    r->widx = (i + 1) % r->slots;
    r->pending++;
    return HZSDR_OK;
}

extern "C" int hzsdr_ring_read(hzsdr_ring *r, const void **slot_dev, size_t *n_samples) {
    if (!r || !slot_dev || !n_samples) return fail(HZSDR_ERR_INVALID, "hzsdr_ring_read: null");
    HZ_ENTER(r->ctx);
    std::lock_guard<std::mutex> lk(r->mu);
    if (r->reading) return fail(HZSDR_ERR_INVALID, "hzsdr_ring_read: previous slot not released (hzsdr_ring_read_done)");
    if (r->pending == 0) return fail(HZSDR_ERR_RING_UNDERRUN, "RingBuffer: Buffer Underrun");
    const size_t i = r->ridx;
    HZ_CUDA(cudaStreamWaitEvent(r->ctx->stream, r->landed[i], 0));  // compute waits for the copy, the host does not
    *slot_dev = r->dev + i * r->slot_bytes;
    *n_samples = r->fill[i];
    r->reading = true;
    return HZSDR_OK;
}

extern "C" int hzsdr_ring_read_done(hzsdr_ring *r) {
    if (!r) return fail(HZSDR_ERR_INVALID, "hzsdr_ring_read_done: null");
    HZ_ENTER(r->ctx);
    std::lock_guard<std::mutex> lk(r->mu);
    if (!r->reading) return fail(HZSDR_ERR_INVALID, "hzsdr_ring_read_done: no slot is being read");
    const size_t i = r->ridx;
    HZ_CUDA(cudaEventRecord(r->consumed[i], r->ctx->stream));
    r->consumed_valid[i] = 1;
    r->ridx = (i + 1) % r->slots;
    r->pending--;
    r->reading = false;
    return HZSDR_OK;
}

// =================================================================================================
// NCCL (loaded lazily with dlopen so the library has no link-time NCCL dependency and shares the
// copy PyTorch already loaded when the harness is Python).  Only multi-GPU Beamform uses it.
// =================================================================================================
namespace {
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclFloat32 = 7 };
enum { ncclSum = 0 };
struct Nccl {
    void *h = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*Reduce)(const void *, void *, size_t, int, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
};
Nccl g_nccl;
std::once_flag g_nccl_once;

void load_nccl() {
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        g_nccl.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.h) break;
    }
    if (!g_nccl.h) return;
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(g_nccl.h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(g_nccl.h, "ncclCommInitRank");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(g_nccl.h, "ncclCommDestroy");
    g_nccl.Reduce = (decltype(g_nccl.Reduce))dlsym(g_nccl.h, "ncclReduce");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(g_nccl.h, "ncclAllReduce");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(g_nccl.h, "ncclGetErrorString");
    g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommDestroy && g_nccl.Reduce && g_nccl.AllReduce;
}

int need_nccl() {
    std::call_once(g_nccl_once, load_nccl);
    if (!g_nccl.ok) return fail(HZSDR_ERR_NCCL, "libnccl.so.2 could not be loaded: %s", dlerror() ? dlerror() : "missing symbols");
    return HZSDR_OK;
}
int nccl_fail(const char *what, int rc) {
    return fail(HZSDR_ERR_NCCL, "%s: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "nccl error");
}
}  // namespace

struct hzsdr_comm {
    hzsdr_ctx *ctx = nullptr;
    ncclComm_t comm = nullptr;
    int nranks = 0, rank = 0;
};

extern "C" int hzsdr_comm_unique_id(void *id_out) {
    if (!id_out) return fail(HZSDR_ERR_INVALID, "hzsdr_comm_unique_id: null");
    int rc = need_nccl();
    if (rc) return rc;
    ncclUniqueId id;
    if ((rc = g_nccl.GetUniqueId(&id)) != 0) return nccl_fail("ncclGetUniqueId", rc);
    static_assert(sizeof(id) == HZSDR_NCCL_UNIQUE_ID_BYTES, "unique id size");
    memcpy(id_out, &id, sizeof(id));
    return HZSDR_OK;
}

extern "C" int hzsdr_comm_create(hzsdr_ctx *ctx, int nranks, int rank, const void *id, hzsdr_comm **out) {
    HZ_ENTER(ctx);
    if (!out || !id || nranks < 1 || rank < 0 || rank >= nranks) return fail(HZSDR_ERR_INVALID, "hzsdr_comm_create: bad arguments");
    *out = nullptr;
    int rc = need_nccl();
    if (rc) return rc;
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    hzsdr_comm *c = new hzsdr_comm();
    c->ctx = ctx;
    c->nranks = nranks;
    c->rank = rank;
    if ((rc = g_nccl.CommInitRank(&c->comm, nranks, uid, rank)) != 0) {
        delete c;
        return nccl_fail("ncclCommInitRank", rc);
    }
    *out = c;
    return HZSDR_OK;
}

extern "C" int hzsdr_comm_destroy(hzsdr_comm *c) {
    if (!c) return HZSDR_OK;
    if (c->ctx) cudaSetDevice(c->ctx->device);
    if (c->comm && g_nccl.ok) g_nccl.CommDestroy(c->comm);
    delete c;
    return HZSDR_OK;
}

extern "C" int hzsdr_comm_reduce_c64(hzsdr_comm *c, void *buf, size_t n, int root) {
    if (!c) return fail(HZSDR_ERR_INVALID, "hzsdr_comm_reduce_c64: null");
    HZ_ENTER(c->ctx);
    int rc = g_nccl.Reduce(buf, buf, 2 * n, ncclFloat32, ncclSum, root, c->comm, c->ctx->stream);
    if (rc) return nccl_fail("ncclReduce", rc);
    return HZSDR_OK;
}

extern "C" int hzsdr_comm_allreduce_c64(hzsdr_comm *c, void *buf, size_t n) {
    if (!c) return fail(HZSDR_ERR_INVALID, "hzsdr_comm_allreduce_c64: null");
    HZ_ENTER(c->ctx);
    int rc = g_nccl.AllReduce(buf, buf, 2 * n, ncclFloat32, ncclSum, c->comm, c->ctx->stream);
    if (rc) return nccl_fail("ncclAllReduce", rc);
    return HZSDR_OK;
}
