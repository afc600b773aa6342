// common.cuh -- shared plumbing for libhzsdrcuda.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <mutex>
#include <string>
#include <utility>

#include "../../include/hzsdr_cuda.h"

namespace hz {

constexpr int kNumSMsB200 = 148;
constexpr int kDecimateBlock = 32 * 1024;  // stream/decimate.go:41-42

// ---- thread-local error string (hzsdr_last_error) ----------------------------------------
void set_error(const char *fmt, ...);
int fail(int status, const char *fmt, ...);

#define HZ_CUDA(expr)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return ::hz::fail(HZSDR_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                              __FILE__, __LINE__);                                             \
    } while (0)

#define HZ_CHECK_LAUNCH() HZ_CUDA(cudaGetLastError())

// ---- which launches may overlap their predecessors -------------------------------------------
// Kernels that start with `griddepcontrol.launch_dependents` let the NEXT kernel of the stream begin
// while they drain (programmatic dependent launch).  That is only safe when the next launch neither
// reads nor writes anything an unfinished predecessor writes, and does not write what one still
// reads (e.g. two ConvolutionReaders in series: the second reads the first's output).  The context
// keeps the byte spans touched by every overlappable launch since the last fully serialised one;
// a launch that conflicts with any of them -- or finds the window full -- goes out WITHOUT the
// attribute (the stream then orders it after everything before it) and restarts the window.
// Kernels outside this scheme never trigger early, so whatever follows them is ordered as usual.
// An overlappable kernel must not FINISH before its predecessors ("B done" has to imply "A done"
// for every later operation of the stream), so the last of its CTAs to get to the end executes
// `griddepcontrol.wait` before exiting.  "Last" is counted in a per-launch slot of
// `hzsdr_ctx::overlap_done` (kSlots > kMax, so a slot is never reused before a fully serialised
// launch has drained its earlier user).  Having every CTA wait instead costs 5% on the C2 chain.
struct OverlapWindow {
    struct Span {
        uintptr_t lo, hi;  // [lo, hi)
        bool hits(const Span &o) const { return lo < o.hi && o.lo < hi; }
    };
    static constexpr int kMax = 192, kSlots = 256;
    Span reads[kMax], writes[kMax];
    int n = 0;
    unsigned seq = 0;  // launches admitted so far; slot of the latest = (seq - 1) % kSlots
    int slot() const { return (int)((seq - 1u) % (unsigned)kSlots); }
    static Span span(const void *p, size_t bytes) { return Span{(uintptr_t)p, (uintptr_t)p + bytes}; }
    // true: launch with cudaLaunchAttributeProgrammaticStreamSerialization.  `pred_ok`: the operation
    // right before this launch in the stream is itself an admitted overlappable kernel (or a
    // non-kernel node, which a programmatic edge cannot bypass).  After any OTHER kernel the attribute
    // is withheld: such a kernel never executes griddepcontrol.launch_dependents, its implicit trigger
    // at block exit does not promise that its writes are flushed for a dependent that skips
    // griddepcontrol.wait, and its spans are not in this window -- so the successor stays fully ordered.
    // record without checking (a launch re-recording its own spans after the window was restarted under it)
    bool push(Span r, Span w) {
        if (n >= kMax) return false;
        reads[n] = r;
        writes[n] = w;
        n++;
        return true;
    }
    bool admit(Span r, Span w, bool pred_ok = true) {
        bool ok = pred_ok && n < kMax;
        for (int i = 0; ok && i < n; i++)
            if (w.hits(writes[i]) || w.hits(reads[i]) || r.hits(writes[i])) ok = false;
        if (!ok) n = 0;
        reads[n] = r;
        writes[n] = w;
        n++;
        seq++;
        return ok;
    }
};

#ifdef __CUDACC__
// first / last statement of an overlappable kernel
__device__ __forceinline__ void overlap_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }
// Two ways to end an overlappable kernel.  Streaming kernels with many short-lived CTAs: every thread
// waits (a counter's atomic round trip at the end of each CTA costs them 20%; the wait itself is
// free there because a later CTA only starts once an earlier launch's CTA has left).  Persistent
// kernels whose CTAs live for the whole launch (the chain kernels): only the last CTA waits --
// called by ONE thread per CTA, after that thread's own work (the CTA cannot exit before it returns).
__device__ __forceinline__ void overlap_join_all() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void overlap_join(uint32_t *done) {
    if (atomicAdd(done, 1u) == gridDim.x - 1u) {
        *done = 0u;  // nobody else touches the slot any more: ready for its next user
        asm volatile("griddepcontrol.wait;" ::: "memory");
    }
}
#endif

// Per-device "already done" flags for function attributes and __constant__ tables: both belong to a
// device, and one process may hold contexts on several (one Context per GPU in a multi-GPU ReadBeamform).
constexpr int kMaxDevices = 64;
struct PerDevice {
    std::mutex mu;
    std::atomic<bool> done[kMaxDevices];
    int value[kMaxDevices];  // per-device result of the one-time work (e.g. resident CTAs per SM)
    PerDevice() {
        for (int d = 0; d < kMaxDevices; d++) done[d].store(false), value[d] = 0;
    }
    static int index(int device) { return device >= 0 && device < kMaxDevices ? device : 0; }
    // Runs fn(value&) -- which returns an hzsdr status -- once per device, under the lock; the device is
    // marked done only after fn succeeded, so a second context on the same device used from another
    // thread either waits for the attribute / occupancy calls to finish or repeats them, never skips them.
    template <class F>
    int once(int device, F &&fn) {
        const int d = index(device);
        if (done[d].load(std::memory_order_acquire)) return HZSDR_OK;
        std::lock_guard<std::mutex> lk(mu);
        if (done[d].load(std::memory_order_relaxed)) return HZSDR_OK;
        const int rc = fn(value[d]);
        if (rc == HZSDR_OK) done[d].store(true, std::memory_order_release);
        return rc;
    }
    int get(int device) const { return value[index(device)]; }
};

// fills `cfg` for a launch on the context's stream, with the attribute when `overlap` allows it
inline void overlap_launch_config(cudaLaunchConfig_t &cfg, cudaLaunchAttribute *attr, bool overlap) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = overlap ? 1 : 0;
}

}  // namespace hz

namespace hz {
// ---- staged host path shared by the *_host entry points ---------------------------------------
// Three slots of device staging and three streams: the H2D copy of piece k+1, the kernel of piece k
// and the D2H copy of piece k-1 run concurrently; events order them, the host never waits inside.
// Sources must be pinned (hzsdr_pinned_alloc / ring slots) for the copies to be asynchronous.
struct HostPipe {
    static constexpr int kDepth = 3;
    struct Slot {
        void *in = nullptr, *out = nullptr;
        size_t in_bytes = 0, out_bytes = 0;
        cudaEvent_t in_done = nullptr, k_done = nullptr, out_done = nullptr;
        bool used = false;
    } slot[kDepth];
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    uint64_t pieces = 0;

    cudaError_t init() {
        if (copy_in) return cudaSuccess;
        cudaError_t e = cudaStreamCreateWithFlags(&copy_in, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&copy_out, cudaStreamNonBlocking);
        for (auto &sl : slot) {
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&sl.in_done, cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&sl.k_done, cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&sl.out_done, cudaEventDisableTiming);
        }
        return e;
    }
    // the next slot, grown to hold in_bytes / out_bytes (growing drains the pipe: first use only)
    cudaError_t next(cudaStream_t compute, size_t in_bytes, size_t out_bytes, Slot **out) {
        Slot &sl = slot[pieces % kDepth];
        cudaError_t e = cudaSuccess;
        if (in_bytes > sl.in_bytes || out_bytes > sl.out_bytes) {
            if ((e = drain(compute)) != cudaSuccess) return e;
            if (in_bytes > sl.in_bytes) {
                if (sl.in) cudaFree(sl.in);
                sl.in = nullptr, sl.in_bytes = 0;
                if ((e = cudaMalloc(&sl.in, in_bytes)) != cudaSuccess) return e;
                sl.in_bytes = in_bytes;
            }
            if (out_bytes > sl.out_bytes) {
                if (sl.out) cudaFree(sl.out);
                sl.out = nullptr, sl.out_bytes = 0;
                if ((e = cudaMalloc(&sl.out, out_bytes)) != cudaSuccess) return e;
                sl.out_bytes = out_bytes;
            }
        }
        // the slot's input may be overwritten once the kernel that last read it is done
        if (sl.used) e = cudaStreamWaitEvent(copy_in, sl.k_done, 0);
        *out = &sl;
        return e;
    }
    // after the piece's H2D copies were enqueued on copy_in: the kernel needs them landed and the
    // slot's previous result drained
    cudaError_t before_kernel(cudaStream_t compute, Slot &sl) {
        cudaError_t e = cudaEventRecord(sl.in_done, copy_in);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(compute, sl.in_done, 0);
        if (e == cudaSuccess && sl.used) e = cudaStreamWaitEvent(compute, sl.out_done, 0);
        return e;
    }
    // after the piece's kernels were enqueued on `compute`: copy_out may start once they are done
    cudaError_t after_kernel(cudaStream_t compute, Slot &sl) {
        cudaError_t e = cudaEventRecord(sl.k_done, compute);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(copy_out, sl.k_done, 0);
        return e;
    }
    // after the piece's D2H copies were enqueued on copy_out
    cudaError_t done(Slot &sl) {
        cudaError_t e = cudaEventRecord(sl.out_done, copy_out);
        sl.used = true;
        pieces++;
        return e;
    }
    cudaError_t drain(cudaStream_t compute) {
        cudaError_t e = copy_in ? cudaStreamSynchronize(copy_in) : cudaSuccess;
        if (e == cudaSuccess) e = cudaStreamSynchronize(compute);
        if (e == cudaSuccess && copy_out) e = cudaStreamSynchronize(copy_out);
        return e;
    }
    void destroy() {
        for (auto &sl : slot) {
            if (sl.in) cudaFree(sl.in);
            if (sl.out) cudaFree(sl.out);
            if (sl.in_done) cudaEventDestroy(sl.in_done);
            if (sl.k_done) cudaEventDestroy(sl.k_done);
            if (sl.out_done) cudaEventDestroy(sl.out_done);
            sl = Slot{};
        }
        if (copy_in) cudaStreamDestroy(copy_in);
        if (copy_out) cudaStreamDestroy(copy_out);
        copy_in = copy_out = nullptr;
    }
};
}  // namespace hz

// one GPU + one stream.  Public as an opaque handle.
struct hzsdr_ctx {
    int device = -1;
    cudaStream_t stream = nullptr;
    cudaDeviceProp prop{};
    int sm_count = 0;
    // scratch for multi-kernel operations (big FFTs); used in stream order, grown on demand
    void *workspace = nullptr;
    size_t workspace_bytes = 0;
    hz::OverlapWindow overlap;  // spans of the launches that may still be running early (see above)
    // Which API call enqueued the stream's latest overlappable kernel.  api_seq counts outermost API
    // entries on this context (HZ_ENTER); a launch may carry the programmatic-serialization attribute
    // only if the previous overlappable launch came from this call or the one right before it -- i.e.
    // no other entry point (whose kernels are outside the overlap scheme) ran in between.  Entry points
    // that launch a non-scheme kernel AFTER a scheme one inside the same call reset scheme_seq themselves.
    mutable uint64_t api_seq = 0;
    mutable int api_depth = 0;
    uint64_t scheme_seq = ~0ull;
    bool overlap_pred_ok() const { return scheme_seq == api_seq || scheme_seq + 1 == api_seq; }
    void overlap_launched() { scheme_seq = api_seq; }
    void overlap_broken() { scheme_seq = ~0ull; }  // a kernel outside the scheme was just enqueued
    uint32_t *overlap_done = nullptr;  // device, OverlapWindow::kSlots zeroed counters
    hz::HostPipe host_pipe;            // staging of hzsdr_beamform_host / hzsdr_channelizer_exec_host
};

namespace hz {

// Every entry point pins the calling thread to the context's device first: cgo callers have no
// thread affinity, so the CUDA "current device" can never be relied on.
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    const hzsdr_ctx *ctx;
    const bool counted;
    // counted = false: the producer side of a ring (hzsdr_ring_write_*), which runs on its own thread beside the
    // context's owner and must not touch the owner's unsynchronised API bookkeeping
    explicit DeviceGuard(const hzsdr_ctx *c, bool counted_ = true) : ctx(c), counted(counted_) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != ctx->device) ok = (cudaSetDevice(ctx->device) == cudaSuccess);
        if (counted && ctx->api_depth++ == 0) ctx->api_seq++;  // entry points call each other: count the outermost only
    }
    ~DeviceGuard() {
        if (counted) ctx->api_depth--;
        // leave the device selected: restoring costs a call per entry and nothing relies on it
    }
};

#define HZ_ENTER_PRODUCER(ctx)                                                        \
    if (!(ctx)) return ::hz::fail(HZSDR_ERR_INVALID, "%s: null context", __func__);   \
    ::hz::DeviceGuard _guard(ctx, false);                                             \
    if (!_guard.ok) return ::hz::fail(HZSDR_ERR_CUDA, "%s: cudaSetDevice(%d) failed", __func__, (ctx)->device)
#define HZ_ENTER(ctx)                                                                 \
    if (!(ctx)) return ::hz::fail(HZSDR_ERR_INVALID, "%s: null context", __func__);   \
    ::hz::DeviceGuard _guard(ctx);                                                    \
    if (!_guard.ok) return ::hz::fail(HZSDR_ERR_CUDA, "%s: cudaSetDevice(%d) failed", __func__, (ctx)->device)

// at least `bytes` of device scratch owned by the context (api.cu); contents are valid only until the
// next operation on the context that asks for scratch
int ctx_workspace(hzsdr_ctx *ctx, size_t bytes, void **out);

// grid sizing for streaming kernels: a multiple of the SM count, capped by the work
inline int stream_grid(const hzsdr_ctx *ctx, size_t work_items, int threads, int blocks_per_sm) {
    size_t need = (work_items + (size_t)threads - 1) / (size_t)threads;
    size_t cap = (size_t)ctx->sm_count * (size_t)blocks_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

// ---- compile-time loop -----------------------------------------------------------------------
template <int... Is, class F>
__host__ __device__ __forceinline__ void static_for_impl(std::integer_sequence<int, Is...>, F &&f) {
    (f(std::integral_constant<int, Is>{}), ...);
}
template <int N, class F>
__host__ __device__ __forceinline__ void static_for(F &&f) {
    static_for_impl(std::make_integer_sequence<int, N>{}, static_cast<F &&>(f));
}

// ---- streaming loads/stores: data touched once, keep it out of L1 ---------------------------
__device__ __forceinline__ uint32_t ld_stream_u32(const void *p) {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint2 ld_stream_u64(const void *p) {
    uint2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 ld_stream_u128(const void *p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ld_stream_f4(const void *p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
// same, for buffers the kernel also writes (in-place ops): no .nc
__device__ __forceinline__ float4 ld_inplace_f4(const void *p) {
    float4 v;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p)
                 : "memory");
    return v;
}
__device__ __forceinline__ float2 ld_stream_f2(const void *p) {
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream_f4(void *p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void st_stream_f2(void *p, float2 v) {
    asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}

// ---- asynchronous global -> shared copies (LDGSTS): no registers, no waiting until the data is used
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---- packed fp32 pairs --------------------------------------------------------------------
// sm_100's FADD2 / FMUL2 / FFMA2 work on a 64-bit register pair and take, as free operand forms, the
// pair as it is, the pair with its halves swapped, one scalar register broadcast to both halves and
// a negation of either half.  ptxas folds the pack / unpack moves below into those forms, so complex
// arithmetic on (re, im) pairs costs half the issue slots of the scalar code; every component is the
// same IEEE operation (round-to-nearest add / mul / fma) as before, so results are bit-identical.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float2 upk2(f32x2 v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
// (a.x + b.x, a.y + b.y)
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a.x, a.y)), "l"(pk2(b.x, b.y)));
    return upk2(r);
}
// (a.x * b.x, a.y * b.y)
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a.x, a.y)), "l"(pk2(b.x, b.y)));
    return upk2(r);
}
// (fma(a.x, b.x, c.x), fma(a.y, b.y, c.y))
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pk2(a.x, a.y)), "l"(pk2(b.x, b.y)), "l"(pk2(c.x, c.y)));
    return upk2(r);
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return add2(a, make_float2(-b.x, -b.y)); }

// ---- integer -> float32 conversion, bit-exact against the reference --------------------------
// x / d for the two divisors the reference uses, as three FP32 ops instead of a div.rn
// subroutine: q = x*r; e = fma(-q, d, x) (exact residual); q' = fma(e, r, q).  Verified
// exhaustively (exact rational arithmetic) to equal IEEE x/d for every u8 code with d = 127.5
// and every i16 code with d = 32767; tests/test_gpu_parity.py re-checks all codes on the device.
__device__ __forceinline__ float div_exact(float x, float d, float r) {
    float q = x * r;
    float e = fmaf(-q, d, x);
    return fmaf(e, r, q);
}
// ---- integer -> float without the conversion unit -----------------------------------------------
// One PRMT drops the integer into the mantissa of a float whose exponent makes the mantissa LSB the
// right weight, one FADD removes the offset; both run on the full-rate ALU/FMA pipes (I2F.S8/U8 go
// through the quarter-rate XU pipe).  All results are exact.
//   u8 : 0x4700bb00 = 2^15 + b  (LSB 2^-8, so .5 is representable)  ->  - 32895.5 = b - 127.5
//   i8 : flip the sign bit first (b ^ 0x80 = b + 128)               ->  - 32896.0 = b
//   i16: 0x4B00hhll = 2^23 + (v ^ 0x8000)                           ->  - 8421376 = v
template <int K>
__device__ __forceinline__ float u8_centered(uint32_t w) {  // byte K of w, minus 127.5
    return __uint_as_float(__byte_perm(w, 0x47000000u, 0x7504u | (K << 4))) - 32895.5f;
}
template <int K>
__device__ __forceinline__ float i8_exact(uint32_t w_flipped) {  // byte K of (w ^ 0x80808080), as a signed value
    return __uint_as_float(__byte_perm(w_flipped, 0x47000000u, 0x7504u | (K << 4))) - 32896.0f;
}
template <int Hh>
__device__ __forceinline__ float i16_exact(uint32_t w_flipped) {  // half Hh of (w ^ 0x80008000), as a signed value
    return __uint_as_float(__byte_perm(w_flipped, 0x4B000000u, 0x7400u | ((2 * Hh + 1) << 4) | (2 * Hh))) - 8421376.0f;
}
// the same, both components of one IQ sample with a single packed add (bytes K, K+1 / both halves)
template <int K>
__device__ __forceinline__ float2 u8_centered_pair(uint32_t w) {
    return add2(make_float2(__uint_as_float(__byte_perm(w, 0x47000000u, 0x7504u | (K << 4))),
                            __uint_as_float(__byte_perm(w, 0x47000000u, 0x7504u | ((K + 1) << 4)))),
                make_float2(-32895.5f, -32895.5f));
}
template <int K>
__device__ __forceinline__ float2 i8_exact_pair(uint32_t w_flipped) {
    return add2(make_float2(__uint_as_float(__byte_perm(w_flipped, 0x47000000u, 0x7504u | (K << 4))),
                            __uint_as_float(__byte_perm(w_flipped, 0x47000000u, 0x7504u | ((K + 1) << 4)))),
                make_float2(-32896.0f, -32896.0f));
}
__device__ __forceinline__ float2 i16_exact_pair(uint32_t w_flipped) {
    return add2(make_float2(__uint_as_float(__byte_perm(w_flipped, 0x4B000000u, 0x7410u)),
                            __uint_as_float(__byte_perm(w_flipped, 0x4B000000u, 0x7432u))),
                make_float2(-8421376.0f, -8421376.0f));
}

// The reference's conversions (bit-exact):
//   iq_u8.go:116-119 == iq_u8_amd64.s:79-80: (float32(b) - 127.5) / 127.5
//   iq_i8.go:114-117: float32(b) / 128 (exact: power of two)
//   iq_i16.go:142-143: float32(v) / 32767
__device__ __forceinline__ float u8_div(float centered) { return div_exact(centered, 127.5f, 1.0f / 127.5f); }
__device__ __forceinline__ float i16_div(float v) { return div_exact(v, 32767.0f, 1.0f / 32767.0f); }

// RawTraits<FMT>: one IQ sample packed in the low bits of a 32-bit word.
//   conv(w)      the reference's value, bit-exact
//   unscaled2(w) the same value as unscaled(w) with the offset removed by ONE packed add: for the
//                FFT chain kernels, which are issue-bound; the streaming kernels keep the scalar form
//                (the pair alignment costs them registers, and occupancy is what hides their HBM latency)
//   unscaled(w)  (b-127.5, ...) / (b, ...) / (v, ...): exact, to be multiplied by scale() by callers
//                that fold the scale into a following multiply (tolerance-bound paths only: for u8 and
//                i16 the fold rounds differently from the reference's division by <= 1 ulp)
template <int FMT>
struct RawTraits;
template <>
struct RawTraits<HZSDR_FORMAT_U8> {
    static constexpr int bytes = 2;
    static __device__ __forceinline__ float scale() { return 1.0f / 127.5f; }
    static __device__ __forceinline__ float2 unscaled(uint32_t w) { return make_float2(u8_centered<0>(w), u8_centered<1>(w)); }
    static __device__ __forceinline__ float2 unscaled_hi(uint32_t w) { return make_float2(u8_centered<2>(w), u8_centered<3>(w)); }
    static __device__ __forceinline__ float2 unscaled2(uint32_t w) { return u8_centered_pair<0>(w); }
    static __device__ __forceinline__ float2 conv(uint32_t w) { return make_float2(u8_div(u8_centered<0>(w)), u8_div(u8_centered<1>(w))); }
    static __device__ __forceinline__ float2 conv_hi(uint32_t w) { return make_float2(u8_div(u8_centered<2>(w)), u8_div(u8_centered<3>(w))); }
};
template <>
struct RawTraits<HZSDR_FORMAT_I8> {
    static constexpr int bytes = 2;
    static __device__ __forceinline__ float scale() { return 0.0078125f; }
    static __device__ __forceinline__ float2 unscaled(uint32_t w) {
        w ^= 0x80808080u;
        return make_float2(i8_exact<0>(w), i8_exact<1>(w));
    }
    static __device__ __forceinline__ float2 unscaled_hi(uint32_t w) {
        w ^= 0x80808080u;
        return make_float2(i8_exact<2>(w), i8_exact<3>(w));
    }
    static __device__ __forceinline__ float2 unscaled2(uint32_t w) { return i8_exact_pair<0>(w ^ 0x80808080u); }
    static __device__ __forceinline__ float2 conv(uint32_t w) {
        const float2 v = unscaled(w);
        return make_float2(v.x * 0.0078125f, v.y * 0.0078125f);
    }
    static __device__ __forceinline__ float2 conv_hi(uint32_t w) {
        const float2 v = unscaled_hi(w);
        return make_float2(v.x * 0.0078125f, v.y * 0.0078125f);
    }
};
template <>
struct RawTraits<HZSDR_FORMAT_I16> {
    static constexpr int bytes = 4;
    static __device__ __forceinline__ float scale() { return 1.0f / 32767.0f; }
    static __device__ __forceinline__ float2 unscaled(uint32_t w) {
        w ^= 0x80008000u;
        return make_float2(i16_exact<0>(w), i16_exact<1>(w));
    }
    static __device__ __forceinline__ float2 unscaled2(uint32_t w) { return i16_exact_pair(w ^ 0x80008000u); }
    static __device__ __forceinline__ float2 conv(uint32_t w) {
        const float2 v = unscaled(w);
        return make_float2(i16_div(v.x), i16_div(v.y));
    }
};

// Go's complex64 multiply: products and sums in fp64, one narrowing per component
// (internal/simd/mult.go:29-33 as compiled by gc; oracle: go_complex64_mul).
__device__ __forceinline__ float2 go_cmul(float2 a, float2 b) {
    double ar = a.x, ai = a.y, br = b.x, bi = b.y;
    return make_float2((float)(ar * br - ai * bi), (float)(ar * bi + ai * br));
}
// fp32 complex multiply for the tolerance-bound paths (<= 1 ulp per component from go_cmul):
// (fma(a.x, b.x, -(a.y b.y)), fma(a.x, b.y, a.y b.x)) in two packed instructions
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    // (the half-negated, swapped pair has to be the FIRST operand: the only slot where ptxas folds it)
    return fma2(b, make_float2(a.x, a.x), mul2(make_float2(-b.y, b.x), make_float2(a.y, a.y)));
}

// cmul with the result's halves exchanged, (im, re): what the re/im-swapped inverse transform eats.
// Swapping a finished pair costs two MOVs; producing it swapped costs nothing.
__device__ __forceinline__ float2 cmul_swapped(float2 a, float2 b) {
    return fma2(make_float2(b.y, b.x), make_float2(a.x, a.x), mul2(make_float2(b.x, -b.y), make_float2(a.y, a.y)));
}

// acc += swap(x * h): the accumulator holds (im, re)
__device__ __forceinline__ float2 cmac_swapped(float2 acc, float2 x, float2 h) {
    acc = fma2(make_float2(h.y, h.x), make_float2(x.x, x.x), acc);
    return fma2(make_float2(h.x, -h.y), make_float2(x.y, x.y), acc);
}

}  // namespace hz
