// fft.cu -- K6 (FFT planner, ConvolveFreq) and the fused chain kernel
//           Convert -> Shift -> FFT -> xH -> IFFT -> Decimate.
#include <algorithm>
#include <map>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "fft.cuh"
#include "nco.cuh"
#include "batch_host.h"
#include "nco_launch.h"
#include "fft_kernels.cuh"

namespace hz {

// ---- twiddle tables, cached per (device, N) ------------------------------------------------------
static std::mutex g_tw_mu;
static std::map<std::pair<int, int>, float2 *> g_tw;

static int get_twiddles(hzsdr_ctx *ctx, int n, const float2 **out) {
    std::lock_guard<std::mutex> lk(g_tw_mu);
    auto key = std::make_pair(ctx->device, n);
    auto it = g_tw.find(key);
    if (it != g_tw.end()) {
        *out = it->second;
        return HZSDR_OK;
    }
    std::vector<float2> h(n);
    for (int m = 0; m < n; m++) {
        const double a = 2.0 * M_PI * (double)m / (double)n;
        h[m] = make_float2((float)cos(a), (float)sin(a));
    }
    float2 *d = nullptr;
    HZ_CUDA(cudaMalloc((void **)&d, sizeof(float2) * n));
    // synchronous pageable copy: the host vector may die on return
    HZ_CUDA(cudaMemcpy(d, h.data(), sizeof(float2) * n, cudaMemcpyHostToDevice));
    g_tw[key] = d;
    *out = d;
    return HZSDR_OK;
}

// RawTraits<FMT>::scale() on the host
static float format_scale(int fmt) {
    switch (fmt) {
        case HZSDR_FORMAT_U8: return 1.0f / 127.5f;
        case HZSDR_FORMAT_I8: return 0.0078125f;
        case HZSDR_FORMAT_I16: return 1.0f / 32767.0f;
        default: return 1.0f;
    }
}

// the [tw | twB | twC] table set of chain1024.cu, cached per device
static std::map<int, float2 *> g_tw1024;
static int get_chain1024_tables(hzsdr_ctx *ctx, const float2 **out) {
    std::lock_guard<std::mutex> lk(g_tw_mu);
    auto it = g_tw1024.find(ctx->device);
    if (it != g_tw1024.end()) {
        *out = it->second;
        return HZSDR_OK;
    }
    std::vector<float2> t(kChain1024TableLen);
    chain1024_twiddles(t.data());
    float2 *d = nullptr;
    HZ_CUDA(cudaMalloc((void **)&d, sizeof(float2) * t.size()));
    HZ_CUDA(cudaMemcpy(d, t.data(), sizeof(float2) * t.size(), cudaMemcpyHostToDevice));
    g_tw1024[ctx->device] = d;
    *out = d;
    return HZSDR_OK;
}

#define HZ_DISPATCH_N(n, CALL)                                                        \
    switch (n) {                                                                      \
        case 2: return CALL(2);                                                       \
        case 4: return CALL(4);                                                       \
        case 8: return CALL(8);                                                       \
        case 16: return CALL(16);                                                     \
        case 32: return CALL(32);                                                     \
        case 64: return CALL(64);                                                     \
        case 128: return CALL(128);                                                   \
        case 256: return CALL(256);                                                   \
        case 512: return CALL(512);                                                   \
        case 1024: return CALL(1024);                                                 \
        case 2048: return CALL(2048);                                                 \
        case 4096: return CALL(4096);                                                 \
        case 8192: return CALL(8192);                                                 \
        case 16384: return CALL(16384);                                               \
        default: return fail(HZSDR_ERR_UNSUPPORTED, "FFT length %zu: need a power of two in [2, 16384]", (size_t)(n)); \
    }

static int dispatch_fft(hzsdr_ctx *ctx, size_t n, int dir, const float2 *src, float2 *dst, size_t batch, const float2 *tw) {
#define CALL(NN) launch_fft<NN>(ctx, dir, src, dst, batch, tw)
    HZ_DISPATCH_N(n, CALL)
#undef CALL
}
static int dispatch_convolve(hzsdr_ctx *ctx, size_t n, const float2 *src, float2 *dst, size_t nblocks, const float2 *tw,
                             const float2 *H, size_t src_stride, size_t h_stride = 0) {
#define CALL(NN) launch_convolve<NN>(ctx, src, dst, nblocks, tw, H, src_stride, h_stride)
    HZ_DISPATCH_N(n, CALL)
#undef CALL
}
static int dispatch_chain(hzsdr_ctx *ctx, size_t n, int fmt, const ChainParams &prm, const NcoTable &nco) {
#define CALL(NN) launch_chain<NN>(ctx, fmt, prm, nco)
    HZ_DISPATCH_N(n, CALL)
#undef CALL
}

static bool fft_len_ok(size_t n) { return n >= 2 && n <= 16384 && (n & (n - 1)) == 0; }

// any supported length: one kernel up to 16384 points, two through the context's scratch beyond
int fft_any(hzsdr_ctx *ctx, size_t n, int direction, const float2 *src, float2 *dst, size_t batch) {
    if (fft_len_ok(n)) {
        const float2 *tw = nullptr;
        int rc = get_twiddles(ctx, (int)n, &tw);
        if (rc) return rc;
        return dispatch_fft(ctx, n, direction, src, dst, batch, tw);
    }
    if (!bigfft_len_ok(n))
        return fail(HZSDR_ERR_UNSUPPORTED, "FFT length %zu: need a power of two in [2, 2^20]", n);
    int n1, n2;
    bigfft_factors(n, &n1, &n2);
    const float2 *tw1 = nullptr, *tw2 = nullptr;
    int rc = get_twiddles(ctx, n1, &tw1);
    if (!rc) rc = get_twiddles(ctx, n2, &tw2);
    if (rc) return rc;
    void *ws = nullptr;
    rc = ctx_workspace(ctx, n * batch * sizeof(float2), &ws);
    if (rc) return rc;
    return launch_bigfft(ctx, n, direction == HZSDR_FFT_FORWARD ? FFT_FWD : FFT_BWD, src, dst, (float2 *)ws, batch, tw1, tw2);
}

}  // namespace hz

using namespace hz;

// =================================================================================================
// C ABI: planner
// =================================================================================================
struct hzsdr_fft_plan {
    hzsdr_ctx *ctx;
    size_t n;
    int direction;
    const float2 *tw;
};

extern "C" int hzsdr_fft_plan_create(hzsdr_ctx *ctx, size_t iq_len, size_t freq_len, int direction,
                                     hzsdr_fft_plan **out) {
    HZ_ENTER(ctx);
    if (!out) return fail(HZSDR_ERR_INVALID, "hzsdr_fft_plan_create: null out");
    *out = nullptr;
    if (iq_len != freq_len)  // testutils/fft.go:127-138
        return fail(HZSDR_ERR_DST_TOO_SMALL, "hzsdr_fft_plan_create: iq length %zu != frequency length %zu", iq_len, freq_len);
    if (!fft_len_ok(iq_len) && !bigfft_len_ok(iq_len))
        return fail(HZSDR_ERR_UNSUPPORTED, "hzsdr_fft_plan_create: length %zu: need a power of two in [2, 2^20]", iq_len);
    if (direction != HZSDR_FFT_FORWARD && direction != HZSDR_FFT_BACKWARD)
        return fail(HZSDR_ERR_INVALID, "hzsdr_fft_plan_create: direction %d", direction);
    const float2 *tw = nullptr;
    if (fft_len_ok(iq_len)) {
        int rc = get_twiddles(ctx, (int)iq_len, &tw);
        if (rc) return rc;
    }
    *out = new hzsdr_fft_plan{ctx, iq_len, direction, tw};
    return HZSDR_OK;
}

extern "C" int hzsdr_fft_exec(hzsdr_fft_plan *plan, const void *src, void *dst, size_t batch) {
    if (!plan) return fail(HZSDR_ERR_INVALID, "hzsdr_fft_exec: null plan");
    HZ_ENTER(plan->ctx);
    if (batch == 0) return HZSDR_OK;
    if (!src || !dst) return fail(HZSDR_ERR_INVALID, "hzsdr_fft_exec: null buffer");
    if (batch > 0xffffffffull) return fail(HZSDR_ERR_INVALID, "hzsdr_fft_exec: batch too large");
    if (!plan->tw)  // beyond 16384 points: two kernels through the context's scratch (bigfft.cu)
        return fft_any(plan->ctx, plan->n, plan->direction, (const float2 *)src, (float2 *)dst, batch);
    return dispatch_fft(plan->ctx, plan->n, plan->direction, (const float2 *)src, (float2 *)dst, batch, plan->tw);
}

extern "C" int hzsdr_fft_plan_destroy(hzsdr_fft_plan *plan) {
    delete plan;  // twiddles belong to the per-device cache
    return HZSDR_OK;
}

extern "C" int hzsdr_convolve_freq(hzsdr_ctx *ctx, const void *src, void *dst, const void *filter, size_t n_fft,
                                   size_t n_blocks) {
    HZ_ENTER(ctx);
    if (!fft_len_ok(n_fft))
        return fail(HZSDR_ERR_UNSUPPORTED, "hzsdr_convolve_freq: length %zu: need a power of two in [2, 16384]", n_fft);
    if (n_blocks == 0) return HZSDR_OK;
    if (!src || !dst || !filter) return fail(HZSDR_ERR_INVALID, "hzsdr_convolve_freq: null buffer");
    if (n_fft == 1024 && n_blocks <= 0x3fffffu) {
        // warp-per-block kernel of chain1024.cu with complex64 input, no mixer, nothing decimated
        const float2 *tw1k = nullptr;
        int rc = get_chain1024_tables(ctx, &tw1k);
        if (rc) return rc;
        ChainParams prm{};
        prm.src = (const uint8_t *)src;
        prm.dst = (float2 *)dst;
        prm.tw = tw1k;
        prm.H = (const float2 *)filter;
        prm.nblocks = (uint32_t)n_blocks;
        prm.z0 = 0;
        prm.D = 1;
        prm.db_log2 = 10;  // "decimate block" = one FFT block: every sample is kept
        prm.M = 1024;
        prm.inv_d = 0.99999994f;
        static const NcoTable none{};
        return launch_chain1024(ctx, HZSDR_FORMAT_C64, prm, none);
    }
    const float2 *tw = nullptr;
    int rc = get_twiddles(ctx, (int)n_fft, &tw);
    if (rc) return rc;
    return dispatch_convolve(ctx, n_fft, (const float2 *)src, (float2 *)dst, n_blocks, tw, (const float2 *)filter, n_fft);
}

// internal: ConvolveFreq over windows that start every `src_stride` samples (fir.cu, overlap-save)
namespace hz {
int convolve_windows(hzsdr_ctx *ctx, const float2 *src, float2 *dst, const float2 *filter, size_t n_fft, size_t n_windows,
                     size_t src_stride) {
    const float2 *tw = nullptr;
    int rc = get_twiddles(ctx, (int)n_fft, &tw);
    if (rc) return rc;
    return dispatch_convolve(ctx, n_fft, src, dst, n_windows, tw, filter, src_stride);
}
}  // namespace hz

// fft.Convolve / fft.CrossCorrelate (fft/convolution.go:97-139): dst = IFFT(FFT(iq1) * FFT(iq2)) or,
// for the correlation, IFFT(FFT(iq1) * conj(FFT(iq2))), over `batch` length-n vectors.
namespace hz {
__global__ void __launch_bounds__(256) k_conj(float2 *x, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) x[i].y = -x[i].y;
}
// a[i] *= b[i] or a[i] *= conj(b[i]) (fft/convolution.go:113-115,131-136), fp32 like the fused kernel
__global__ void __launch_bounds__(256) k_spectrum_mul(float2 *a, const float2 *__restrict__ b, size_t n, int conj_b) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float2 h = b[i];
        if (conj_b) h.y = -h.y;
        a[i] = cmul(a[i], h);
    }
}
}  // namespace hz

extern "C" int hzsdr_fft_convolve(hzsdr_ctx *ctx, void *dst, const void *iq1, const void *iq2, size_t n, size_t batch,
                                  int cross_correlate, void *scratch) {
    HZ_ENTER(ctx);
    if (!fft_len_ok(n) && !bigfft_len_ok(n))
        return fail(HZSDR_ERR_UNSUPPORTED, "hzsdr_fft_convolve: length %zu: need a power of two in [2, 2^20]", n);
    if (batch == 0) return HZSDR_OK;
    if (!dst || !iq1 || !iq2 || !scratch) return fail(HZSDR_ERR_INVALID, "hzsdr_fft_convolve: null buffer (scratch: n*batch complex64)");
    if (!fft_len_ok(n)) {  // the Kerberos aligner's 65536 points (rtl/kerberos/internal/align.go:44-55): unfused
        int rc = fft_any(ctx, n, HZSDR_FFT_FORWARD, (const float2 *)iq2, (float2 *)scratch, batch);
        if (!rc) rc = fft_any(ctx, n, HZSDR_FFT_FORWARD, (const float2 *)iq1, (float2 *)dst, batch);
        if (rc) return rc;
        const size_t tot = n * batch;
        k_spectrum_mul<<<(int)std::min<size_t>((tot + 255) / 256, (size_t)ctx->sm_count * 8), 256, 0, ctx->stream>>>(
            (float2 *)dst, (const float2 *)scratch, tot, cross_correlate);
        HZ_CHECK_LAUNCH();
        return fft_any(ctx, n, HZSDR_FFT_BACKWARD, (const float2 *)dst, (float2 *)dst, batch);
    }
    const float2 *tw = nullptr;
    int rc = get_twiddles(ctx, (int)n, &tw);
    if (rc) return rc;
    rc = dispatch_fft(ctx, n, HZSDR_FFT_FORWARD, (const float2 *)iq2, (float2 *)scratch, batch, tw);  // freq2
    if (rc) return rc;
    if (cross_correlate) {  // freq1[i] * complex(real(freq2[i]), -imag(freq2[i])), fft/convolution.go:131-136
        const size_t tot = n * batch;
        k_conj<<<(int)std::min<size_t>((tot + 255) / 256, (size_t)ctx->sm_count * 8), 256, 0, ctx->stream>>>((float2 *)scratch, tot);
        HZ_CHECK_LAUNCH();
    }
    // forward of iq1, x freq2, backward -- fused, freq1 never leaves the SM
    return dispatch_convolve(ctx, n, (const float2 *)iq1, (float2 *)dst, batch, tw, (const float2 *)scratch, n, n);
}

// =================================================================================================
// C ABI: fused chain
// =================================================================================================
// Descriptors of a batched launch and the pool of accumulator segments they point into: pinned staging,
// kStages deep so that the host never rewrites a slot an earlier copy has yet to read, and one device image
// (copies and launches are ordered on the context's stream).  Layout: [StreamDesc x cap | NcoSegment pool].
struct BatchStaging {
    static constexpr int kStages = 3;
    uint8_t *host[kStages] = {};
    cudaEvent_t done[kStages] = {};
    bool used[kStages] = {};
    uint8_t *dev = nullptr;
    size_t cap = 0;  // streams
    uint64_t calls = 0;
    int stage = 0;
    size_t nbatch = 0, nsegs = 0;

    static size_t seg_cap(size_t cap) { return cap * (size_t)kMaxSegsPerLaunch; }
    void release() {
        for (int i = 0; i < kStages; i++) {
            if (host[i]) cudaFreeHost(host[i]);
            if (done[i]) cudaEventDestroy(done[i]);
            host[i] = nullptr;
            done[i] = nullptr;
            used[i] = false;
        }
        if (dev) cudaFree(dev);
        dev = nullptr;
        cap = 0;
    }
    int reserve(hzsdr_ctx *ctx, size_t streams) {
        if (streams <= cap) return HZSDR_OK;
        HZ_CUDA(cudaStreamSynchronize(ctx->stream));
        release();
        const size_t bytes = sizeof(StreamDesc) * streams + sizeof(NcoSegment) * seg_cap(streams);
        for (int i = 0; i < kStages; i++) {
            HZ_CUDA(cudaHostAlloc((void **)&host[i], bytes, cudaHostAllocPortable));
            HZ_CUDA(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
        }
        HZ_CUDA(cudaMalloc((void **)&dev, bytes));
        cap = streams;
        return HZSDR_OK;
    }
    int begin() {  // next staging slot, once the copy that last read it is done
        stage = (int)(calls % kStages);
        if (used[stage]) HZ_CUDA(cudaEventSynchronize(done[stage]));
        nbatch = nsegs = 0;
        return HZSDR_OK;
    }
    StreamDesc *descs() { return reinterpret_cast<StreamDesc *>(host[stage]); }
    const StreamDesc *dev_descs() const { return reinterpret_cast<const StreamDesc *>(dev); }
    const NcoSegment *dev_pool() const { return reinterpret_cast<const NcoSegment *>(dev + sizeof(StreamDesc) * cap); }
    // one more stream whose buffer is covered by a single launch table
    void push(const void *src, void *dst, const NcoTable &table) {
        NcoSegment *pool = reinterpret_cast<NcoSegment *>(host[stage] + sizeof(StreamDesc) * cap);
        fill_desc(descs()[nbatch++], pool, (uint32_t)nsegs, src, dst, table);
        nsegs += (size_t)table.count;
    }
    int upload(cudaStream_t st) {
        if (!nbatch) return HZSDR_OK;
        const size_t seg_off = sizeof(StreamDesc) * cap;
        HZ_CUDA(cudaMemcpyAsync(dev, host[stage], sizeof(StreamDesc) * nbatch, cudaMemcpyHostToDevice, st));
        HZ_CUDA(cudaMemcpyAsync(dev + seg_off, host[stage] + seg_off, sizeof(NcoSegment) * nsegs, cudaMemcpyHostToDevice, st));
        HZ_CUDA(cudaEventRecord(done[stage], st));
        used[stage] = true;
        calls++;
        return HZSDR_OK;
    }
};

struct hzsdr_chain {
    hzsdr_ctx *ctx = nullptr;
    hzsdr_chain_config cfg{};
    uint32_t decim_block = 0, db_log2 = 0, per_block = 0;
    float inv_d = 0.f;
    const float2 *tw = nullptr;
    float2 *H = nullptr;  // device copy of the filter
    float2 *tw1024 = nullptr;  // this chain's table set of the N = 1024 kernel (chain1024.cu): [32][32 | twB | twC]
    // split form (even decimation factor): 32 x 32 tables that carry e^{-i r dP} * scale for a launch's
    // dominant phase step, cached by dP -- built on the host and copied in stream order on a miss
    static constexpr int kSplitSlots = 16;  // one per accumulator binade that can dominate a launch: they recur every 2*pi wrap
    bool can_split = false;
    float2 *split_dev = nullptr, *split_stage = nullptr;  // kSplitSlots x 32 x 32: device tables, pinned staging
    uint64_t split_dp[kSplitSlots] = {};
    int split_used = 0;
    float2 *twk = nullptr;     // tables of the N = K * 1024 kernel (chaink.cu): [(K-1)*1024 twiddles | N permuted filter]
    const float2 *tw1k_plain = nullptr;  // the device's plain [32][32] W_1024 table (shared, not owned)
    int kfac = 0;              // K = n_fft / 1024 when twk is set
    float2 *tw16k = nullptr;   // tables of the N = 16384 kernel (chain16k.cu): [31*32 | 15*1024 | 16384 permuted filter]
    // overlap-save form of the N = 16384 kernel (cfg.overlap_save_taps > 0): windows every os_hop samples,
    // os_hist = 16384 - os_hop raw samples (and the accumulator segments that cover them) carried between calls
    uint32_t os_hop = 0, os_hist = 0;
    uint8_t *hist_raw = nullptr;       // device, 2 x os_hist raw samples (ping-pong): the tail of the previous call's buffer,
    int hist_cur = 0;                  // written by that call's own kernel (no copy in the stream: the launches overlap)
    std::vector<HostSeg> os_tail;      // their NCO segments, in coordinates [0, os_hist)
    bool os_started = false;           // false: stream start, the history is silence
    hzsdr_nco nco{};
    // staging for the end-to-end path
    void *stage_in = nullptr;
    size_t stage_in_bytes = 0;
    void *stage_out = nullptr;
    size_t stage_out_bytes = 0;
    // pipelined end-to-end path: kPipeDepth staging slots, three streams
    static constexpr int kPipeDepth = 3;
    struct PipeSlot {
        void *in = nullptr, *out = nullptr;
        size_t in_bytes = 0, out_bytes = 0;
        cudaEvent_t in_done = nullptr, k_done = nullptr, out_done = nullptr;
        bool used = false;
    } pipe[kPipeDepth];
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    uint64_t submitted = 0;
};

extern "C" int hzsdr_chain_create(hzsdr_ctx *ctx, const hzsdr_chain_config *cfg, hzsdr_chain **out) {
    HZ_ENTER(ctx);
    if (!out || !cfg) return fail(HZSDR_ERR_INVALID, "hzsdr_chain_create: null argument");
    *out = nullptr;
    if (cfg->src_format != HZSDR_FORMAT_U8 && cfg->src_format != HZSDR_FORMAT_I8 && cfg->src_format != HZSDR_FORMAT_I16)
        return fail(HZSDR_ERR_FORMAT_UNKNOWN, "hzsdr_chain_create: raw source format expected, got %d", cfg->src_format);
    if (!fft_len_ok(cfg->n_fft))
        return fail(HZSDR_ERR_UNSUPPORTED, "hzsdr_chain_create: filter length %zu: need a power of two in [2, 16384]", cfg->n_fft);
    if (!cfg->filter_host) return fail(HZSDR_ERR_INVALID, "hzsdr_chain_create: null filter");
    if (cfg->decimate == 0 || cfg->sample_rate == 0) return fail(HZSDR_ERR_INVALID, "hzsdr_chain_create: decimate / sample_rate must be > 0");
    const uint32_t db = cfg->decimate_block ? cfg->decimate_block : (uint32_t)kDecimateBlock;
    if ((db & (db - 1)) != 0 || db > (1u << 24))
        return fail(HZSDR_ERR_UNSUPPORTED, "hzsdr_chain_create: decimate block %u must be a power of two <= 2^24", db);
    if (cfg->i16_lsb_bits < 0 || cfg->i16_lsb_bits > 16 || (cfg->i16_lsb_bits && cfg->src_format != HZSDR_FORMAT_I16))
        return fail(HZSDR_ERR_INVALID, "hzsdr_chain_create: i16_lsb_bits = %d", cfg->i16_lsb_bits);
    if (cfg->overlap_save_taps) {
        // the spectral fold behind the x16 decimation needs window starts on multiples of 16, the first-pass
        // shortcut for a silent history whole rows of 1024: os_hist = (taps - 1) rounded up to 1024
        if (cfg->n_fft != 16384 || cfg->decimate % 16 != 0 || db < 16384)
            return fail(HZSDR_ERR_UNSUPPORTED, "hzsdr_chain_create: the fused overlap-save form needs n_fft = 16384 and a decimation factor "
                        "that is a multiple of 16 (other shapes: hzsdr_fir_* on a converted stream)");
        if (cfg->overlap_save_taps > 8193)
            return fail(HZSDR_ERR_UNSUPPORTED, "hzsdr_chain_create: overlap-save with %u taps: at most 8193 (half a window of history)",
                        cfg->overlap_save_taps);
    }
    hzsdr_chain *c = new hzsdr_chain();
    c->ctx = ctx;
    c->cfg = *cfg;
    c->cfg.filter_host = nullptr;
    if (cfg->overlap_save_taps) {
        c->os_hist = ((cfg->overlap_save_taps - 1 + 1023) / 1024) * 1024;
        if (c->os_hist == 0) c->os_hist = 1024;
        c->os_hop = 16384 - c->os_hist;
    }
    c->decim_block = db;
    while ((1u << c->db_log2) < db) c->db_log2++;
    c->per_block = db / cfg->decimate;
    c->inv_d = nextafterf(1.0f / (float)cfg->decimate, 0.0f);
    c->nco.sample_rate = cfg->sample_rate;
    c->nco.ts = 0.0;
    int rc = get_twiddles(ctx, (int)cfg->n_fft, &c->tw);
    if (rc) {
        delete c;
        return rc;
    }
    cudaError_t e = cudaMalloc((void **)&c->H, sizeof(float2) * cfg->n_fft);
    if (e == cudaSuccess) e = cudaMemcpy(c->H, cfg->filter_host, sizeof(float2) * cfg->n_fft, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && cfg->n_fft == 1024 && db >= 1024) {
        std::vector<float2> t(kChain1024TableLen);
        chain1024_twiddles(t.data());
        e = cudaMalloc((void **)&c->tw1024, sizeof(float2) * t.size());
        if (e == cudaSuccess) e = cudaMemcpy(c->tw1024, t.data(), sizeof(float2) * t.size(), cudaMemcpyHostToDevice);
        c->can_split = (cfg->decimate % 2) == 0;  // (the table cache is allocated on first use: chain_split_table)
    }
    if (e == cudaSuccess && (cfg->n_fft == 2048 || cfg->n_fft == 4096 || cfg->n_fft == 8192) && db >= cfg->n_fft) {
        const int k = (int)(cfg->n_fft / 1024);
        std::vector<float2> t((size_t)(k - 1) * 1024 + cfg->n_fft);
        chaink_tables(k, reinterpret_cast<const float2 *>(cfg->filter_host), t.data(), t.data() + (size_t)(k - 1) * 1024);
        e = cudaMalloc((void **)&c->twk, sizeof(float2) * t.size());
        if (e == cudaSuccess) e = cudaMemcpy(c->twk, t.data(), sizeof(float2) * t.size(), cudaMemcpyHostToDevice);
        if (e == cudaSuccess && get_chain1024_tables(ctx, &c->tw1k_plain) != HZSDR_OK) e = cudaErrorMemoryAllocation;
        c->kfac = k;
    }
    if (e == cudaSuccess && cfg->n_fft == 16384 && cfg->decimate % 16 == 0 && db >= 16384) {
        std::vector<float2> t(31 * 32 + 15 * 1024 + 16384);
        chain16k_twiddles(t.data(), t.data() + 31 * 32);
        chain16k_permute_filter(reinterpret_cast<const float2 *>(cfg->filter_host), t.data() + 31 * 32 + 15 * 1024);
        e = cudaMalloc((void **)&c->tw16k, sizeof(float2) * t.size());
        if (e == cudaSuccess) e = cudaMemcpy(c->tw16k, t.data(), sizeof(float2) * t.size(), cudaMemcpyHostToDevice);
        if (e == cudaSuccess && c->os_hist) {
            const size_t hb = (size_t)c->os_hist * hzsdr_format_size(cfg->src_format);
            e = cudaMalloc((void **)&c->hist_raw, 2 * hb);
            if (e == cudaSuccess) e = cudaMemset(c->hist_raw, 0, 2 * hb);
        }
    }
    if (e != cudaSuccess) {
        if (c->hist_raw) cudaFree(c->hist_raw);
        if (c->H) cudaFree(c->H);
        if (c->tw1024) cudaFree(c->tw1024);
        if (c->tw16k) cudaFree(c->tw16k);
        if (c->twk) cudaFree(c->twk);
        if (c->split_dev) cudaFree(c->split_dev);
        if (c->split_stage) cudaFreeHost(c->split_stage);
        delete c;
        return fail(HZSDR_ERR_CUDA, "hzsdr_chain_create: %s", cudaGetErrorString(e));
    }
    *out = c;
    return HZSDR_OK;
}

extern "C" int hzsdr_chain_destroy(hzsdr_chain *c) {
    if (!c) return HZSDR_OK;
    HZ_ENTER(c->ctx);
    cudaStreamSynchronize(c->ctx->stream);
    if (c->H) cudaFree(c->H);
    if (c->hist_raw) cudaFree(c->hist_raw);
    if (c->tw1024) cudaFree(c->tw1024);
    if (c->tw16k) cudaFree(c->tw16k);
    if (c->twk) cudaFree(c->twk);
    if (c->split_dev) cudaFree(c->split_dev);
    if (c->split_stage) cudaFreeHost(c->split_stage);
    if (c->stage_in) cudaFree(c->stage_in);
    if (c->stage_out) cudaFree(c->stage_out);
    if (c->copy_in) { cudaStreamSynchronize(c->copy_in); cudaStreamDestroy(c->copy_in); }
    if (c->copy_out) { cudaStreamSynchronize(c->copy_out); cudaStreamDestroy(c->copy_out); }
    for (auto &sl : c->pipe) {
        if (sl.in) cudaFree(sl.in);
        if (sl.out) cudaFree(sl.out);
        if (sl.in_done) cudaEventDestroy(sl.in_done);
        if (sl.k_done) cudaEventDestroy(sl.k_done);
        if (sl.out_done) cudaEventDestroy(sl.out_done);
    }
    delete c;
    return HZSDR_OK;
}

// The split table (chain1024.cu, SPLIT) for the phase step of a launch's dominant segment: from the
// chain's cache, or built on the host and copied in stream order into a slot nothing in flight reads.
static int chain_split_table(hzsdr_chain *c, const NcoTable &table, const float2 **tw_out, uint64_t *dp_out) {
    uint64_t dp = 0;
    uint32_t longest = 0;
    for (int k = 0; k < table.count; k++)
        if (table.seg[k].count > longest && table.seg[k].dp) longest = table.seg[k].count, dp = table.seg[k].dp;
    if (!c->split_dev) {  // the channelizer's chains only get here for their stream-start buffers
        const size_t bytes = sizeof(float2) * 32 * 32 * hzsdr_chain::kSplitSlots;
        HZ_CUDA(cudaMalloc((void **)&c->split_dev, bytes));
        HZ_CUDA(cudaHostAlloc((void **)&c->split_stage, bytes, cudaHostAllocPortable));
    }
    int slot = 0;
    while (slot < c->split_used && c->split_dp[slot] != dp) slot++;
    if (slot == c->split_used) {
        if (c->split_used == hzsdr_chain::kSplitSlots) {  // cache full: start over once nothing reads it any more
            HZ_CUDA(cudaStreamSynchronize(c->ctx->stream));
            c->split_used = 0;
            slot = 0;
        }
        float2 *stage = c->split_stage + (size_t)slot * 1024;
        chain1024_split_twiddles(stage, dp, format_scale(c->cfg.src_format));
        HZ_CUDA(cudaMemcpyAsync(c->split_dev + (size_t)slot * 1024, stage, sizeof(float2) * 1024, cudaMemcpyHostToDevice,
                                c->ctx->stream));
        c->split_dp[slot] = dp;
        c->split_used++;
    }
    *tw_out = c->split_dev + (size_t)slot * 1024;
    *dp_out = dp;
    return HZSDR_OK;
}

static size_t chain_unit(const hzsdr_chain *c) {
    return c->cfg.n_fft > c->decim_block ? c->cfg.n_fft : c->decim_block;  // both powers of two: lcm = max
}

extern "C" int hzsdr_chain_out_len(const hzsdr_chain *c, size_t n, size_t *n_out) {
    if (!c || !n_out) return fail(HZSDR_ERR_INVALID, "hzsdr_chain_out_len: null");
    const size_t lz = (n / c->cfg.n_fft) * c->cfg.n_fft;  // ConvolutionReader drops the partial block
    *n_out = (lz / c->decim_block) * c->per_block;        // DecimateReader drops the partial block
    return HZSDR_OK;
}

// Overlap-save over one buffer (chain16k.cu, OS).  Windows are anchored at the start of the call: window w
// covers call samples [w*hop - os_hist, w*hop + os_hop), its first os_hist samples being history (the carried
// raw tail of the previous buffer for w = 0).  Launch coordinates put sample 0 os_hist samples in front of
// the launch's first new sample; a launch's NCO table covers its windows in those coordinates, the history
// part with the segments the previous call left behind.  Usually one launch per call.
static int chain_exec_os(hzsdr_chain *c, const void *src, size_t n, void *dst) {
    const size_t L = c->os_hop, hist = c->os_hist, N = 16384;
    const int sb = hzsdr_format_size(c->cfg.src_format);
    if (((uintptr_t)src % 16) != 0) return fail(HZSDR_ERR_INVALID, "hzsdr_chain_exec: overlap-save needs a 16-byte aligned source");
    std::vector<HostSeg> segs;
    double ts = c->nco.ts;
    build_segments(c->nco.sample_rate, n, &ts, segs);
    // extended coordinates: [0, hist) = history, [hist, hist + n) = this buffer
    std::vector<HostSeg> ext;
    if (c->os_started && !c->os_tail.empty())
        ext = c->os_tail;
    else
        ext.push_back(HostSeg{0, hist, 0.0, 0.0});  // silence in front of the stream: any phase will do
    for (HostSeg h : segs) {
        h.j0 += hist;
        ext.push_back(h);
    }
    const size_t W = (n + L - 1) / L;
    size_t wa = 0, k0 = 0;
    while (wa < W) {
        const size_t lo = wa * L;
        while (k0 < ext.size() && (size_t)(ext[k0].j0 + ext[k0].count) <= lo) k0++;
        size_t wb = wa, k = k0;
        while (wb < W) {
            const size_t hi = wb * L + N;
            size_t kk = k;
            while (kk < ext.size() && (size_t)ext[kk].j0 < hi) kk++;
            if (kk - k0 > (size_t)kMaxSegsPerLaunch) break;
            k = kk;
            wb++;
            if ((wb - wa) * L >= ((size_t)1 << 30)) break;
        }
        if (wb == wa)
            return fail(HZSDR_ERR_UNSUPPORTED, "NCO: more than %d accumulator segments inside one overlap-save window", kMaxSegsPerLaunch);
        const size_t hi = (wb - 1) * L + N;
        NcoTable table;
        table.count = 0;
        for (size_t q = k0; q < ext.size() && (size_t)ext[q].j0 < hi; q++) {
            HostSeg h = ext[q];
            const size_t h_end = (size_t)(h.j0 + h.count);
            if ((size_t)h.j0 < lo) {
                const size_t d = lo - (size_t)h.j0;
                h.base += (double)d * h.step;
                h.j0 = lo;
                h.count -= d;
            }
            if (h_end > hi) h.count = hi - (size_t)h.j0;
            table.seg[table.count++] = to_device_segment(h, lo, c->cfg.shift_hz);
        }
        ChainParams prm{};
        prm.src = (const uint8_t *)src + ((ptrdiff_t)lo - (ptrdiff_t)hist) * sb;  // launch coordinate 0
        prm.dst = (float2 *)dst;
        prm.nblocks = (uint32_t)(wb - wa);
        prm.z0 = (uint32_t)lo;
        prm.D = c->cfg.decimate;
        prm.M = c->per_block;
        prm.db_log2 = c->db_log2;
        prm.inv_d = c->inv_d;
        prm.lsb_shift = c->cfg.i16_lsb_bits ? 16 - c->cfg.i16_lsb_bits : 0;
        prm.tw = c->tw16k;
        prm.tw3 = c->tw16k + 31 * 32;
        prm.tw1k = c->tw16k + 31 * 32 + 15 * 1024;
        prm.hist = c->hist_raw + (size_t)c->hist_cur * hist * sb;
        // the call's last launch also saves the buffer's last os_hist raw samples for the next call
        prm.hist_out = wb == W ? c->hist_raw + (size_t)(c->hist_cur ^ 1) * hist * sb : nullptr;
        prm.tail_src = (const uint8_t *)src + (n - hist) * sb;
        prm.tail_bytes = (uint32_t)(hist * sb);
        prm.os_hop = (uint32_t)L;
        prm.os_head = wa == 0 ? (uint32_t)hist : 0u;
        prm.os_valid = (uint32_t)(hist + n - lo);
        prm.os_zend = (uint32_t)n;
        prm.os_zero_head = (wa == 0 && !c->os_started) ? 1u : 0u;
        int rc = launch_chain16k(c->ctx, c->cfg.src_format, prm, table);
        if (rc) return rc;
        wa = wb;
    }
    // carry the buffer's last os_hist raw samples and the segments that cover them
    c->hist_cur ^= 1;
    c->os_tail.clear();
    for (HostSeg h : segs) {
        const size_t h_end = (size_t)(h.j0 + h.count);
        if (h_end <= n - hist) continue;
        if ((size_t)h.j0 < n - hist) {
            const size_t d = (n - hist) - (size_t)h.j0;
            h.base += (double)d * h.step;
            h.j0 = n - hist;
            h.count -= d;
        }
        h.j0 -= (n - hist);
        c->os_tail.push_back(h);
    }
    c->os_started = true;
    c->nco.ts = ts;
    return HZSDR_OK;
}

extern "C" int hzsdr_chain_exec(hzsdr_chain *c, const void *src, size_t n, void *dst, size_t dst_len, size_t *n_out) {
    if (!c) return fail(HZSDR_ERR_INVALID, "hzsdr_chain_exec: null chain");
    HZ_ENTER(c->ctx);
    if (n_out) *n_out = 0;
    const size_t unit = chain_unit(c);
    if (n % unit)
        return fail(HZSDR_ERR_INVALID, "hzsdr_chain_exec: n = %zu must be a multiple of %zu (block boundaries are anchored at "
                    "stream start; buffer the remainder in the reader)", n, unit);
    if (n == 0) return HZSDR_OK;
    if (n > 0x7fffffffull) return fail(HZSDR_ERR_INVALID, "hzsdr_chain_exec: at most 2^31-1 samples per call");
    size_t total = 0;
    hzsdr_chain_out_len(c, n, &total);
    if (dst_len < total) return fail(HZSDR_ERR_DST_TOO_SMALL, "hzsdr_chain_exec: %zu < %zu", dst_len, total);
    if (!src || !dst) return fail(HZSDR_ERR_INVALID, "hzsdr_chain_exec: null buffer");
    const int sb = hzsdr_format_size(c->cfg.src_format);
    if (((uintptr_t)src % sb) || ((uintptr_t)dst % 8)) return fail(HZSDR_ERR_INVALID, "hzsdr_chain_exec: misaligned buffer");
    if (c->os_hop) {
        int rc = chain_exec_os(c, src, n, dst);
        if (rc) return rc;
        if (n_out) *n_out = total;
        return HZSDR_OK;
    }

    std::vector<HostSeg> segs;
    double ts = c->nco.ts;
    build_segments(c->nco.sample_rate, n, &ts, segs);
    std::vector<NcoLaunch> launches;
    int rc = plan_nco_launches(segs, n, c->cfg.n_fft, c->cfg.shift_hz, launches);
    if (rc) return rc;
    for (const NcoLaunch &L : launches) {
        ChainParams prm{};
        prm.src = (const uint8_t *)src + L.first * sb;
        prm.dst = (float2 *)dst;
        prm.tw = c->tw;
        prm.H = c->H;
        prm.nblocks = (uint32_t)(L.count / c->cfg.n_fft);
        prm.z0 = (uint32_t)L.first;
        prm.D = c->cfg.decimate;
        prm.M = c->per_block;
        prm.db_log2 = c->db_log2;
        prm.inv_d = c->inv_d;
        prm.lsb_shift = c->cfg.i16_lsb_bits ? 16 - c->cfg.i16_lsb_bits : 0;
        if (c->tw1024) {
            prm.tw = c->tw1024;
            if (c->can_split) {
                rc = chain_split_table(c, L.table, &prm.tw, &prm.dp_nom);
                if (rc) return rc;
                prm.tw_bc = c->tw1024 + 32 * 32;
                prm.split = 1;
            }
            rc = launch_chain1024(c->ctx, c->cfg.src_format, prm, L.table);
        } else if (c->twk) {
            prm.tw = c->tw1k_plain;
            prm.tw3 = c->twk;
            prm.tw1k = c->twk + (size_t)(c->kfac - 1) * 1024;
            rc = launch_chaink(c->ctx, c->cfg.src_format, c->kfac, prm, L.table);
        } else if (c->tw16k && ((uintptr_t)prm.src % 16) == 0) {
            prm.tw = c->tw16k;
            prm.tw3 = c->tw16k + 31 * 32;
            prm.tw1k = c->tw16k + 31 * 32 + 15 * 1024;
            rc = launch_chain16k(c->ctx, c->cfg.src_format, prm, L.table);
        } else {
            rc = dispatch_chain(c->ctx, c->cfg.n_fft, c->cfg.src_format, prm, L.table);
        }
        if (rc) return rc;
    }
    c->nco.ts = ts;
    if (n_out) *n_out = total;
    return HZSDR_OK;
}

// K consecutive buffers of the chain's stream in one call (what a reader that drains K ring slots does).
//
// N = 1024 chains with an even decimation factor: ONE launch of the batched kernel (chain1024.cu, BATCH + SPLIT)
// over all K buffers -- a buffer is a "stream" of the channelizer form whose NCO segments continue where the
// previous buffer's ended.  The table prologue, the launch and its tail are paid once per K buffers instead of
// once per buffer (a 2^22-sample buffer is only two blocks per warp), and the kernel's per-lane tables live in
// Tensor Memory.  Buffers that do not fit a descriptor (a stream-start buffer with dozens of accumulator
// segments) and other chain shapes go through hzsdr_chain_exec one by one, in order: the launches overlap on
// the device and the host loop costs ~2 us per buffer.
extern "C" int hzsdr_chain_exec_batch(hzsdr_chain *c, const void *const *srcs, size_t n_each, void *const *dsts,
                                      size_t dst_len_each, size_t count, size_t *n_out_each) {
    if (!c) return fail(HZSDR_ERR_INVALID, "hzsdr_chain_exec_batch: null chain");
    HZ_ENTER(c->ctx);
    if (n_out_each) *n_out_each = 0;
    if (count && (!srcs || !dsts)) return fail(HZSDR_ERR_INVALID, "hzsdr_chain_exec_batch: null buffer table");
    size_t total = 0;
    hzsdr_chain_out_len(c, n_each, &total);
    const size_t blocks_each = n_each / 1024;
    const bool batched = c->tw1024 && c->can_split && !c->os_hop && count >= 2 && n_each && n_each % chain_unit(c) == 0 &&
                         n_each <= 0x7fffffffull && dst_len_each >= total && count >= 8;  // (a few buffers: their own launches overlap just as well)
    if (!batched) {
        size_t got = 0;
        for (size_t k = 0; k < count; k++) {
            int rc = hzsdr_chain_exec(c, srcs[k], n_each, dsts[k], dst_len_each, &got);
            if (rc) return rc;
        }
        if (n_out_each) *n_out_each = got;
        return HZSDR_OK;
    }
    const int sb = hzsdr_format_size(c->cfg.src_format);
    {
        std::vector<BufSpan> all;
        all.reserve(2 * count);
        for (size_t k = 0; k < count; k++) {
            if (!srcs[k] || !dsts[k]) return fail(HZSDR_ERR_INVALID, "hzsdr_chain_exec_batch: null buffer %zu", k);
            if (((uintptr_t)srcs[k] % sb) || ((uintptr_t)dsts[k] % 8)) return fail(HZSDR_ERR_INVALID, "hzsdr_chain_exec_batch: misaligned buffer %zu", k);
            all.push_back({(uintptr_t)srcs[k], (uintptr_t)srcs[k] + n_each * (size_t)sb, false});
            all.push_back({(uintptr_t)dsts[k], (uintptr_t)dsts[k] + total * sizeof(float2), true});
        }
        // inside one kernel nothing is ordered: buffers that write what another buffer of the call reads or writes keep
        // the call's order by going one at a time
        if (write_conflict(std::move(all))) {
            size_t got = 0;
            for (size_t k = 0; k < count; k++) {
                int rc = hzsdr_chain_exec(c, srcs[k], n_each, dsts[k], dst_len_each, &got);
                if (rc) return rc;
            }
            if (n_out_each) *n_out_each = got;
            return HZSDR_OK;
        }
    }
    // Launches of up to kParamStreams buffers / kParamSegs segments, descriptors in the kernel parameters.  The
    // spans of every buffer go through the context's OverlapWindow like a single launch's, so consecutive batched
    // launches (and calls) overlap on the device when -- and only when -- they touch disjoint memory.
    ChainParams prm{};
    prm.tw = c->tw1024;  // [32 x 32 plain | twB | twC]: only twB / twC are read by the batched SPLIT kernel
    prm.H = c->H;
    prm.nblocks = (uint32_t)blocks_each;
    prm.z0 = 0;  // every buffer starts on a DecimateReader block boundary (n_each is a multiple of it)
    prm.D = c->cfg.decimate;
    prm.M = c->per_block;
    prm.db_log2 = c->db_log2;
    prm.inv_d = c->inv_d;
    prm.lsb_shift = c->cfg.i16_lsb_bits ? 16 - c->cfg.i16_lsb_bits : 0;
    prm.streams = nullptr;
    prm.seg_pool = nullptr;
    prm.split = 1;
    prm.tw_bc = nullptr;
    BatchTable tbl;
    uint32_t nb = 0, ns = 0;
    std::vector<BufSpan> pending;  // spans of the buffers in `tbl`
    auto flush = [&]() -> int {
        if (!nb) return HZSDR_OK;
        prm.nstreams = nb;
        const bool may = admit_spans(c->ctx, std::move(pending));
        pending.clear();
        const int rc2 = launch_chain1024_batch(c->ctx, c->cfg.src_format, prm, &tbl, may);
        nb = ns = 0;
        return rc2;
    };
    std::vector<HostSeg> segs;
    std::vector<NcoLaunch> launches;
    int rc = HZSDR_OK;
    for (size_t k = 0; k < count; k++) {
        double ts = c->nco.ts;
        build_segments(c->nco.sample_rate, n_each, &ts, segs);
        rc = plan_nco_launches(segs, n_each, c->cfg.n_fft, c->cfg.shift_hz, launches);
        if (rc) return rc;
        if (launches.size() == 1 && launches[0].table.count <= kParamSegs) {
            const NcoTable &t = launches[0].table;
            if (nb == (uint32_t)kParamStreams || ns + (uint32_t)t.count > (uint32_t)kParamSegs) {
                rc = flush();
                if (rc) return rc;
            }
            pending.push_back({(uintptr_t)srcs[k], (uintptr_t)srcs[k] + n_each * (size_t)sb, false});
            pending.push_back({(uintptr_t)dsts[k], (uintptr_t)dsts[k] + total * sizeof(float2), true});
            fill_desc(tbl.desc[nb], tbl.seg, ns, srcs[k], dsts[k], t);
            nb++;
            ns += (uint32_t)t.count;
            c->nco.ts = ts;
        } else {  // more accumulator segments than one table holds: in order; carries c->nco.ts itself
            rc = flush();
            if (rc) return rc;
            size_t got = 0;
            rc = hzsdr_chain_exec(c, srcs[k], n_each, dsts[k], dst_len_each, &got);
            if (rc) return rc;
        }
    }
    rc = flush();
    if (rc) return rc;
    if (n_out_each) *n_out_each = total;
    return HZSDR_OK;
}

extern "C" int hzsdr_chain_exec_host(hzsdr_chain *c, const void *src_host, size_t n, void *dst_host, size_t dst_len,
                                     size_t *n_out) {
    if (!c) return fail(HZSDR_ERR_INVALID, "hzsdr_chain_exec_host: null chain");
    HZ_ENTER(c->ctx);
    if (n_out) *n_out = 0;
    size_t total = 0;
    hzsdr_chain_out_len(c, n, &total);
    if (dst_len < total) return fail(HZSDR_ERR_DST_TOO_SMALL, "hzsdr_chain_exec_host: %zu < %zu", dst_len, total);
    const size_t in_bytes = n * hzsdr_format_size(c->cfg.src_format), out_bytes = total * 8;
    if (in_bytes > c->stage_in_bytes) {
        if (c->stage_in) cudaFree(c->stage_in);
        c->stage_in = nullptr;
        c->stage_in_bytes = 0;
        HZ_CUDA(cudaMalloc(&c->stage_in, in_bytes));
        c->stage_in_bytes = in_bytes;
    }
    if (out_bytes > c->stage_out_bytes) {
        if (c->stage_out) cudaFree(c->stage_out);
        c->stage_out = nullptr;
        c->stage_out_bytes = 0;
        HZ_CUDA(cudaMalloc(&c->stage_out, out_bytes ? out_bytes : 8));
        c->stage_out_bytes = out_bytes;
    }
    if (in_bytes) HZ_CUDA(cudaMemcpyAsync(c->stage_in, src_host, in_bytes, cudaMemcpyHostToDevice, c->ctx->stream));
    size_t got = 0;
    int rc = hzsdr_chain_exec(c, c->stage_in, n, c->stage_out, total, &got);
    if (rc) return rc;
    if (got) HZ_CUDA(cudaMemcpyAsync(dst_host, c->stage_out, got * 8, cudaMemcpyDeviceToHost, c->ctx->stream));
    HZ_CUDA(cudaStreamSynchronize(c->ctx->stream));
    if (n_out) *n_out = got;
    return HZSDR_OK;
}

static int chain_pipe_init(hzsdr_chain *c) {
    if (c->copy_in) return HZSDR_OK;
    HZ_CUDA(cudaStreamCreateWithFlags(&c->copy_in, cudaStreamNonBlocking));
    HZ_CUDA(cudaStreamCreateWithFlags(&c->copy_out, cudaStreamNonBlocking));
    for (auto &sl : c->pipe) {
        HZ_CUDA(cudaEventCreateWithFlags(&sl.in_done, cudaEventDisableTiming));
        HZ_CUDA(cudaEventCreateWithFlags(&sl.k_done, cudaEventDisableTiming));
        HZ_CUDA(cudaEventCreateWithFlags(&sl.out_done, cudaEventDisableTiming));
    }
    return HZSDR_OK;
}

extern "C" int hzsdr_chain_submit_host(hzsdr_chain *c, const void *src_host, size_t n, void *dst_host, size_t dst_len,
                                       size_t *n_out) {
    if (!c) return fail(HZSDR_ERR_INVALID, "hzsdr_chain_submit_host: null chain");
    HZ_ENTER(c->ctx);
    if (n_out) *n_out = 0;
    size_t total = 0;
    hzsdr_chain_out_len(c, n, &total);
    if (dst_len < total) return fail(HZSDR_ERR_DST_TOO_SMALL, "hzsdr_chain_submit_host: %zu < %zu", dst_len, total);
    if (n % chain_unit(c)) return fail(HZSDR_ERR_INVALID, "hzsdr_chain_submit_host: n = %zu must be a multiple of %zu", n, chain_unit(c));
    if (n == 0) return HZSDR_OK;
    int rc0 = chain_pipe_init(c);
    if (rc0) return rc0;
    auto &sl = c->pipe[c->submitted % hzsdr_chain::kPipeDepth];
    const size_t in_bytes = n * hzsdr_format_size(c->cfg.src_format), out_bytes = total * 8;
    if (in_bytes > sl.in_bytes || out_bytes > sl.out_bytes) {
        // growing a slot: drain whatever still uses it first (rare: first use or a larger buffer)
        HZ_CUDA(cudaStreamSynchronize(c->copy_in));
        HZ_CUDA(cudaStreamSynchronize(c->ctx->stream));
        HZ_CUDA(cudaStreamSynchronize(c->copy_out));
        if (in_bytes > sl.in_bytes) {
            if (sl.in) cudaFree(sl.in);
            sl.in = nullptr; sl.in_bytes = 0;
            HZ_CUDA(cudaMalloc(&sl.in, in_bytes));
            sl.in_bytes = in_bytes;
        }
        if (out_bytes > sl.out_bytes) {
            if (sl.out) cudaFree(sl.out);
            sl.out = nullptr; sl.out_bytes = 0;
            HZ_CUDA(cudaMalloc(&sl.out, out_bytes ? out_bytes : 8));
            sl.out_bytes = out_bytes;
        }
    }
    // H2D may start once the kernel that last read this slot's input is done
    if (sl.used) HZ_CUDA(cudaStreamWaitEvent(c->copy_in, sl.k_done, 0));
    HZ_CUDA(cudaMemcpyAsync(sl.in, src_host, in_bytes, cudaMemcpyHostToDevice, c->copy_in));
    HZ_CUDA(cudaEventRecord(sl.in_done, c->copy_in));
    // the kernel needs the input landed and the slot's previous result drained
    HZ_CUDA(cudaStreamWaitEvent(c->ctx->stream, sl.in_done, 0));
    if (sl.used) HZ_CUDA(cudaStreamWaitEvent(c->ctx->stream, sl.out_done, 0));
    size_t got = 0;
    int rc = hzsdr_chain_exec(c, sl.in, n, sl.out, total, &got);
    if (rc) return rc;
    HZ_CUDA(cudaEventRecord(sl.k_done, c->ctx->stream));
    HZ_CUDA(cudaStreamWaitEvent(c->copy_out, sl.k_done, 0));
    if (got) HZ_CUDA(cudaMemcpyAsync(dst_host, sl.out, got * 8, cudaMemcpyDeviceToHost, c->copy_out));
    HZ_CUDA(cudaEventRecord(sl.out_done, c->copy_out));
    sl.used = true;
    c->submitted++;
    if (n_out) *n_out = got;
    return HZSDR_OK;
}

// The reader side of the driver hand-off: the next unread slot of a pinned ring (raw samples a producer
// thread wrote with hzsdr_ring_write_peek / write_poke; their H2D copy is already in flight on the ring's copy
// stream) goes through the chain, and the result travels to dst_host behind the kernel.  Nothing waits on
// the host: events order copy -> kernel -> copy, the slot is released to the producer in stream order.
extern "C" int hzsdr_chain_submit_ring(hzsdr_chain *c, hzsdr_ring *ring, void *dst_host, size_t dst_len, size_t *n_out) {
    if (!c || !ring) return fail(HZSDR_ERR_INVALID, "hzsdr_chain_submit_ring: null");
    HZ_ENTER(c->ctx);
    if (n_out) *n_out = 0;
    const void *slot = nullptr;
    size_t n = 0;
    int rc = hzsdr_ring_read(ring, &slot, &n);  // HZSDR_ERR_RING_UNDERRUN when the producer is behind
    if (rc) return rc;
    size_t total = 0;
    hzsdr_chain_out_len(c, n, &total);
    if (dst_len < total) rc = fail(HZSDR_ERR_DST_TOO_SMALL, "hzsdr_chain_submit_ring: %zu < %zu", dst_len, total);
    if (!rc && n % chain_unit(c)) rc = fail(HZSDR_ERR_INVALID, "hzsdr_chain_submit_ring: slot holds %zu samples, not a multiple of %zu", n, chain_unit(c));
    if (!rc) rc = chain_pipe_init(c);
    if (rc || n == 0) {
        hzsdr_ring_read_done(ring);  // (the slot is dropped: the ring must not wedge on a bad buffer)
        return rc;
    }
    auto &sl = c->pipe[c->submitted % hzsdr_chain::kPipeDepth];
    const size_t out_bytes = total * 8;
    if (out_bytes > sl.out_bytes) {
        HZ_CUDA(cudaStreamSynchronize(c->ctx->stream));
        HZ_CUDA(cudaStreamSynchronize(c->copy_out));
        if (sl.out) cudaFree(sl.out);
        sl.out = nullptr;
        sl.out_bytes = 0;
        HZ_CUDA(cudaMalloc(&sl.out, out_bytes ? out_bytes : 8));
        sl.out_bytes = out_bytes;
    }
    if (sl.used) HZ_CUDA(cudaStreamWaitEvent(c->ctx->stream, sl.out_done, 0));  // the slot's previous result has left
    size_t got = 0;
    rc = hzsdr_chain_exec(c, slot, n, sl.out, total, &got);
    const int rc2 = hzsdr_ring_read_done(ring);  // "consumed" is recorded behind the kernel on the context's stream
    if (rc) return rc;
    if (rc2) return rc2;
    HZ_CUDA(cudaEventRecord(sl.k_done, c->ctx->stream));
    HZ_CUDA(cudaStreamWaitEvent(c->copy_out, sl.k_done, 0));
    if (got) HZ_CUDA(cudaMemcpyAsync(dst_host, sl.out, got * 8, cudaMemcpyDeviceToHost, c->copy_out));
    HZ_CUDA(cudaEventRecord(sl.out_done, c->copy_out));
    sl.used = true;
    c->submitted++;
    if (n_out) *n_out = got;
    return HZSDR_OK;
}

extern "C" int hzsdr_chain_wait_host(hzsdr_chain *c) {
    if (!c) return fail(HZSDR_ERR_INVALID, "hzsdr_chain_wait_host: null chain");
    HZ_ENTER(c->ctx);
    if (c->copy_in) HZ_CUDA(cudaStreamSynchronize(c->copy_in));
    HZ_CUDA(cudaStreamSynchronize(c->ctx->stream));
    if (c->copy_out) HZ_CUDA(cudaStreamSynchronize(c->copy_out));
    return HZSDR_OK;
}

extern "C" int hzsdr_chain_get_ts(const hzsdr_chain *c, double *ts) {
    if (!c || !ts) return fail(HZSDR_ERR_INVALID, "hzsdr_chain_get_ts: null");
    *ts = c->nco.ts;
    return HZSDR_OK;
}

extern "C" int hzsdr_chain_set_ts(hzsdr_chain *c, double ts) {
    if (!c) return fail(HZSDR_ERR_INVALID, "hzsdr_chain_set_ts: null");
    c->nco.ts = ts;
    // overlap-save: the carried history belongs to the old accumulator -- the stream restarts (silence in front)
    c->os_started = false;
    c->os_tail.clear();
    return HZSDR_OK;
}

// =================================================================================================
// C ABI: channelizer -- many independent streams through the fused chain, one launch per buffer set
// (BASELINE config 5).  Each stream is what the reference would build as its own reader chain
// (stream/convert.go:37, shifter.go:89, convolution.go:36, decimate.go:34); streams share format,
// rate, filter and decimation and differ in mixer frequency and carried NCO time.
// =================================================================================================
struct hzsdr_channelizer {
    hzsdr_ctx *ctx = nullptr;
    std::vector<hzsdr_chain *> chains;  // per-stream state (ts, shift) + the single-stream fall-back
    static constexpr int kStages = 4;   // rotating descriptor staging so back-to-back execs never collide
    BatchStaging batch;
};

extern "C" int hzsdr_channelizer_destroy(hzsdr_channelizer *z) {
    if (!z) return HZSDR_OK;
    HZ_ENTER(z->ctx);
    cudaStreamSynchronize(z->ctx->stream);
    for (auto *c : z->chains) hzsdr_chain_destroy(c);
    z->batch.release();
    delete z;
    return HZSDR_OK;
}

extern "C" int hzsdr_channelizer_create(hzsdr_ctx *ctx, const hzsdr_chain_config *cfg, const double *shift_hz,
                                        size_t n_streams, hzsdr_channelizer **out) {
    HZ_ENTER(ctx);
    if (!out || !cfg || !shift_hz || n_streams == 0) return fail(HZSDR_ERR_INVALID, "hzsdr_channelizer_create: bad arguments");
    *out = nullptr;
    if (cfg->overlap_save_taps) return fail(HZSDR_ERR_UNSUPPORTED, "hzsdr_channelizer_create: overlap-save chains are single-stream");
    hzsdr_channelizer *z = new hzsdr_channelizer();
    z->ctx = ctx;
    for (size_t s = 0; s < n_streams; s++) {
        hzsdr_chain_config c = *cfg;
        c.shift_hz = shift_hz[s];
        hzsdr_chain *ch = nullptr;
        int rc = hzsdr_chain_create(ctx, &c, &ch);
        if (rc) {
            hzsdr_channelizer_destroy(z);
            return rc;
        }
        z->chains.push_back(ch);
    }
    {
        int rc = z->batch.reserve(ctx, n_streams);
        if (rc) {
            hzsdr_channelizer_destroy(z);
            return rc;
        }
    }
    *out = z;
    return HZSDR_OK;
}

// streams [first, first + count) of the channelizer; srcs / dsts are indexed from `first`
static int channelizer_run(hzsdr_channelizer *z, size_t first, size_t count, const void *const *srcs, size_t n,
                           void *const *dsts, size_t dst_len, size_t *n_out_each);

extern "C" int hzsdr_channelizer_exec(hzsdr_channelizer *z, const void *const *srcs, size_t n, void *const *dsts,
                                      size_t dst_len, size_t *n_out_each) {
    if (!z) return fail(HZSDR_ERR_INVALID, "hzsdr_channelizer_exec: null");
    HZ_ENTER(z->ctx);
    if (n_out_each) *n_out_each = 0;
    if (!srcs || !dsts) return fail(HZSDR_ERR_INVALID, "hzsdr_channelizer_exec: null buffer table");
    return channelizer_run(z, 0, z->chains.size(), srcs, n, dsts, dst_len, n_out_each);
}

// End to end: srcs_host / dsts_host are arrays of n_streams HOST pointers (pinned for the copies to
// overlap).  The streams go through the context's staging pipe in groups: the raw samples of group
// g+1 cross PCIe while group g is in the kernel and group g-1's decimated output travels back.
extern "C" int hzsdr_channelizer_submit_host(hzsdr_channelizer *z, const void *const *srcs_host, size_t n,
                                             void *const *dsts_host, size_t dst_len, size_t *n_out_each) {
    if (!z) return fail(HZSDR_ERR_INVALID, "hzsdr_channelizer_submit_host: null");
    HZ_ENTER(z->ctx);
    if (n_out_each) *n_out_each = 0;
    if (!srcs_host || !dsts_host) return fail(HZSDR_ERR_INVALID, "hzsdr_channelizer_submit_host: null buffer table");
    hzsdr_chain *c0 = z->chains[0];
    const size_t unit = chain_unit(c0);
    if (n % unit) return fail(HZSDR_ERR_INVALID, "hzsdr_channelizer_submit_host: n = %zu must be a multiple of %zu", n, unit);
    size_t total = 0;
    hzsdr_chain_out_len(c0, n, &total);
    if (dst_len < total) return fail(HZSDR_ERR_DST_TOO_SMALL, "hzsdr_channelizer_submit_host: %zu < %zu", dst_len, total);
    if (n == 0) return HZSDR_OK;
    const size_t ns = z->chains.size();
    for (size_t s = 0; s < ns; s++)
        if (!srcs_host[s] || !dsts_host[s]) return fail(HZSDR_ERR_INVALID, "hzsdr_channelizer_submit_host: null buffer for stream %zu", s);
    hzsdr_ctx *ctx = z->ctx;
    HostPipe &hp = ctx->host_pipe;
    HZ_CUDA(hp.init());
    const size_t sb = (size_t)hzsdr_format_size(c0->cfg.src_format);
    const size_t in_each = n * sb, out_each = (total ? total : 1) * 8;
    // group size: about 64 MiB of raw samples per piece, at least 8 pieces when there are enough streams
    size_t group = ((size_t)64 << 20) / (in_each ? in_each : 1);
    if (group > (ns + 7) / 8) group = (ns + 7) / 8;
    if (group < 1) group = 1;
    std::vector<const void *> src_dev(group);
    std::vector<void *> dst_dev(group);
    for (size_t first = 0; first < ns; first += group) {
        const size_t cnt = ns - first < group ? ns - first : group;
        HostPipe::Slot *sl = nullptr;
        HZ_CUDA(hp.next(ctx->stream, group * in_each, group * out_each, &sl));
        for (size_t k = 0; k < cnt; k++) {
            src_dev[k] = (const uint8_t *)sl->in + k * in_each;
            dst_dev[k] = (uint8_t *)sl->out + k * out_each;
            HZ_CUDA(cudaMemcpyAsync((void *)src_dev[k], srcs_host[first + k], in_each, cudaMemcpyHostToDevice, hp.copy_in));
        }
        HZ_CUDA(hp.before_kernel(ctx->stream, *sl));
        size_t got = 0;
        int rc = channelizer_run(z, first, cnt, src_dev.data(), n, dst_dev.data(), total, &got);
        if (rc) return rc;
        HZ_CUDA(hp.after_kernel(ctx->stream, *sl));
        if (got)
            for (size_t k = 0; k < cnt; k++)
                HZ_CUDA(cudaMemcpyAsync(dsts_host[first + k], dst_dev[k], got * 8, cudaMemcpyDeviceToHost, hp.copy_out));
        HZ_CUDA(hp.done(*sl));
    }
    if (n_out_each) *n_out_each = total;
    return HZSDR_OK;
}

static int channelizer_run(hzsdr_channelizer *z, size_t first, size_t count, const void *const *srcs, size_t n,
                           void *const *dsts, size_t dst_len, size_t *n_out_each) {
    hzsdr_chain *c0 = z->chains[0];
    const size_t unit = chain_unit(c0);
    if (n % unit) return fail(HZSDR_ERR_INVALID, "hzsdr_channelizer_exec: n = %zu must be a multiple of %zu", n, unit);
    size_t total = 0;
    hzsdr_chain_out_len(c0, n, &total);
    if (dst_len < total) return fail(HZSDR_ERR_DST_TOO_SMALL, "hzsdr_channelizer_exec: %zu < %zu", dst_len, total);
    if (n == 0) return HZSDR_OK;
    if (n > 0x7fffffffull) return fail(HZSDR_ERR_INVALID, "hzsdr_channelizer_exec: at most 2^31-1 samples per stream per call");

    int rc0 = z->batch.begin();
    if (rc0) return rc0;
    std::vector<HostSeg> segs;
    std::vector<NcoLaunch> launches;
    std::vector<size_t> singles;
    const bool batchable = c0->tw1024 != nullptr;  // the batched kernel is the N = 1024 one
    for (size_t s = 0; s < count; s++) {
        hzsdr_chain *c = z->chains[first + s];
        if (!srcs[s] || !dsts[s]) return fail(HZSDR_ERR_INVALID, "hzsdr_channelizer_exec: null buffer for stream %zu", first + s);
        bool ok = false;
        if (batchable) {
            double ts = c->nco.ts;
            build_segments(c->nco.sample_rate, n, &ts, segs);
            int rc = plan_nco_launches(segs, n, c->cfg.n_fft, c->cfg.shift_hz, launches);
            if (rc) return rc;
            if (launches.size() == 1) {
                z->batch.push(srcs[s], dsts[s], launches[0].table);
                c->nco.ts = ts;
                ok = true;
            }
        }
        if (!ok) singles.push_back(s);
    }
    // streams whose buffer needs more than one segment table, or another FFT length
    for (size_t s : singles) {
        size_t got = 0;
        int rc = hzsdr_chain_exec(z->chains[first + s], srcs[s], n, dsts[s], dst_len, &got);
        if (rc) return rc;
    }
    if (z->batch.nbatch) {
        const uint32_t nbatch = (uint32_t)z->batch.nbatch;
        ChainParams prm{};
        const float2 *plain = nullptr;  // the chains' own tables may be in split form
        int rc = get_chain1024_tables(z->ctx, &plain);
        if (rc) return rc;
        prm.tw = plain;
        prm.H = c0->H;
        prm.nblocks = (uint32_t)(n / c0->cfg.n_fft);
        prm.z0 = 0;
        prm.D = c0->cfg.decimate;
        prm.M = c0->per_block;
        prm.db_log2 = c0->db_log2;
        prm.inv_d = c0->inv_d;
        prm.lsb_shift = c0->cfg.i16_lsb_bits ? 16 - c0->cfg.i16_lsb_bits : 0;
        prm.split = c0->can_split ? 1 : 0;
        prm.tw_bc = nullptr;  // twB / twC follow the plain table
        // Steady state (a few accumulator segments per stream): launches of <= 64 streams whose descriptors travel in
        // the kernel parameters -- no copy in the stream, and a launch may overlap its predecessor when every stream's
        // spans are clear of what is still in flight.  Heavy tables (every stream at its start: ~90 segments each):
        // one launch, descriptors through device memory.
        const bool in_params = c0->can_split && z->batch.nsegs <= (size_t)nbatch * 6;
        if (in_params) {
            const StreamDesc *hd = z->batch.descs();
            const NcoSegment *pool = reinterpret_cast<const NcoSegment *>(z->batch.host[z->batch.stage] + sizeof(StreamDesc) * z->batch.cap);
            const size_t sbytes = (size_t)hzsdr_format_size(c0->cfg.src_format);
            {   // inside one kernel nothing is ordered (and the chunks below may overlap): no stream may write what another
                // reads or writes (batch_host.h)
                std::vector<BufSpan> spans;
                spans.reserve(2 * (size_t)nbatch);
                for (uint32_t k = 0; k < nbatch; k++) {
                    spans.push_back({(uintptr_t)hd[k].src, (uintptr_t)hd[k].src + n * sbytes, false});
                    spans.push_back({(uintptr_t)hd[k].dst, (uintptr_t)hd[k].dst + total * sizeof(float2), true});
                }
                if (write_conflict(std::move(spans)))
                    return fail(HZSDR_ERR_INVALID, "hzsdr_channelizer_exec: a stream's output buffer overlaps another stream's buffer");
            }
            BatchTable tbl;
            std::vector<int> seg_count(nbatch);
            for (uint32_t k = 0; k < nbatch; k++) seg_count[k] = hd[k].count;
            for (const auto &chunk : plan_param_launches(seg_count, (uint32_t)kParamStreams, (uint32_t)kParamSegs)) {
                uint32_t ns = 0;
                std::vector<BufSpan> spans;
                spans.reserve(2 * (size_t)chunk.second);
                for (uint32_t i = 0; i < chunk.second; i++) {
                    const StreamDesc &d = hd[chunk.first + i];
                    tbl.desc[i] = d;
                    tbl.desc[i].seg_off = ns;
                    for (int q = 0; q < d.count; q++) tbl.seg[ns + q] = pool[d.seg_off + q];
                    ns += (uint32_t)d.count;
                    spans.push_back({(uintptr_t)d.src, (uintptr_t)d.src + n * sbytes, false});
                    spans.push_back({(uintptr_t)d.dst, (uintptr_t)d.dst + total * sizeof(float2), true});
                }
                const bool may = admit_spans(z->ctx, std::move(spans));
                prm.streams = nullptr;
                prm.seg_pool = nullptr;
                prm.nstreams = chunk.second;
                rc = launch_chain1024_batch(z->ctx, c0->cfg.src_format, prm, &tbl, may);
                if (rc) return rc;
            }
        } else {
            rc = z->batch.upload(z->ctx->stream);
            if (rc) return rc;
            prm.streams = z->batch.dev_descs();
            prm.seg_pool = z->batch.dev_pool();
            prm.nstreams = nbatch;
            rc = launch_chain1024_batch(z->ctx, c0->cfg.src_format, prm, nullptr, false);
            if (rc) return rc;
        }
    }
    if (n_out_each) *n_out_each = total;
    return HZSDR_OK;
}

extern "C" int hzsdr_channelizer_get_ts(const hzsdr_channelizer *z, double *ts_out) {
    if (!z || !ts_out) return fail(HZSDR_ERR_INVALID, "hzsdr_channelizer_get_ts: null");
    for (size_t s = 0; s < z->chains.size(); s++) ts_out[s] = z->chains[s]->nco.ts;
    return HZSDR_OK;
}

extern "C" int hzsdr_channelizer_set_ts(hzsdr_channelizer *z, const double *ts) {
    if (!z || !ts) return fail(HZSDR_ERR_INVALID, "hzsdr_channelizer_set_ts: null");
    for (size_t s = 0; s < z->chains.size(); s++) z->chains[s]->nco.ts = ts[s];
    return HZSDR_OK;
}
