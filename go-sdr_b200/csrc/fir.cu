// fir.cu -- true (linear) FIR convolution + decimation of a complex64 stream.
//
// AN EXTENSION, not a parity item: the reference's stream.ConvolutionReader is block-circular with
// no history (stream/convolution.go:57-81); BASELINE's "overlap-save FIR" and "polyphase FIR +
// decimate" have no reference counterpart (SURVEY.md finding 3).  Definition implemented here:
//     z[n] = sum_{k < taps} h[k] * y[n-k]      over the whole stream, y[n < 0] = 0
//     out  = z[D*i]                            continuous decimation phase
// checked against a direct fp64 FIR (tests/test_gpu_parity.py::test_fir_*).  History (taps-1
// samples) and the stream position are carried between calls.
//
//   overlap-save  windows of N samples every L = N-(taps-1) samples go through the fused
//                 FFT -> xH -> IFFT kernel (k_convolve with a source stride, no window copy); a
//                 decimating gather keeps the valid part.
//   polyphase     direct form, only the kept outputs are computed: one warp per output, lanes
//                 split the taps (conflict-free shared-memory reads), shuffle reduction.
#include <vector>

#include "common.cuh"

namespace hz {
int convolve_windows(hzsdr_ctx *ctx, const float2 *src, float2 *dst, const float2 *filter, size_t n_fft, size_t n_windows,
                     size_t src_stride);

// dst[o] = W[b*N + m], i = i0 + o*D, b = i / L, m = i - b*L + (taps-1)
__global__ void __launch_bounds__(256) k_os_gather(const float2 *__restrict__ W, float2 *__restrict__ dst, size_t cnt, size_t i0,
                                                    uint32_t D, uint32_t L, uint32_t N, uint32_t hist) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < cnt; o += stride) {
        const size_t i = i0 + o * D;
        const size_t b = i / L;
        const size_t m = i - b * L + hist;
        dst[o] = W[b * N + m];
    }
}

// one warp per output: z = sum_k h[k] * ext[hist + i - k]
constexpr int kPolyWarps = 8;
__global__ void __launch_bounds__(32 * kPolyWarps) k_polyphase(const float2 *__restrict__ ext, const float2 *__restrict__ taps,
                                                                float2 *__restrict__ dst, size_t cnt, size_t i0, uint32_t D,
                                                                uint32_t ntaps) {
    extern __shared__ float2 sh[];  // taps, reversed: sh[k'] = h[ntaps-1-k'] so that both reads ascend
    for (uint32_t k = threadIdx.x; k < ntaps; k += blockDim.x) sh[k] = taps[ntaps - 1 - k];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (size_t o = (size_t)blockIdx.x * kPolyWarps + warp; o < cnt; o += (size_t)gridDim.x * kPolyWarps) {
        // ext index of y[i - (ntaps-1)] is exactly i (ext carries ntaps-1 samples of history in front)
        const float2 *x = ext + (i0 + o * D);
        float2 acc = make_float2(0.f, 0.f);
        for (uint32_t k = lane; k < ntaps; k += 32) {
            const float2 h = sh[k], v = __ldg(x + k);
            acc.x = fmaf(h.x, v.x, acc.x);
            acc.x = fmaf(-h.y, v.y, acc.x);
            acc.y = fmaf(h.x, v.y, acc.y);
            acc.y = fmaf(h.y, v.x, acc.y);
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, s);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, s);
        }
        if (lane == 0) dst[o] = acc;
    }
}

}  // namespace hz

using namespace hz;

struct hzsdr_fir {
    hzsdr_ctx *ctx = nullptr;
    size_t ntaps = 0, hist = 0;
    unsigned D = 1;
    int method = 0;
    size_t n_fft = 0, L = 0;
    float2 *taps = nullptr;  // device, time domain
    float2 *H = nullptr;     // device, FFT_N(taps)/N
    float2 *ext = nullptr;   // [history | new samples | zero pad]
    size_t ext_cap = 0;
    float2 *win = nullptr;   // overlap-save window outputs
    size_t win_cap = 0;
    float2 *history = nullptr;  // hist samples
    uint64_t pos = 0;           // stream samples consumed so far
};

extern "C" int hzsdr_fir_destroy(hzsdr_fir *f) {
    if (!f) return HZSDR_OK;
    HZ_ENTER(f->ctx);
    cudaStreamSynchronize(f->ctx->stream);
    for (float2 *p : {f->taps, f->H, f->ext, f->win, f->history})
        if (p) cudaFree(p);
    delete f;
    return HZSDR_OK;
}

extern "C" int hzsdr_fir_create(hzsdr_ctx *ctx, const float *taps, size_t ntaps, unsigned decimate, int method, hzsdr_fir **out) {
    HZ_ENTER(ctx);
    if (!out || !taps || ntaps == 0 || decimate == 0) return fail(HZSDR_ERR_INVALID, "hzsdr_fir_create: bad arguments");
    *out = nullptr;
    if (method != HZSDR_FIR_AUTO && method != HZSDR_FIR_OVERLAP_SAVE && method != HZSDR_FIR_POLYPHASE)
        return fail(HZSDR_ERR_INVALID, "hzsdr_fir_create: method %d", method);
    size_t n_fft = 64;
    while (n_fft < 4 * ntaps && n_fft < 16384) n_fft <<= 1;  // >= 75% of every window is new output
    if (method == HZSDR_FIR_AUTO) method = (4.0 * ntaps / decimate < 40.0 || ntaps - 1 > n_fft / 2) ? HZSDR_FIR_POLYPHASE : HZSDR_FIR_OVERLAP_SAVE;
    if (method == HZSDR_FIR_OVERLAP_SAVE && ntaps - 1 > n_fft / 2)
        return fail(HZSDR_ERR_UNSUPPORTED, "hzsdr_fir_create: %zu taps need an FFT longer than 16384 for overlap-save", ntaps);
    if (method == HZSDR_FIR_POLYPHASE && ntaps * sizeof(float2) > 200 * 1024)
        return fail(HZSDR_ERR_UNSUPPORTED, "hzsdr_fir_create: %zu taps do not fit shared memory", ntaps);
    hzsdr_fir *f = new hzsdr_fir();
    f->ctx = ctx;
    f->ntaps = ntaps;
    f->hist = ntaps - 1;
    f->D = decimate;
    f->method = method;
    f->n_fft = n_fft;
    f->L = n_fft - (ntaps - 1);
    auto bail = [&](int rc) {
        hzsdr_fir_destroy(f);
        return rc;
    };
    cudaError_t e = cudaMalloc((void **)&f->taps, sizeof(float2) * ntaps);
    if (e == cudaSuccess) e = cudaMemcpy(f->taps, taps, sizeof(float2) * ntaps, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc((void **)&f->history, sizeof(float2) * (f->hist ? f->hist : 1));
    if (e == cudaSuccess) e = cudaMemset(f->history, 0, sizeof(float2) * (f->hist ? f->hist : 1));
    if (e != cudaSuccess) return bail(fail(HZSDR_ERR_CUDA, "hzsdr_fir_create: %s", cudaGetErrorString(e)));
    if (method == HZSDR_FIR_OVERLAP_SAVE) {
        // H = FFT_N(zero-padded taps) / N, computed with the library's own transform
        e = cudaMalloc((void **)&f->H, sizeof(float2) * n_fft);
        if (e == cudaSuccess) e = cudaMemset(f->H, 0, sizeof(float2) * n_fft);
        if (e == cudaSuccess) e = cudaMemcpy(f->H, taps, sizeof(float2) * ntaps, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return bail(fail(HZSDR_ERR_CUDA, "hzsdr_fir_create: %s", cudaGetErrorString(e)));
        hzsdr_fft_plan *p = nullptr;
        int rc = hzsdr_fft_plan_create(ctx, n_fft, n_fft, HZSDR_FFT_FORWARD, &p);
        if (rc == HZSDR_OK) rc = hzsdr_fft_exec(p, f->H, f->H, 1);
        if (rc == HZSDR_OK) rc = hzsdr_scale(ctx, f->H, n_fft, 1.0f / (float)n_fft);
        hzsdr_fft_plan_destroy(p);
        if (rc == HZSDR_OK) rc = hzsdr_ctx_sync(ctx);
        if (rc != HZSDR_OK) return bail(rc);
    }
    *out = f;
    return HZSDR_OK;
}

extern "C" int hzsdr_fir_reset(hzsdr_fir *f) {
    if (!f) return fail(HZSDR_ERR_INVALID, "hzsdr_fir_reset: null");
    HZ_ENTER(f->ctx);
    if (f->hist) HZ_CUDA(cudaMemsetAsync(f->history, 0, sizeof(float2) * f->hist, f->ctx->stream));
    f->pos = 0;
    return HZSDR_OK;
}

extern "C" int hzsdr_fir_exec(hzsdr_fir *f, const void *src, size_t n, void *dst, size_t dst_len, size_t *n_out) {
    if (!f) return fail(HZSDR_ERR_INVALID, "hzsdr_fir_exec: null");
    HZ_ENTER(f->ctx);
    if (n_out) *n_out = 0;
    if (n == 0) return HZSDR_OK;
    if (!src || !dst) return fail(HZSDR_ERR_INVALID, "hzsdr_fir_exec: null buffer");
    // kept outputs: stream indices g = D*i inside [pos, pos + n)
    const uint64_t g0 = (f->pos + f->D - 1) / f->D * f->D;
    const size_t cnt = g0 < f->pos + n ? (size_t)((f->pos + n - 1 - g0) / f->D + 1) : 0;
    if (dst_len < cnt) return fail(HZSDR_ERR_DST_TOO_SMALL, "hzsdr_fir_exec: %zu < %zu", dst_len, cnt);
    cudaStream_t st = f->ctx->stream;
    const size_t n_win = f->method == HZSDR_FIR_OVERLAP_SAVE ? (n + f->L - 1) / f->L : 0;
    const size_t ext_need = f->method == HZSDR_FIR_OVERLAP_SAVE ? (n_win - 1) * f->L + f->n_fft : f->hist + n;
    if (ext_need > f->ext_cap) {
        HZ_CUDA(cudaStreamSynchronize(st));
        if (f->ext) cudaFree(f->ext);
        f->ext = nullptr;
        f->ext_cap = 0;
        HZ_CUDA(cudaMalloc((void **)&f->ext, sizeof(float2) * ext_need));
        f->ext_cap = ext_need;
    }
    // ext = [history | new samples | zeros]
    if (f->hist) HZ_CUDA(cudaMemcpyAsync(f->ext, f->history, sizeof(float2) * f->hist, cudaMemcpyDeviceToDevice, st));
    HZ_CUDA(cudaMemcpyAsync(f->ext + f->hist, src, sizeof(float2) * n, cudaMemcpyDeviceToDevice, st));
    if (ext_need > f->hist + n) HZ_CUDA(cudaMemsetAsync(f->ext + f->hist + n, 0, sizeof(float2) * (ext_need - f->hist - n), st));
    const size_t i0 = (size_t)(g0 - f->pos);
    if (cnt) {
        if (f->method == HZSDR_FIR_OVERLAP_SAVE) {
            if (n_win * f->n_fft > f->win_cap) {
                HZ_CUDA(cudaStreamSynchronize(st));
                if (f->win) cudaFree(f->win);
                f->win = nullptr;
                f->win_cap = 0;
                HZ_CUDA(cudaMalloc((void **)&f->win, sizeof(float2) * n_win * f->n_fft));
                f->win_cap = n_win * f->n_fft;
            }
            int rc = convolve_windows(f->ctx, f->ext, f->win, f->H, f->n_fft, n_win, f->L);
            if (rc) return rc;
            const int grid = (int)std::min<size_t>((cnt + 255) / 256, (size_t)f->ctx->sm_count * 8);
            k_os_gather<<<grid, 256, 0, st>>>(f->win, (float2 *)dst, cnt, i0, f->D, (uint32_t)f->L, (uint32_t)f->n_fft, (uint32_t)f->hist);
            HZ_CHECK_LAUNCH();
        } else {
            const size_t smem = sizeof(float2) * f->ntaps;
            if (smem > 48 * 1024) HZ_CUDA(cudaFuncSetAttribute((const void *)k_polyphase, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const int grid = (int)std::min<size_t>((cnt + kPolyWarps - 1) / kPolyWarps, (size_t)f->ctx->sm_count * 8);
            k_polyphase<<<grid, 32 * kPolyWarps, smem, st>>>(f->ext, f->taps, (float2 *)dst, cnt, i0, f->D, (uint32_t)f->ntaps);
            HZ_CHECK_LAUNCH();
        }
    }
    // carry the last taps-1 samples
    if (f->hist) HZ_CUDA(cudaMemcpyAsync(f->history, f->ext + n, sizeof(float2) * f->hist, cudaMemcpyDeviceToDevice, st));
    f->pos += n;
    if (n_out) *n_out = cnt;
    return HZSDR_OK;
}
