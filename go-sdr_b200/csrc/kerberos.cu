// kerberos.cu -- the coherent-receiver helpers of rtl/kerberos/internal:
//   FFTShiftAndScale (reader.go:57-64), GraftReaders' per-buffer work (graft.go:96-125),
//   checkAlignment's peak search (align.go:125-146) and PhaseOffsets (align.go:244-272).
// The transforms themselves are the planner's (fft.cu / bigfft.cu); the reference plans 65536 points.
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "fft.cuh"
#include "fft_kernels.cuh"

namespace hz {

// data[i], data[half+i] = data[half+i]/scale, data[i]/scale   (IEEE fp32 division, like the reference)
__global__ void __launch_bounds__(256) k_fftshift_scale(float2 *data, uint32_t n, uint32_t batch, float scale) {
    const uint32_t half = n / 2;
    const size_t total = (size_t)half * batch, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += stride) {
        const size_t b = w / half, i = w - b * half;
        float2 *v = data + b * n;
        const float2 lo = v[i], hi = v[half + i];
        v[i] = make_float2(__fdiv_rn(hi.x, scale), __fdiv_rn(hi.y, scale));
        v[half + i] = make_float2(__fdiv_rn(lo.x, scale), __fdiv_rn(lo.y, scale));
    }
}

// checkAlignment's search: the FIRST index with the largest power, exact zeros skipped, power
// computed in fp32 without contraction (align.go:133-140); then wrapped into (-n/2, n/2].
__global__ void __launch_bounds__(256) k_correlate_peak(const float2 *__restrict__ cc, uint32_t n, int32_t *__restrict__ out) {
    const float2 *v = cc + (size_t)blockIdx.x * n;
    float best = -INFINITY;
    uint32_t best_i = 0xffffffffu;
    for (uint32_t i = threadIdx.x; i < n; i += 256) {
        const float2 el = v[i];
        if (el.x == 0.0f && el.y == 0.0f) continue;
        const float pw = __fadd_rn(__fmul_rn(el.x, el.x), __fmul_rn(el.y, el.y));
        if (pw > best) {  // ascending i within a thread: strict > keeps the first
            best = pw;
            best_i = i;
        }
    }
    __shared__ float s_pw[256];
    __shared__ uint32_t s_i[256];
    s_pw[threadIdx.x] = best;
    s_i[threadIdx.x] = best_i;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            const float op = s_pw[threadIdx.x + s];
            const uint32_t oi = s_i[threadIdx.x + s];
            if (op > s_pw[threadIdx.x] || (op == s_pw[threadIdx.x] && oi < s_i[threadIdx.x])) {
                s_pw[threadIdx.x] = op;
                s_i[threadIdx.x] = oi;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        int32_t r = s_i[0] == 0xffffffffu ? -1 : (int32_t)s_i[0];  // maxPowI stays -1 when every element is zero
        if (r > (int32_t)(n / 2)) r -= (int32_t)n;
        out[blockIdx.x] = r;
    }
}

// PhaseOffsets: sum over i of Phase(complex128(conjMult(b0[i], bj[i]))), fp64 (align.go:258-264)
__global__ void __launch_bounds__(256) k_phase_sum(const float2 *__restrict__ bufs, uint32_t n, double *__restrict__ out) {
    const float2 *b0 = bufs, *bj = bufs + (size_t)blockIdx.x * n;
    double acc = 0.0;
    for (uint32_t i = threadIdx.x; i < n; i += 256) {
        const float2 a = b0[i], b = bj[i];
        const float2 m = go_cmul(a, make_float2(b.x, -b.y));
        acc += atan2((double)m.y, (double)m.x);
    }
    __shared__ double s[256];
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if (threadIdx.x < k) s[threadIdx.x] += s[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = s[0];
}

}  // namespace hz

using namespace hz;

extern "C" int hzsdr_fftshift_scale(hzsdr_ctx *ctx, void *data, size_t n, size_t batch, float scale) {
    HZ_ENTER(ctx);
    if (n == 0 || batch == 0) return HZSDR_OK;
    if (!data) return fail(HZSDR_ERR_INVALID, "hzsdr_fftshift_scale: null buffer");
    if (n > 0xffffffffull || batch > 0xffffffffull) return fail(HZSDR_ERR_INVALID, "hzsdr_fftshift_scale: too large");
    const size_t total = (n / 2) * batch;
    if (total == 0) return HZSDR_OK;
    k_fftshift_scale<<<(int)std::min<size_t>((total + 255) / 256, (size_t)ctx->sm_count * 8), 256, 0, ctx->stream>>>(
        (float2 *)data, (uint32_t)n, (uint32_t)batch, scale);
    HZ_CHECK_LAUNCH();
    return HZSDR_OK;
}

extern "C" int hzsdr_graft(hzsdr_ctx *ctx, const void *iq, size_t n_readers, size_t fft_size, void *dst, void *freq) {
    HZ_ENTER(ctx);
    if (n_readers == 0 || fft_size == 0) return fail(HZSDR_ERR_INVALID, "hzsdr_graft: empty input");
    if (!iq || !dst || !freq) return fail(HZSDR_ERR_INVALID, "hzsdr_graft: null buffer");
    if (freq == iq || freq == dst) return fail(HZSDR_ERR_INVALID, "hzsdr_graft: freq must be a separate buffer");
    // one forward plan per reader into its slice of freqBuf (graft.go:69-78) == a batched transform
    int rc = fft_any(ctx, fft_size, HZSDR_FFT_FORWARD, (const float2 *)iq, (float2 *)freq, n_readers);
    if (rc) return rc;
    rc = hzsdr_fftshift_scale(ctx, freq, fft_size, n_readers, (float)fft_size);  // graft.go:112
    if (rc) return rc;
    return fft_any(ctx, fft_size * n_readers, HZSDR_FFT_BACKWARD, (const float2 *)freq, (float2 *)dst, 1);  // graft.go:80,117
}

extern "C" int hzsdr_correlate_peak(hzsdr_ctx *ctx, const void *cc, size_t n, size_t batch, int32_t *offsets_host) {
    HZ_ENTER(ctx);
    if (batch == 0) return HZSDR_OK;
    if (!cc || !offsets_host || n == 0) return fail(HZSDR_ERR_INVALID, "hzsdr_correlate_peak: null / empty argument");
    if (n > 0x7fffffffull || batch > 65535) return fail(HZSDR_ERR_INVALID, "hzsdr_correlate_peak: too large");
    void *ws = nullptr;
    int rc = ctx_workspace(ctx, batch * sizeof(int32_t), &ws);
    if (rc) return rc;
    k_correlate_peak<<<(int)batch, 256, 0, ctx->stream>>>((const float2 *)cc, (uint32_t)n, (int32_t *)ws);
    HZ_CHECK_LAUNCH();
    HZ_CUDA(cudaMemcpyAsync(offsets_host, ws, batch * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    HZ_CUDA(cudaStreamSynchronize(ctx->stream));
    return HZSDR_OK;
}

extern "C" int hzsdr_phase_offsets(hzsdr_ctx *ctx, const void *bufs, size_t n_chan, size_t n, float *out_host) {
    HZ_ENTER(ctx);
    if (n_chan == 0) return HZSDR_OK;
    if (!bufs || !out_host || n == 0) return fail(HZSDR_ERR_INVALID, "hzsdr_phase_offsets: null / empty argument");
    if (n > 0xffffffffull || n_chan > 65535) return fail(HZSDR_ERR_INVALID, "hzsdr_phase_offsets: too large");
    void *ws = nullptr;
    int rc = ctx_workspace(ctx, n_chan * sizeof(double), &ws);
    if (rc) return rc;
    k_phase_sum<<<(int)n_chan, 256, 0, ctx->stream>>>((const float2 *)bufs, (uint32_t)n, (double *)ws);
    HZ_CHECK_LAUNCH();
    std::vector<double> ph(n_chan);
    HZ_CUDA(cudaMemcpyAsync(ph.data(), ws, n_chan * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    HZ_CUDA(cudaStreamSynchronize(ctx->stream));
    ph[0] = 1;  // align.go:265 -- the reference sets the reference channel's *sum* to 1 before averaging
    for (size_t j = 0; j < n_chan; j++) {
        const double a = ph[j] / (double)n;
        out_host[2 * j] = (float)cos(a);  // complex64(cmplx.Rect(1, a))
        out_host[2 * j + 1] = (float)sin(a);
    }
    return HZSDR_OK;
}
