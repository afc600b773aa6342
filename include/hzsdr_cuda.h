/*
 * hzsdr_cuda.h -- C ABI of libhzsdrcuda.so: hz.tools/sdr's IQ sample chain on NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary.  The Go package `cuda/` (go-sdr_b200/go/cuda, cgo) and the
 * `sdr.cuda`-tagged twins of conv.go and the stream package bind exactly these symbols; the C++ mirror of
 * the reader API (go-sdr_b200/host) and the Python ctypes binding used by tests/ and bench.py
 * call the same ones.  Every entry point names the reference interface it stands in for
 * (file:line relative to the hztools/go-sdr tree).
 *
 * Conventions
 *  - Every function returns an hzsdr_status (0 = ok).  A human-readable message for the last
 *    failure on the calling thread is available from hzsdr_last_error() -- the same shape as the
 *    reference's `rvToErr(C.int) error` for librtlsdr (rtl/error.go:31-36).  Status codes 4..7 map
 *    1:1 onto the reference's sentinel errors (iq.go:27-39, conv.go:30).
 *  - There is NO CPU fallback.  Without an sm_100 device hzsdr_ctx_create fails with
 *    HZSDR_ERR_NO_DEVICE (the same spirit as the reference's SIMD CPU-feature gate, which panics:
 *    internal/simd/enabled_amd64.go:35-50).
 *  - Lengths are in IQ samples (one I/Q pair), never bytes, as everywhere in the reference.
 *    Sample layouts are the reference's: interleaved [2]uint8 / [2]int8 / [2]int16 / complex64
 *    (iq_u8.go:35, iq_i8.go:31, iq_i16.go:50, iq_c64.go:38).
 *  - Pointers named *_dev are device pointers valid on the context's GPU; *_host are host
 *    pointers.  Kernels are enqueued on the context's CUDA stream and return immediately;
 *    hzsdr_ctx_sync / hzsdr_download / *_exec_host wait.  A context may be used from any OS
 *    thread (cgo gives no thread affinity); calls on one context must not race each other.
 */
#ifndef HZSDR_CUDA_H
#define HZSDR_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(HZSDR_BUILD) && defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

typedef enum hzsdr_status {
    HZSDR_OK = 0,
    HZSDR_ERR_NO_DEVICE = 1,        /* no CUDA device, or not an sm_100 part */
    HZSDR_ERR_CUDA = 2,             /* a CUDA runtime call failed; see hzsdr_last_error() */
    HZSDR_ERR_INVALID = 3,          /* bad argument */
    HZSDR_ERR_DST_TOO_SMALL = 4,    /* sdr.ErrDstTooSmall             iq.go:37-39 */
    HZSDR_ERR_FORMAT_MISMATCH = 5,  /* sdr.ErrSampleFormatMismatch    iq.go:29-31 */
    HZSDR_ERR_FORMAT_UNKNOWN = 6,   /* sdr.ErrSampleFormatUnknown     iq.go:33-35 */
    HZSDR_ERR_CONVERSION_NOT_IMPLEMENTED = 7, /* sdr.ErrConversionNotImplemented conv.go:30 */
    HZSDR_ERR_NOMEM = 8,
    HZSDR_ERR_NCCL = 9,
    HZSDR_ERR_UNSUPPORTED = 10,     /* e.g. an FFT length this build has no kernel for */
    HZSDR_ERR_RING_UNDERRUN = 11    /* stream.ErrRingBufferUnderrun   stream/ring.go:44 */
} hzsdr_status;

/* sdr.SampleFormat ids, iq.go:113-129 */
typedef enum hzsdr_format {
    HZSDR_FORMAT_C64 = 1,
    HZSDR_FORMAT_U8 = 2,
    HZSDR_FORMAT_I16 = 3,
    HZSDR_FORMAT_I8 = 4
} hzsdr_format;

typedef struct hzsdr_ctx hzsdr_ctx;           /* one GPU + one CUDA stream */
typedef struct hzsdr_fft_plan hzsdr_fft_plan; /* fft.Plan            fft/fft.go:52-59 */
typedef struct hzsdr_chain hzsdr_chain;       /* fused Convert->Shift->Convolution->Decimate */
typedef struct hzsdr_ring hzsdr_ring;         /* pinned-host slot ring + async H2D */
typedef struct hzsdr_comm hzsdr_comm;         /* NCCL communicator (multi-GPU Beamform) */

/* ---- library / device ------------------------------------------------------------------- */
const char *hzsdr_last_error(void);
const char *hzsdr_version(void);
/* bytes per IQ sample: SampleFormat.Size(), iq.go:99-110; 0 for an unknown format */
int hzsdr_format_size(int format);
int hzsdr_device_count(int *count);
/* Replaces the reference's SIMD backend probe (internal/simd/enabled_amd64.go:35-50). */
int hzsdr_ctx_create(int device, hzsdr_ctx **out);
int hzsdr_ctx_destroy(hzsdr_ctx *ctx);
int hzsdr_ctx_sync(hzsdr_ctx *ctx);
/* Completes everything the *_submit_host entry points below (hzsdr_beamform_submit_host,
 * hzsdr_channelizer_submit_host) have enqueued on this context, their copy streams included: after
 * it the destination host buffers hold the results.  (sdr.Reader.Read returning, reader.go:39-47.) */
int hzsdr_ctx_wait_host(hzsdr_ctx *ctx);
/* the context's cudaStream_t, for interop (e.g. timing with CUDA events on this stream) */
int hzsdr_ctx_stream(hzsdr_ctx *ctx, void **cuda_stream);
/* what debug.ReadBuildInfo would report for the `cuda` backend (debug/build.go:60-75) */
int hzsdr_ctx_info(hzsdr_ctx *ctx, char *name, size_t name_len, int *sm_major, int *sm_minor,
                   int *sm_count, size_t *hbm_bytes);

/* ---- memory: the device-resident SamplesC64 / raw buffers of the `cuda` Go package ------- */
int hzsdr_dev_alloc(hzsdr_ctx *ctx, size_t bytes, void **out_dev);
int hzsdr_dev_free(hzsdr_ctx *ctx, void *dev);
int hzsdr_dev_memset(hzsdr_ctx *ctx, void *dev, int value, size_t bytes);
/* cudaHostAlloc'd memory the Go side exposes as ordinary SamplesU8/I8/I16 through yikes.Samples
 * (yikes/bytes.go:50-71) and hands to stream.RingBufferOptions.IQBufferAllocator
 * (stream/ring.go:60-64). */
int hzsdr_pinned_alloc(size_t bytes, void **out_host);
int hzsdr_pinned_free(void *host);
/* async on the context stream when host memory is pinned; the host buffer must stay valid until
 * the next hzsdr_ctx_sync (cgo callers therefore only pass library-owned pinned memory). */
int hzsdr_upload(hzsdr_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes);
/* copies and waits: on return dst_host holds the data (the D2H a host sdr.SamplesC64 needs). */
int hzsdr_download(hzsdr_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes);
int hzsdr_copy(hzsdr_ctx *ctx, void *dst_dev, const void *src_dev, size_t bytes);

/* ---- K1  sdr.ConvertBuffer(dst C64, src), conv.go:55-93 ---------------------------------- *
 * -> SamplesU8.ToC64 iq_u8.go:103-121 (+ asm iq_u8_amd64.s:27-90), SamplesI8.ToC64
 * iq_i8.go:99-119, SamplesI16.ToC64 iq_i16.go:137-147.  Bit-exact.  src_len > dst_len is
 * HZSDR_ERR_DST_TOO_SMALL (conv.go:60-62); src_format == C64 is a copy (conv.go:56-58).
 * *n_out = samples converted (= src_len). */
int hzsdr_convert_to_c64(hzsdr_ctx *ctx, int src_format, const void *src_dev, size_t src_len,
                         void *dst_dev, size_t dst_len, size_t *n_out);
/* The rest of ConvertBuffer's 4x4 matrix (conv.go:36-46), bit-exact:
 * complex64 -> u8 / i16 / i8 (iq_c64.go:77-117: fp32 multiply, separate add, truncation toward zero as
 * the amd64 compiler emits it), u8 <-> i8 <-> i16 (iq_u8.go:73-101, iq_i8.go:73-97,
 * iq_i16.go:116-134,150-162).  dst_format == C64 forwards to hzsdr_convert_to_c64. */
int hzsdr_convert(hzsdr_ctx *ctx, int src_format, const void *src_dev, size_t src_len, int dst_format,
                  void *dst_dev, size_t dst_len, size_t *n_out);
/* SamplesI16.ShiftLSBToMSBBits, iq_i16.go:103-111 (what pluto/rx.go:146 does on the CPU) */
int hzsdr_i16_shift_lsb_to_msb(hzsdr_ctx *ctx, void *buf_dev, size_t n, int bits);
/* K1L  LookupTable.Lookup, iq_lookup_table.go:129-147,198-251: dst[i] = table[index(src[i])],
 * index = the IQ byte pair as a little-endian uint16 (:56-64).  table_dev has 65536 samples of
 * table_format; src_format is U8 or I8. */
int hzsdr_lookup(hzsdr_ctx *ctx, int src_format, const void *src_dev, size_t n,
                 int table_format, const void *table_dev, void *dst_dev, size_t dst_len);

/* ---- K2  stream.ShiftBuffer / ShiftReader, stream/shifter.go:44-102 ---------------------- *
 * The reference's closure state: `ts` (fp64 seconds, serially accumulated, wrapped at 2*pi
 * seconds) and the sample rate (shifter.go:67-71).  ts after a call is bit-equal to the
 * reference's accumulator; the host builds a closed-form segment table per call (csrc/nco.h). */
typedef struct hzsdr_nco {
    uint32_t sample_rate;
    double ts;
} hzsdr_nco;
/* in place on a device complex64 buffer */
int hzsdr_shift(hzsdr_ctx *ctx, void *buf_dev, size_t n, double freq_hz, hzsdr_nco *state);
/* fused K1+K2: raw -> complex64 -> mixed, one pass over HBM (ConvertReader + ShiftReader) */
int hzsdr_convert_shift(hzsdr_ctx *ctx, int src_format, const void *src_dev, size_t n,
                        void *dst_dev, size_t dst_len, double freq_hz, hzsdr_nco *state);
/* `count` consecutive buffers of the stream (n_each samples each, e.g. drained ring slots) through the fused
 * ConvertReader + ShiftReader in ONE kernel launch per <= 64 buffers; same results and carried state as `count`
 * calls of hzsdr_convert_shift.  srcs / dsts: host arrays of device pointers. */
int hzsdr_convert_shift_batch(hzsdr_ctx *ctx, int src_format, const void *const *srcs_host, size_t n_each,
                              void *const *dsts_host, size_t dst_len_each, size_t count, double freq_hz, hzsdr_nco *state);

/* ---- K3/K4/K5  Multiply / Gain / Add ------------------------------------------------------ *
 * hzsdr_rotate: simd.RotateComplex internal/simd/mult.go:29-47 via SamplesC64.Multiply
 *   iq_c64.go:128-130; products widened to fp64 exactly as the Go compiler does, so results are
 *   bit-equal to the amd64 build.  (m == 1 is the caller's no-op, stream/multiply.go:59-62.)
 * hzsdr_scale: simd.ScaleComplex internal/simd/mult_simd_amd64.s:27-55 (stream/gain.go:39-57).
 * hzsdr_add: addReader.Read stream/add.go:121-185: dst = ((0 + s0) + s1) + ... in reader order,
 *   fp32; srcs_host is a host array of k device pointers; dst may alias none of them. */
int hzsdr_rotate(hzsdr_ctx *ctx, void *buf_dev, size_t n, float m_re, float m_im);
int hzsdr_scale(hzsdr_ctx *ctx, void *buf_dev, size_t n, float r);
int hzsdr_add(hzsdr_ctx *ctx, void *dst_dev, const void *const *srcs_host, int k, size_t n);
/* stream.Add on I8 / I16 readers: wrapping component-wise integer adds (stream/add.go:95-113) */
int hzsdr_add_int(hzsdr_ctx *ctx, int format, void *dst_dev, const void *const *srcs_host, int k, size_t n);

/* ---- K7  Decimate / Downsample ------------------------------------------------------------ *
 * hzsdr_decimate: stream.DecimateBuffer stream/decimate.go:59-101 (formats U8, I16, C64 -- I8 is
 *   HZSDR_ERR_FORMAT_UNKNOWN as in the reference :85-97).  block == 0: one buffer,
 *   to[i] = from[factor*i], i < n/factor.  block > 0: DecimateReader semantics
 *   (stream/decimate.go:34-51, block = 32768): floor(n/block) whole blocks, phase restarts per
 *   block, floor(block/factor) outputs each, trailing partial block dropped
 *   (stream/read_transformer.go:121-125).  dst_len too small: HZSDR_ERR_DST_TOO_SMALL, *n_out = 0.
 * hzsdr_downsample: stream.DownsampleBuffer stream/downsample.go:68-127 (src U8, I16 or C64; dst
 *   always C64): sequential fp32 sum of `factor` samples divided by float32(factor). */
int hzsdr_decimate(hzsdr_ctx *ctx, int format, const void *src_dev, size_t n, void *dst_dev,
                   size_t dst_len, unsigned factor, size_t block, size_t *n_out);
int hzsdr_downsample(hzsdr_ctx *ctx, int src_format, const void *src_dev, size_t n, void *dst_dev,
                     size_t dst_len, unsigned factor, size_t block, size_t *n_out);

/* ---- K6  fft.Planner / fft.Plan, fft/fft.go:45-59; fft.ConvolveFreq fft/convolution.go:150 - *
 * direction: HZSDR_FFT_FORWARD = fft.Forward (e^{-2 pi i kn/N}), HZSDR_FFT_BACKWARD =
 * fft.Backward (e^{+...}); both unnormalised (the reference has no in-tree FFT; this is the
 * convention the project pins -- see DESIGN.md).  n must be a power of two, 2 <= n <= 2^20
 * (else HZSDR_ERR_UNSUPPORTED); up to 16384 points a transform is one kernel, beyond (the 65536
 * points of rtl/kerberos/internal/align.go:92-100, graft.go:73-80) two through context scratch.  iq_len != freq_len is HZSDR_ERR_DST_TOO_SMALL, the planner
 * contract of testutils/fft.go:127-138. */
#define HZSDR_FFT_FORWARD 1
#define HZSDR_FFT_BACKWARD 0
int hzsdr_fft_plan_create(hzsdr_ctx *ctx, size_t iq_len, size_t freq_len, int direction,
                          hzsdr_fft_plan **out);
/* Plan.Transform over `batch` consecutive length-n vectors; src may equal dst */
int hzsdr_fft_exec(hzsdr_fft_plan *plan, const void *src_dev, void *dst_dev, size_t batch);
int hzsdr_fft_plan_destroy(hzsdr_fft_plan *plan);
/* stream.ConvolutionReader's Proc (stream/convolution.go:62-80) over n_blocks consecutive
 * n_fft-sample blocks: dst = IFFT(FFT(src) * filter), block-circular, one kernel, the spectrum
 * never leaves the SM.  filter_dev: n_fft complex64, frequency domain.  src may equal dst. */
int hzsdr_convolve_freq(hzsdr_ctx *ctx, const void *src_dev, void *dst_dev, const void *filter_dev,
                        size_t n_fft, size_t n_blocks);

/* fft.Convolve / fft.CrossCorrelate (fft/convolution.go:97-139) over `batch`
 * length-n vectors: dst = IFFT(FFT(iq1) * FFT(iq2)), or with conj(FFT(iq2)) when cross_correlate != 0.
 * Unnormalised transforms, so the result carries the factor n.  scratch_dev: n*batch complex64.  dst
 * may alias iq1. */
int hzsdr_fft_convolve(hzsdr_ctx *ctx, void *dst_dev, const void *iq1_dev, const void *iq2_dev, size_t n,
                       size_t batch, int cross_correlate, void *scratch_dev);

/* ---- coherent-receiver helpers: rtl/kerberos/internal --------- */
/* FFTShiftAndScale (rtl/kerberos/internal/reader.go:57-64) over `batch` length-n vectors, in place:
 * the halves of each vector are exchanged and every component divided by `scale` (IEEE fp32). */
int hzsdr_fftshift_scale(hzsdr_ctx *ctx, void *data_dev, size_t n, size_t batch, float scale);
/* One pass of GraftReaders' loop (graft.go:96-125): iq_dev holds n_readers buffers of fft_size
 * complex64, reader-major; each is transformed forward into its slice of freq_dev, shifted and
 * scaled by fft_size, and ONE backward transform of n_readers*fft_size points gives dst_dev (the
 * reader at n_readers x the sample rate).  n_readers*fft_size: a power of two <= 2^20.  freq_dev:
 * n_readers*fft_size complex64 of scratch, distinct from iq_dev and dst_dev. */
int hzsdr_graft(hzsdr_ctx *ctx, const void *iq_dev, size_t n_readers, size_t fft_size, void *dst_dev,
                void *freq_dev);
/* checkAlignment's peak search (align.go:125-146) over `batch` correlation vectors of length n on the
 * device: the first index of maximum power (zeros skipped), minus n when it is beyond n/2; -1 for an
 * all-zero vector.  Results land in offsets_host (the call synchronises). */
int hzsdr_correlate_peak(hzsdr_ctx *ctx, const void *cc_dev, size_t n, size_t batch, int32_t *offsets_host);
/* PhaseOffsets (align.go:244-272): bufs_dev = n_chan buffers of n complex64, channel-major; out_host =
 * n_chan complex64 unit phasors of the mean phase of conj-multiplying channel 0 with channel j (fp64
 * accumulation).  Element 0 reproduces the reference's Rect(1, 1/n).  The call synchronises. */
int hzsdr_phase_offsets(hzsdr_ctx *ctx, const void *bufs_dev, size_t n_chan, size_t n, float *out_host);

/* ---- K8  stream.ReadBeamform data path, stream/beamform.go:148-171 ------------------------ *
 * dst[n] = sum_c w_c * toC64(x_c[n]), accumulated in channel order in fp32 from 0
 * (multiply.go:46-70 + add.go:115-185).  chans_host: host array of nchan device pointers to raw
 * `src_format` buffers of n samples; weights_host: nchan complex64 (BeamformConfig.Angles /
 * SetPhaseAngles, beamform.go:131-145). */
int hzsdr_beamform(hzsdr_ctx *ctx, int src_format, const void *const *chans_host, int nchan,
                   const float *weights_host, size_t n, void *dst_dev);
/* The same from HOST buffers, end to end (ReadBeamform over host readers, beamform.go:148-171): the
 * channels' raw samples travel H2D piece by piece through a 3-slot staging pipe, overlapped with
 * the kernel and with the D2H copy of the finished beam into dst_host (n complex64).  Only enqueues;
 * hzsdr_ctx_wait_host completes it.  Buffers should be pinned (hzsdr_pinned_alloc / ring slots) and
 * must stay valid until then.  Channels at one pitch inside one block travel as a single 2-D copy. */
int hzsdr_beamform_submit_host(hzsdr_ctx *ctx, int src_format, const void *const *chans_host_mem, int nchan,
                               const float *weights_host, size_t n, void *dst_host_mem);
/* stream.BeamformAngles2D / BeamformAngles (beamform.go:57-128): fp64 host math, no GPU needed.
 * antennas_xy: 2*n doubles; out_weights: 2*n floats (complex64). */
int hzsdr_beamform_angles_2d(double frequency_hz, double angle_deg, const double center_xy[2],
                             const double *antennas_xy, int n, float *out_weights);

/* ---- fused chain: ConvertReader -> ShiftReader -> ConvolutionReader -> DecimateReader ------ *
 * (stream/convert.go:37, shifter.go:89, convolution.go:36, decimate.go:34) as ONE kernel per
 * buffer: raw samples are read once, the decimated complex64 written once. */
typedef struct hzsdr_chain_config {
    int src_format;            /* U8, I8 or I16 */
    uint32_t sample_rate;      /* Reader.SampleRate() */
    double shift_hz;           /* ShiftReader's rf.Hz */
    size_t n_fft;              /* len(filter): ConvolutionReader block length */
    const void *filter_host;   /* n_fft complex64, frequency domain (host memory, copied) */
    uint32_t decimate;         /* DecimateReader factor (>= 1) */
    uint32_t decimate_block;   /* 0 -> 32768, the reference's fixed block (decimate.go:41) */
    int i16_lsb_bits;          /* 0, or the ADC width for ShiftLSBToMSBBits (pluto: 12) */
    /* 0: block-circular ConvolutionReader, the reference's semantics (stream/convolution.go:57-81).
     * T > 0 (EXTENSION, BASELINE config 3's "overlap-save FIR"): `filter_host` is the n_fft-bin spectrum of a
     * T-tap filter (FFT of the zero-padded taps, / n_fft) and the chain computes the TRUE linear convolution
     * z[n] = sum_k h[k] y[n-k] of the mixed stream (y[n < 0] = 0 at stream start, history carried between
     * calls) by overlap-save inside the fused kernel, then decimates.  n_fft = 16384, decimate % 16 == 0,
     * T <= 8193.  hzsdr_chain_set_ts restarts the stream (the carried history is dropped). */
    uint32_t overlap_save_taps;
} hzsdr_chain_config;
int hzsdr_chain_create(hzsdr_ctx *ctx, const hzsdr_chain_config *cfg, hzsdr_chain **out);
int hzsdr_chain_destroy(hzsdr_chain *chain);
/* samples the chain emits for n input samples that start on a block boundary */
int hzsdr_chain_out_len(const hzsdr_chain *chain, size_t n, size_t *n_out);
/* device-resident: n must be a multiple of lcm(n_fft, decimate_block) */
int hzsdr_chain_exec(hzsdr_chain *chain, const void *src_dev, size_t n, void *dst_dev,
                     size_t dst_len, size_t *n_out);
/* `count` consecutive buffers of the stream (n_each samples each, e.g. `count` drained ring slots) in
 * one call: srcs_host / dsts_host are host arrays of device pointers; every buffer emits *n_out_each.
 * Same results as `count` calls of hzsdr_chain_exec.  n_fft = 1024 chains with an even decimation
 * factor run the whole batch as ONE kernel launch once it is large enough to fill the chip several
 * times over (tables staged once per batch, in Tensor Memory); other shapes launch per buffer. */
int hzsdr_chain_exec_batch(hzsdr_chain *chain, const void *const *srcs_host, size_t n_each,
                           void *const *dsts_host, size_t dst_len_each, size_t count, size_t *n_out_each);
/* end to end: H2D of the raw buffer, the fused kernel, D2H of the result, then wait */
int hzsdr_chain_exec_host(hzsdr_chain *chain, const void *src_host, size_t n, void *dst_host,
                          size_t dst_len, size_t *n_out);
/* pipelined end to end (what a GPU-backed Reader does between Read calls): enqueue H2D of a
 * pinned raw buffer, the fused kernel and the D2H into a pinned result buffer on three streams
 * with 3-deep device staging, and return without waiting; buffer k+1's H2D overlaps buffer k's
 * kernel and buffer k-1's D2H.  *n_out is known at submit time.  Both host buffers must be pinned
 * (hzsdr_pinned_alloc / ring slots) and stay untouched until hzsdr_chain_wait_host returns. */
int hzsdr_chain_submit_host(hzsdr_chain *chain, const void *src_host, size_t n, void *dst_host,
                            size_t dst_len, size_t *n_out);
int hzsdr_chain_wait_host(hzsdr_chain *chain);
/* the carried NCO state (checkpoint / resume of a stream; shifter.go:68) */
int hzsdr_chain_get_ts(const hzsdr_chain *chain, double *ts);
int hzsdr_chain_set_ts(hzsdr_chain *chain, double ts);

/* ---- FIR + decimate by true linear convolution (EXTENSION: no reference counterpart) ----------- *
 * The reference's ConvolutionReader is block-circular (stream/convolution.go:57-81); BASELINE's
 * "overlap-save FIR" and "fused polyphase FIR+decimate" are defined here as
 *     z[n] = sum_k taps[k] * y[n-k]  (y[n<0] = 0, history carried between calls),  out = z[D*i]
 * over a device-resident complex64 stream.  taps: ntaps complex64 (interleaved re, im), host memory.
 * method: overlap-save runs windows through the fused FFT -> xH -> IFFT kernel; polyphase computes
 * only the kept outputs in direct form; AUTO picks by taps/decimate. */
#define HZSDR_FIR_AUTO 0
#define HZSDR_FIR_OVERLAP_SAVE 1
#define HZSDR_FIR_POLYPHASE 2
typedef struct hzsdr_fir hzsdr_fir;
int hzsdr_fir_create(hzsdr_ctx *ctx, const float *taps, size_t ntaps, unsigned decimate, int method, hzsdr_fir **out);
int hzsdr_fir_destroy(hzsdr_fir *fir);
int hzsdr_fir_reset(hzsdr_fir *fir);
int hzsdr_fir_exec(hzsdr_fir *fir, const void *src_dev, size_t n, void *dst_dev, size_t dst_len, size_t *n_out);

/* ---- fused polyphase decimator on a RAW stream (EXTENSION: BASELINE north_star's "fused polyphase
 * FIR+decimate"; no reference counterpart) ----------------------------------------------------- *
 * ConvertReader (iq_u8.go:111-121 / iq_i8.go:107-119 / iq_i16.go:141-145) -> ShiftReader
 * (stream/shifter.go:66-85, the fp64 time accumulator carried between calls) -> FIR with REAL taps ->
 * keep every `decimate`-th sample, as ONE kernel that computes only the kept outputs:
 *     z[n] = sum_k taps[k] * y[n-k]  (y[n<0] = 0, raw history carried between calls),  out[i] = z[D*i]
 * with a continuous decimation phase over the stream (not DecimateReader's 32768-sample restarts).
 * Cheaper than the FFT forms when ntaps / decimate is below ~16; long filters: hzsdr_chain_* with
 * overlap_save_taps.  taps: ntaps floats, host memory.  Any n per call. */
typedef struct hzsdr_polyphase hzsdr_polyphase;
int hzsdr_polyphase_create(hzsdr_ctx *ctx, int src_format, uint32_t sample_rate, double shift_hz, const float *taps,
                           size_t ntaps, unsigned decimate, int i16_lsb_bits, hzsdr_polyphase **out);
int hzsdr_polyphase_destroy(hzsdr_polyphase *p);
/* outputs the NEXT n samples of the stream will produce */
int hzsdr_polyphase_out_len(const hzsdr_polyphase *p, size_t n, size_t *n_out);
int hzsdr_polyphase_exec(hzsdr_polyphase *p, const void *src_dev, size_t n, void *dst_dev, size_t dst_len, size_t *n_out);
int hzsdr_polyphase_get_ts(const hzsdr_polyphase *p, double *ts);
int hzsdr_polyphase_set_ts(hzsdr_polyphase *p, double ts);

/* ---- channelizer: n_streams independent chains, ONE kernel launch per set of buffers ------- *
 * BASELINE config 5.  Every stream is what the reference builds as its own reader chain
 * (stream/convert.go:37 -> shifter.go:89 -> convolution.go:36 -> decimate.go:34); the streams share
 * cfg's format, sample rate, filter and decimation and differ in mixer frequency (shift_hz[s];
 * cfg->shift_hz is ignored) and carried NCO time.  srcs_host / dsts_host are host arrays of
 * n_streams device pointers; every stream consumes n samples and produces *n_out_each. */
typedef struct hzsdr_channelizer hzsdr_channelizer;
int hzsdr_channelizer_create(hzsdr_ctx *ctx, const hzsdr_chain_config *cfg, const double *shift_hz,
                             size_t n_streams, hzsdr_channelizer **out);
int hzsdr_channelizer_destroy(hzsdr_channelizer *chz);
int hzsdr_channelizer_exec(hzsdr_channelizer *chz, const void *const *srcs_host, size_t n,
                           void *const *dsts_host, size_t dst_len, size_t *n_out_each);
/* End to end: srcs_host_mem / dsts_host_mem are host arrays of n_streams HOST buffers (pinned).  The
 * streams cross PCIe in groups through the context's staging pipe, overlapped with the kernel and
 * the return copies.  Only enqueues; hzsdr_ctx_wait_host completes it. */
int hzsdr_channelizer_submit_host(hzsdr_channelizer *chz, const void *const *srcs_host_mem, size_t n,
                                  void *const *dsts_host_mem, size_t dst_len, size_t *n_out_each);
int hzsdr_channelizer_get_ts(const hzsdr_channelizer *chz, double *ts_out /* n_streams */);
int hzsdr_channelizer_set_ts(hzsdr_channelizer *chz, const double *ts /* n_streams */);

/* ---- pinned-host ring: stream.RingBuffer with a cudaHostAlloc allocator, stream/ring.go ---- *
 * Producers (SDR driver callbacks) write raw samples into the next pinned slot
 * (UnsafeRingBuffer.WritePeekUnsafePointer / WritePoke, ring.go:359-379); poke starts the async
 * H2D of that slot on a copy stream; the consumer gets the device copy in order.  Overrun
 * overwrites the oldest unread slot like the reference (ring.go:170-186, :266). */
int hzsdr_ring_create(hzsdr_ctx *ctx, int format, size_t slots, size_t slot_len, hzsdr_ring **out);
int hzsdr_ring_destroy(hzsdr_ring *ring);
int hzsdr_ring_write_peek(hzsdr_ring *ring, void **slot_host);
int hzsdr_ring_write_poke(hzsdr_ring *ring, size_t n_samples);
/* next unread slot's device copy, made visible to the context stream; HZSDR_ERR_RING_UNDERRUN if
 * nothing is pending (the BlockReads=false behaviour, ring.go:216-219) */
int hzsdr_ring_read(hzsdr_ring *ring, const void **slot_dev, size_t *n_samples);
int hzsdr_ring_read_done(hzsdr_ring *ring);
/* One ring and one thread on each side: write_peek / write_poke may be called from a producer thread
 * (an SDR driver callback) while the context's owner thread reads and computes.
 *
 * Driver hand-off end to end (stream/ring.go:344-392 feeding the ReadTransformer chain): the next unread
 * slot goes through the fused chain and the result is copied to dst_host_mem (pinned) behind the kernel.
 * HZSDR_ERR_RING_UNDERRUN when no slot is pending.  Only enqueues; hzsdr_chain_wait_host completes it. */
int hzsdr_chain_submit_ring(hzsdr_chain *chain, hzsdr_ring *ring, void *dst_host_mem, size_t dst_len, size_t *n_out);

/* ---- multi-GPU Beamform: NCCL sum of per-GPU partial beams over NVLink --------------------- */
#define HZSDR_NCCL_UNIQUE_ID_BYTES 128
int hzsdr_comm_unique_id(void *id_out /* HZSDR_NCCL_UNIQUE_ID_BYTES */);
int hzsdr_comm_create(hzsdr_ctx *ctx, int nranks, int rank, const void *id, hzsdr_comm **out);
int hzsdr_comm_destroy(hzsdr_comm *comm);
/* in-place fp32 sum of n complex64 samples onto `root` (ncclReduce), on the context stream */
int hzsdr_comm_reduce_c64(hzsdr_comm *comm, void *buf_dev, size_t n, int root);
int hzsdr_comm_allreduce_c64(hzsdr_comm *comm, void *buf_dev, size_t n);

/* ---- multi-GPU Beamform, reduction fused into the kernel over NVLink peer memory ----------- *
 * One process per GPU.  Channels are sharded across ranks; every rank computes the partial beam of
 * its channels and writes the part that belongs to output slice s (n/nranks samples) directly into
 * rank s's memory (CUDA IPC peer stores) while it computes -- a reduce-scatter overlapped with the
 * math.  dst_slice receives this rank's n/nranks finished samples, summed in rank order; n must be
 * a multiple of 128 * nranks.
 * create: allocates the staging area (sized for exchanges of up to max_batch buffers, 1..64) and returns
 * its IPC handle; exchange the handles of all ranks by any means (torch.distributed, MPI, a pipe),
 * rank-major, and pass them to connect.
 * exec_batch: ONE exchange (one kernel, one flag round, one finishing kernel) over nbuf buffers of n
 * samples: chans_host[k * nchan + c] = channel c of buffer k (device pointers), dst_slices_host[k]
 * receives this rank's n/nranks samples of buffer k.  Batching is what makes the exchange NVLink-bound
 * instead of latency-bound: at 8 GPUs one 2^20-sample buffer is ~10 us of NVLink against ~35 us of
 * launch + flag latency.  exec = exec_batch with nbuf = 1.  Every rank must make the same sequence of
 * calls.  nbuf * nchan <= 1024. */
#define HZSDR_IPC_HANDLE_BYTES 64
typedef struct hzsdr_beam_group hzsdr_beam_group;
int hzsdr_beam_group_create(hzsdr_ctx *ctx, int nranks, int rank, size_t n, size_t max_batch, void *handle_out,
                            hzsdr_beam_group **out);
int hzsdr_beam_group_connect(hzsdr_beam_group *group, const void *all_handles /* nranks * 64 bytes */);
int hzsdr_beam_group_exec(hzsdr_beam_group *group, int src_format, const void *const *chans_host, int nchan,
                          const float *weights_host, void *dst_slice_dev);
int hzsdr_beam_group_exec_batch(hzsdr_beam_group *group, int src_format, const void *const *chans_host, int nchan,
                                const float *weights_host, size_t nbuf, void *const *dst_slices_host);
/* exec only enqueues: the finishing sum runs on a side stream so the next buffer's compute does not
 * queue behind the wait for the peers.  join makes the context stream (and thus hzsdr_ctx_sync and
 * later kernels) wait for every slice produced so far. */
int hzsdr_beam_group_join(hzsdr_beam_group *group);
int hzsdr_beam_group_destroy(hzsdr_beam_group *group);

#if defined(HZSDR_BUILD) && defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* HZSDR_CUDA_H */
