// hzsdr.hpp -- C++ mirror of hz.tools/sdr's reader API for the IQ sample chain, GPU-backed.
//
// The product's host language is Go (go-sdr_b200/go: a cgo `cuda/` package plus `sdr.cuda`-tagged
// twins of conv.go and stream/*.go).  No Go toolchain exists in the build image, so this header is
// the compiled, tested host side above the C ABI: same names, argument meaning, block rules and
// error behaviour as the reference, so that go-sdr_b200/host/test_host.cpp reads like the
// reference's own tests.  Everything here calls libhzsdrcuda.so through include/hzsdr_cuda.h;
// there is no CPU implementation of any sample arithmetic in this file.
//
//   sdr::Samples / SamplesU8 / I8 / I16 / C64   iq.go:59-87, iq_u8.go:35, iq_i8.go:31, iq_i16.go:50, iq_c64.go:38
//   cuda::SamplesC64 (device-resident)          the 5th Samples type of the `sdr.cuda` build
//   sdr::Reader, ReadFull, ReadAtLeast          reader.go:39-113
//   sdr::ConvertBuffer, CopySamples             conv.go:55-93, copy.go:31-52
//   stream::ConvertReader / ShiftReader / ConvolutionReader / DecimateReader / DownsampleReader /
//           Multiply / Gain / Add / ReadBeamform   stream/*.go
//   fft::Planner / Plan / ConvolveFreq          fft/fft.go:45-59, fft/convolution.go:150-192
//
// Error convention: Go's `(int, error)` becomes sdr::Result{n, err}; the reference's sentinel
// errors are global objects compared by identity (`r.err == sdr::ErrDstTooSmall`), exactly as Go
// compares them with `==`.  Errors latch on a stage as they do on the reference's pipes
// (stream/read_transformer.go:121-135).
#pragma once
#include <algorithm>
#include <array>
#include <complex>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/hzsdr_cuda.h"

namespace sdr {

// ---- errors ----------------------------------------------------------------------------------
struct Error {
    std::string msg;
};
using Err = std::shared_ptr<const Error>;
inline Err make_err(const char *m) { return std::make_shared<const Error>(Error{m}); }
inline const Err ErrSampleFormatMismatch = make_err("sdr: iq sample formats do not match");        // iq.go:29-31
inline const Err ErrSampleFormatUnknown = make_err("sdr: iq sample format is not understood");     // iq.go:33-35
inline const Err ErrDstTooSmall = make_err("sdr: destination sample buffer is too small");         // iq.go:37-39
inline const Err ErrConversionNotImplemented = make_err("sdr: unknown format conversion");         // conv.go:30
inline const Err ErrShortBuffer = make_err("sdr: short read");                                     // reader.go:31
inline const Err ErrUnexpectedEOF = make_err("sdr: expected EOF");                                 // reader.go:35
inline const Err ErrPipeClosed = make_err("sdr: pipe closed");                                     // pipe.go:30
inline const Err EOF_ = make_err("EOF");                                                           // io.EOF

struct Result {
    int n = 0;
    Err err;
};

// library status -> the reference's sentinel, or a fresh error carrying hzsdr_last_error()
inline Err from_status(int rc) {
    switch (rc) {
        case HZSDR_OK: return nullptr;
        case HZSDR_ERR_DST_TOO_SMALL: return ErrDstTooSmall;
        case HZSDR_ERR_FORMAT_MISMATCH: return ErrSampleFormatMismatch;
        case HZSDR_ERR_FORMAT_UNKNOWN: return ErrSampleFormatUnknown;
        case HZSDR_ERR_CONVERSION_NOT_IMPLEMENTED: return ErrConversionNotImplemented;
        default: return std::make_shared<const Error>(Error{std::string("hzsdrcuda: ") + hzsdr_last_error()});
    }
}

// ---- sample formats and buffers ---------------------------------------------------------------
enum class SampleFormat : uint8_t { C64 = 1, U8 = 2, I16 = 3, I8 = 4 };  // iq.go:113-129
inline int FormatSize(SampleFormat f) { return hzsdr_format_size((int)f); }  // iq.go:99-110

class Samples {  // iq.go:59-87
   public:
    virtual ~Samples() = default;
    virtual SampleFormat Format() const = 0;
    virtual int Length() const = 0;
    int Size() const { return FormatSize(Format()) * Length(); }
    virtual std::shared_ptr<Samples> Slice(int start, int end) const = 0;  // aliases, like a Go slice
    virtual void *Data() const = 0;
    virtual bool OnDevice() const { return false; }
};
using SamplesPtr = std::shared_ptr<Samples>;

template <typename T, SampleFormat F>
class HostSamples : public Samples {
   public:
    explicit HostSamples(int n) : store_(std::make_shared<std::vector<T>>(n)), p_(store_->data()), n_(n) {}
    HostSamples(std::shared_ptr<std::vector<T>> store, T *p, int n) : store_(std::move(store)), p_(p), n_(n) {}
    SampleFormat Format() const override { return F; }
    int Length() const override { return n_; }
    std::shared_ptr<Samples> Slice(int a, int b) const override { return std::make_shared<HostSamples>(store_, p_ + a, b - a); }
    void *Data() const override { return p_; }
    T &operator[](int i) { return p_[i]; }
    const T &operator[](int i) const { return p_[i]; }
    T *begin() { return p_; }
    T *end() { return p_ + n_; }

   private:
    std::shared_ptr<std::vector<T>> store_;
    T *p_;
    int n_;
};
using SamplesU8 = HostSamples<std::array<uint8_t, 2>, SampleFormat::U8>;
using SamplesI8 = HostSamples<std::array<int8_t, 2>, SampleFormat::I8>;
using SamplesI16 = HostSamples<std::array<int16_t, 2>, SampleFormat::I16>;
using SamplesC64 = HostSamples<std::complex<float>, SampleFormat::C64>;

inline SamplesPtr MakeSamples(SampleFormat f, int n) {  // iq.go:135-149
    switch (f) {
        case SampleFormat::U8: return std::make_shared<SamplesU8>(n);
        case SampleFormat::I8: return std::make_shared<SamplesI8>(n);
        case SampleFormat::I16: return std::make_shared<SamplesI16>(n);
        case SampleFormat::C64: return std::make_shared<SamplesC64>(n);
    }
    return nullptr;
}

}  // namespace sdr

// ---- the `cuda` package: context + device-resident samples ------------------------------------
namespace cuda {

class Context {
   public:
    // Fails loudly (throws) when no sm_100 GPU is present: there is no CPU fallback.
    explicit Context(int device = 0) {
        int rc = hzsdr_ctx_create(device, &h_);
        if (rc != HZSDR_OK) throw std::runtime_error(std::string("cuda.NewContext: ") + hzsdr_last_error());
    }
    ~Context() { hzsdr_ctx_destroy(h_); }
    Context(const Context &) = delete;
    hzsdr_ctx *h() const { return h_; }
    void Sync() const { hzsdr_ctx_sync(h_); }

   private:
    hzsdr_ctx *h_ = nullptr;
};
using ContextPtr = std::shared_ptr<Context>;

struct DeviceBlock {  // owning device allocation
    ContextPtr ctx;
    void *p = nullptr;
    size_t bytes = 0;
    DeviceBlock(ContextPtr c, size_t b) : ctx(std::move(c)), bytes(b) {
        if (hzsdr_dev_alloc(ctx->h(), b, &p) != HZSDR_OK) throw std::runtime_error(hzsdr_last_error());
    }
    ~DeviceBlock() { hzsdr_dev_free(ctx->h(), p); }
};

// Raw or complex64 samples living in HBM.  A GPU reader handed one of these leaves its output on
// the device; handed a host Samples it copies the result back (SURVEY.md 8(b) ownership rule).
class DeviceSamples : public sdr::Samples {
   public:
    DeviceSamples(ContextPtr ctx, sdr::SampleFormat f, int n)
        : blk_(std::make_shared<DeviceBlock>(ctx, (size_t)std::max(n, 1) * sdr::FormatSize(f))), f_(f), off_(0), n_(n) {}
    DeviceSamples(std::shared_ptr<DeviceBlock> b, sdr::SampleFormat f, size_t off, int n) : blk_(std::move(b)), f_(f), off_(off), n_(n) {}
    sdr::SampleFormat Format() const override { return f_; }
    int Length() const override { return n_; }
    std::shared_ptr<sdr::Samples> Slice(int a, int b) const override {
        return std::make_shared<DeviceSamples>(blk_, f_, off_ + (size_t)a * sdr::FormatSize(f_), b - a);
    }
    void *Data() const override { return (uint8_t *)blk_->p + off_; }
    bool OnDevice() const override { return true; }
    const ContextPtr &ctx() const { return blk_->ctx; }

   private:
    std::shared_ptr<DeviceBlock> blk_;
    sdr::SampleFormat f_;
    size_t off_;
    int n_;
};
inline std::shared_ptr<DeviceSamples> NewSamplesC64(ContextPtr ctx, int n) {
    return std::make_shared<DeviceSamples>(std::move(ctx), sdr::SampleFormat::C64, n);
}

}  // namespace cuda

namespace sdr {

// CopySamples, copy.go:31-52 (+ the device cases the `sdr.cuda` build adds).  Copies min(len).
inline Result CopySamples(cuda::Context &ctx, Samples &dst, const Samples &src) {
    if (dst.Format() != src.Format()) return {0, ErrSampleFormatMismatch};
    const int n = std::min(dst.Length(), src.Length());
    const size_t bytes = (size_t)n * FormatSize(dst.Format());
    int rc = HZSDR_OK;
    if (!dst.OnDevice() && !src.OnDevice())
        std::memcpy(dst.Data(), src.Data(), bytes);
    else if (dst.OnDevice() && src.OnDevice())
        rc = hzsdr_copy(ctx.h(), dst.Data(), src.Data(), bytes);
    else if (dst.OnDevice()) {
        rc = hzsdr_upload(ctx.h(), dst.Data(), src.Data(), bytes);
        if (rc == HZSDR_OK) rc = hzsdr_ctx_sync(ctx.h());
    } else
        rc = hzsdr_download(ctx.h(), dst.Data(), src.Data(), bytes);
    if (rc != HZSDR_OK) return {0, from_status(rc)};
    return {n, nullptr};
}

// ---- Reader -----------------------------------------------------------------------------------
class Reader {  // reader.go:39-51
   public:
    virtual ~Reader() = default;
    virtual Result Read(Samples &s) = 0;
    virtual SampleFormat Format() const = 0;  // Reader.SampleFormat()
    virtual unsigned SampleRate() const = 0;
};
using ReaderPtr = std::shared_ptr<Reader>;

inline Result ReadAtLeast(Reader &r, Samples &buf, int min) {  // reader.go:94-113
    if (buf.Length() < min) return {0, ErrShortBuffer};
    int n = 0;
    Err err;
    while (n < min && !err) {
        auto sl = buf.Slice(n, buf.Length());
        Result rr = r.Read(*sl);
        n += rr.n;
        err = rr.err;
    }
    if (n >= min) return {n, err};
    if (n > 0 && err == EOF_) return {n, ErrUnexpectedEOF};
    return {n, err};
}
inline Result ReadFull(Reader &r, Samples &buf) { return ReadAtLeast(r, buf, buf.Length()); }  // reader.go:72

// A Reader over a fixed host buffer, handing out at most `chunk` samples per Read and io.EOF at
// the end: the role sdr.Pipe + a writer goroutine play in the reference's tests.
class BufferReader : public Reader {
   public:
    BufferReader(SamplesPtr data, unsigned rate, int chunk = 1 << 30) : data_(std::move(data)), rate_(rate), chunk_(chunk) {}
    Result Read(Samples &s) override {
        if (s.Format() != data_->Format()) return {0, ErrSampleFormatMismatch};  // pipe.go:82-84
        if (pos_ >= data_->Length()) return {0, EOF_};
        const int n = std::min({s.Length(), data_->Length() - pos_, chunk_});
        std::memcpy(s.Data(), (uint8_t *)data_->Data() + (size_t)pos_ * FormatSize(s.Format()), (size_t)n * FormatSize(s.Format()));
        pos_ += n;
        return {n, nullptr};
    }
    SampleFormat Format() const override { return data_->Format(); }
    unsigned SampleRate() const override { return rate_; }

   private:
    SamplesPtr data_;
    unsigned rate_;
    int chunk_, pos_ = 0;
};

// sdr.ConvertBuffer, conv.go:55-93: every pair of formats (host or device buffers; host buffers are
// staged through the device).  Same-format is CopySamples.
inline Result ConvertBuffer(cuda::Context &ctx, Samples &dst, const Samples &src) {
    if (src.Format() == dst.Format()) return CopySamples(ctx, dst, src);
    if (src.Length() > dst.Length()) return {0, ErrDstTooSmall};
    const int n = src.Length();
    if (n == 0) return {0, nullptr};
    void *d_src = nullptr, *d_dst = nullptr;
    int rc = HZSDR_OK;
    const size_t sb = (size_t)n * FormatSize(src.Format()), db = (size_t)n * FormatSize(dst.Format());
    if (src.OnDevice())
        d_src = src.Data();
    else {
        rc = hzsdr_dev_alloc(ctx.h(), sb, &d_src);
        if (rc == HZSDR_OK) rc = hzsdr_upload(ctx.h(), d_src, src.Data(), sb);
    }
    if (rc == HZSDR_OK) {
        if (dst.OnDevice())
            d_dst = dst.Data();
        else
            rc = hzsdr_dev_alloc(ctx.h(), db, &d_dst);
    }
    size_t got = 0;
    if (rc == HZSDR_OK) rc = hzsdr_convert(ctx.h(), (int)src.Format(), d_src, n, (int)dst.Format(), d_dst, n, &got);
    if (rc == HZSDR_OK && !dst.OnDevice()) rc = hzsdr_download(ctx.h(), dst.Data(), d_dst, got * FormatSize(dst.Format()));
    if (rc == HZSDR_OK) rc = hzsdr_ctx_sync(ctx.h());
    if (!src.OnDevice() && d_src) hzsdr_dev_free(ctx.h(), d_src);
    if (!dst.OnDevice() && d_dst) hzsdr_dev_free(ctx.h(), d_dst);
    if (rc != HZSDR_OK) return {0, from_status(rc)};
    return {(int)got, nullptr};
}

// ---- Writer -----------------------------------------------------------------------------------
class Writer {  // writer.go:30-44
   public:
    virtual ~Writer() = default;
    virtual Result Write(const Samples &s) = 0;
    virtual SampleFormat Format() const = 0;  // Writer.SampleFormat()
    virtual unsigned SampleRate() const = 0;
};
using WriterPtr = std::shared_ptr<Writer>;

// A Writer that appends to a host buffer and records the size of every Write: the role the read end
// of sdr.Pipe plays in the reference's writer tests.
class BufferWriter : public Writer {
   public:
    BufferWriter(SampleFormat f, unsigned rate) : fmt_(f), rate_(rate) {}
    Result Write(const Samples &s) override {
        if (s.Format() != fmt_) return {0, ErrSampleFormatMismatch};  // pipe.go:101-103
        if (s.OnDevice()) return {0, make_err("BufferWriter: host samples expected")};
        const uint8_t *p = (const uint8_t *)s.Data();
        bytes_.insert(bytes_.end(), p, p + s.Size());
        writes_.push_back(s.Length());
        return {s.Length(), nullptr};
    }
    SampleFormat Format() const override { return fmt_; }
    unsigned SampleRate() const override { return rate_; }
    int Length() const { return (int)(bytes_.size() / FormatSize(fmt_)); }
    const std::vector<uint8_t> &Bytes() const { return bytes_; }
    const std::vector<int> &Writes() const { return writes_; }

   private:
    SampleFormat fmt_;
    unsigned rate_;
    std::vector<uint8_t> bytes_;
    std::vector<int> writes_;
};

// stream.ConvertWriter, stream/convert.go:58-118: a writer taking `input_format` samples that
// converts them (sdr.ConvertBuffer on the GPU, 32 Ki samples at a time like the reference's
// internal buffer) and passes them on to `out` in out's format.
class ConvertWriterGpu : public Writer {
   public:
    static constexpr int kBufSize = 32 * 1024;  // stream/convert.go:68
    ConvertWriterGpu(cuda::Context &ctx, WriterPtr out, SampleFormat input_format)
        : ctx_(ctx), out_(std::move(out)), in_fmt_(input_format), buf_(MakeSamples(out_->Format(), kBufSize)) {}
    Result Write(const Samples &in) override {
        if (in.Format() != in_fmt_) return {0, ErrSampleFormatMismatch};  // :86-88
        int n = 0;
        for (int i = 0; i < in.Length(); i += kBufSize) {  // :94-115
            const int ie = std::min(i + kBufSize, in.Length());
            Result c = ConvertBuffer(ctx_, *buf_, *in.Slice(i, ie));
            if (c.err) return {n, c.err};
            if (ie - i != c.n) return {n, make_err("ConvertWriter: Conversion mismatch")};
            Result w = out_->Write(*buf_->Slice(0, c.n));
            n += w.n;
            if (w.err) return {n, w.err};
        }
        return {n, nullptr};
    }
    SampleFormat Format() const override { return in_fmt_; }               // :120-122
    unsigned SampleRate() const override { return out_->SampleRate(); }    // :124-126

   private:
    cuda::Context &ctx_;
    WriterPtr out_;
    SampleFormat in_fmt_;
    SamplesPtr buf_;
};
inline std::pair<WriterPtr, Err> ConvertWriter(cuda::Context &ctx, WriterPtr out, SampleFormat input_format) {
    if (!MakeSamples(out->Format(), 1)) return {nullptr, ErrSampleFormatUnknown};  // MakeSamples' error, :69-72
    return {std::make_shared<ConvertWriterGpu>(ctx, std::move(out), input_format), nullptr};
}

}  // namespace sdr

// ---- fft ---------------------------------------------------------------------------------------
namespace fft {

enum Direction : int { Forward = HZSDR_FFT_FORWARD, Backward = HZSDR_FFT_BACKWARD };  // fft/fft.go:32-40

class Plan {  // fft/fft.go:52-59
   public:
    virtual ~Plan() = default;
    virtual sdr::Err Transform() = 0;
    virtual sdr::Err Close() = 0;
};
using PlanPtr = std::shared_ptr<Plan>;
// Planner(iq, frequency, direction): both device-resident complex64 buffers of equal length.
using Planner = std::function<std::pair<PlanPtr, sdr::Err>(std::shared_ptr<cuda::DeviceSamples> iq,
                                                           std::shared_ptr<cuda::DeviceSamples> frequency, Direction)>;

class CudaPlan : public Plan {
   public:
    CudaPlan(hzsdr_fft_plan *p, std::shared_ptr<cuda::DeviceSamples> iq, std::shared_ptr<cuda::DeviceSamples> fr, Direction d)
        : p_(p), iq_(std::move(iq)), fr_(std::move(fr)), d_(d) {}
    ~CudaPlan() override { Close(); }
    sdr::Err Transform() override {
        if (!p_) return sdr::make_err("fft: plan closed");
        // Forward reads iq and writes frequency; Backward the other way round (fft/fft.go:32-40)
        const void *src = d_ == Forward ? iq_->Data() : fr_->Data();
        void *dst = d_ == Forward ? fr_->Data() : iq_->Data();
        return sdr::from_status(hzsdr_fft_exec(p_, src, dst, 1));
    }
    sdr::Err Close() override {
        if (p_) hzsdr_fft_plan_destroy(p_);
        p_ = nullptr;
        return nullptr;
    }

   private:
    hzsdr_fft_plan *p_;
    std::shared_ptr<cuda::DeviceSamples> iq_, fr_;
    Direction d_;
};

// The GPU fft.Planner; length mismatch is sdr::ErrDstTooSmall (testutils/fft.go:127-138).
inline Planner CudaPlanner(cuda::ContextPtr ctx) {
    return [ctx](std::shared_ptr<cuda::DeviceSamples> iq, std::shared_ptr<cuda::DeviceSamples> fr, Direction d)
               -> std::pair<PlanPtr, sdr::Err> {
        hzsdr_fft_plan *p = nullptr;
        int rc = hzsdr_fft_plan_create(ctx->h(), iq->Length(), fr->Length(), (int)d, &p);
        if (rc != HZSDR_OK) return {nullptr, sdr::from_status(rc)};
        return {std::make_shared<CudaPlan>(p, iq, fr, d), nullptr};
    };
}

}  // namespace fft

// ---- stream ------------------------------------------------------------------------------------
namespace stream {

using sdr::ConvertWriter;  // stream.ConvertWriter lives next to sdr.ConvertBuffer above

using sdr::Err;
using sdr::Reader;
using sdr::ReaderPtr;
using sdr::Result;
using sdr::SampleFormat;
using sdr::Samples;

constexpr int kBlock = 32 * 1024;  // stream/convert.go:43-44, decimate.go:41-42, downsample.go:54-55

// A GPU stage can hand its output to the next stage without leaving the device.
class DeviceReader : public Reader {
   public:
    virtual cuda::ContextPtr Ctx() const = 0;
    // Fill up to dst.Length() samples of dst (device memory, this stage's output format).
    virtual Result ReadDevice(cuda::DeviceSamples &dst) = 0;
    Result Read(Samples &s) override {
        if (s.Format() != Format()) return {0, wrong_format_};
        if (s.OnDevice()) return ReadDevice(static_cast<cuda::DeviceSamples &>(s));
        if (!host_stage_ || host_stage_->Length() < s.Length())
            host_stage_ = std::make_shared<cuda::DeviceSamples>(Ctx(), Format(), s.Length());
        auto view = std::static_pointer_cast<cuda::DeviceSamples>(host_stage_->Slice(0, s.Length()));
        Result r = ReadDevice(*view);
        if (r.n > 0) {
            int rc = hzsdr_download(Ctx()->h(), s.Data(), view->Data(), (size_t)r.n * sdr::FormatSize(Format()));
            if (rc != HZSDR_OK) return {0, sdr::from_status(rc)};
        }
        return r;
    }

   protected:
    Err wrong_format_ = sdr::ErrSampleFormatMismatch;  // testutils/reader.go:87-97
    std::shared_ptr<cuda::DeviceSamples> host_stage_;
};

// Pull exactly n samples (ReadFull semantics) from `in` into device memory `dst`.
inline Result read_full_to_device(const cuda::ContextPtr &ctx, Reader &in, cuda::DeviceSamples &dst, sdr::SamplesPtr &host_stage) {
    if (auto *dr = dynamic_cast<DeviceReader *>(&in)) return sdr::ReadFull(*dr, dst);
    const int n = dst.Length();
    if (!host_stage || host_stage->Length() < n || host_stage->Format() != in.Format()) host_stage = sdr::MakeSamples(in.Format(), n);
    auto view = host_stage->Slice(0, n);
    Result r = sdr::ReadFull(in, *view);
    if (r.n > 0) {
        int rc = hzsdr_upload(ctx->h(), dst.Data(), view->Data(), (size_t)r.n * sdr::FormatSize(in.Format()));
        if (rc == HZSDR_OK) rc = hzsdr_ctx_sync(ctx->h());  // the staging buffer is reused by the next read
        if (rc != HZSDR_OK) return {0, sdr::from_status(rc)};
    }
    return r;
}

// stream.ReadTransformer (stream/read_transformer.go:45-137) with the Proc running on the GPU:
// ReadFull `in_len` input samples, Proc them into at most `out_len` output samples, hand those out.
// `batch` consecutive blocks go through one kernel launch; a trailing partial block is dropped and
// the error latched, as in the reference (:121-125).
class GpuReadTransformer : public DeviceReader {
   public:
    // proc(in_dev, n_blocks, out_dev) -> samples produced (or error)
    using Proc = std::function<Result(cuda::DeviceSamples &in, int n_blocks, cuda::DeviceSamples &out)>;
    GpuReadTransformer(cuda::ContextPtr ctx, ReaderPtr in, int in_len, int out_len, SampleFormat out_fmt, unsigned out_rate,
                       Proc proc, int batch = 16)
        : ctx_(std::move(ctx)), in_(std::move(in)), in_len_(in_len), out_len_(out_len), fmt_(out_fmt), rate_(out_rate),
          proc_(std::move(proc)), batch_(batch) {}
    cuda::ContextPtr Ctx() const override { return ctx_; }
    SampleFormat Format() const override { return fmt_; }
    unsigned SampleRate() const override { return rate_; }

    Result ReadDevice(cuda::DeviceSamples &dst) override {
        if (avail_ == 0) {
            if (err_) return {0, err_};
            fill();
            if (avail_ == 0) return {0, err_};
        }
        const int n = std::min(dst.Length(), avail_);
        int rc = hzsdr_copy(ctx_->h(), dst.Data(), (uint8_t *)out_->Data() + (size_t)pos_ * sdr::FormatSize(fmt_),
                            (size_t)n * sdr::FormatSize(fmt_));
        if (rc != HZSDR_OK) return {0, sdr::from_status(rc)};
        pos_ += n;
        avail_ -= n;
        return {n, nullptr};
    }

   private:
    void fill() {
        if (!inbuf_) {
            inbuf_ = std::make_shared<cuda::DeviceSamples>(ctx_, in_->Format(), in_len_ * batch_);
            out_ = std::make_shared<cuda::DeviceSamples>(ctx_, fmt_, out_len_ * batch_);
        }
        Result r = read_full_to_device(ctx_, *in_, *inbuf_, host_in_);
        const int blocks = r.n / in_len_;  // the partial block is dropped
        if (r.err) err_ = r.err;
        pos_ = 0;
        if (blocks > 0) {
            auto iv = std::static_pointer_cast<cuda::DeviceSamples>(inbuf_->Slice(0, blocks * in_len_));
            Result p = proc_(*iv, blocks, *out_);
            if (p.err) {
                err_ = p.err;
                return;
            }
            avail_ = p.n;
        }
    }
    cuda::ContextPtr ctx_;
    ReaderPtr in_;
    int in_len_, out_len_;
    SampleFormat fmt_;
    unsigned rate_;
    Proc proc_;
    int batch_;
    std::shared_ptr<cuda::DeviceSamples> inbuf_, out_;
    sdr::SamplesPtr host_in_;
    int pos_ = 0, avail_ = 0;
    Err err_;
};

// stream.ConvertReader, stream/convert.go:37-51 (to = C64)
inline std::pair<ReaderPtr, Err> ConvertReader(cuda::ContextPtr ctx, ReaderPtr in, SampleFormat to) {
    if (to != SampleFormat::C64) return {nullptr, sdr::ErrConversionNotImplemented};
    const SampleFormat from = in->Format();
    auto c = ctx;
    auto proc = [c, from](cuda::DeviceSamples &i, int blocks, cuda::DeviceSamples &o) -> Result {
        size_t got = 0;
        int rc = hzsdr_convert_to_c64(c->h(), (int)from, i.Data(), (size_t)blocks * kBlock, o.Data(), o.Length(), &got);
        return {(int)got, sdr::from_status(rc)};
    };
    const unsigned rate = in->SampleRate();
    return {std::make_shared<GpuReadTransformer>(ctx, std::move(in), kBlock, kBlock, SampleFormat::C64, rate, proc), nullptr};
}

// stream.ShiftReader, stream/shifter.go:89-102: mixes whatever block the consumer asks for; the
// fp64 time accumulator is carried across reads (shifter.go:66-85).
class ShiftReaderGpu : public DeviceReader {
   public:
    ShiftReaderGpu(cuda::ContextPtr ctx, ReaderPtr r, double shift_hz) : ctx_(std::move(ctx)), r_(std::move(r)), shift_(shift_hz) {
        nco_.sample_rate = r_->SampleRate();
        nco_.ts = 0.0;
        wrong_format_ = sdr::ErrSampleFormatUnknown;  // shifter.go:45-50
    }
    cuda::ContextPtr Ctx() const override { return ctx_; }
    SampleFormat Format() const override { return r_->Format(); }
    unsigned SampleRate() const override { return r_->SampleRate(); }
    double Ts() const { return nco_.ts; }
    Result ReadDevice(cuda::DeviceSamples &dst) override {
        Result r;
        if (auto *dr = dynamic_cast<DeviceReader *>(r_.get()))
            r = dr->ReadDevice(dst);
        else {
            if (!host_ || host_->Length() < dst.Length()) host_ = sdr::MakeSamples(SampleFormat::C64, dst.Length());
            auto v = host_->Slice(0, dst.Length());
            r = r_->Read(*v);
            if (r.n > 0) {
                int rc = hzsdr_upload(ctx_->h(), dst.Data(), v->Data(), (size_t)r.n * 8);
                if (rc == HZSDR_OK) rc = hzsdr_ctx_sync(ctx_->h());
                if (rc != HZSDR_OK) return {0, sdr::from_status(rc)};
            }
        }
        if (r.err) return r;  // shifter.go:52-55
        int rc = hzsdr_shift(ctx_->h(), dst.Data(), r.n, shift_, &nco_);
        if (rc != HZSDR_OK) return {0, sdr::from_status(rc)};
        return r;
    }

   private:
    cuda::ContextPtr ctx_;
    ReaderPtr r_;
    double shift_;
    hzsdr_nco nco_{};
    sdr::SamplesPtr host_;
};
inline std::pair<ReaderPtr, Err> ShiftReader(cuda::ContextPtr ctx, ReaderPtr r, double shift_hz) {
    if (r->Format() != SampleFormat::C64) return {nullptr, sdr::ErrSampleFormatUnknown};  // shifter.go:90-95
    return {std::make_shared<ShiftReaderGpu>(std::move(ctx), std::move(r), shift_hz), nullptr};
}

// stream.ConvolutionReader, stream/convolution.go:36-82.  `filter` is in the frequency domain;
// blocks of len(filter) samples, block-circular.  The planner argument of the reference is
// implied: the library's own FFT (fft::CudaPlanner) is fused into the kernel.
inline std::pair<ReaderPtr, Err> ConvolutionReader(cuda::ContextPtr ctx, ReaderPtr r, const std::vector<std::complex<float>> &filter) {
    if (r->Format() != SampleFormat::C64) return {nullptr, sdr::ErrSampleFormatUnknown};
    const int n = (int)filter.size();
    auto fdev = std::make_shared<cuda::DeviceSamples>(ctx, SampleFormat::C64, n);
    int rc = hzsdr_upload(ctx->h(), fdev->Data(), filter.data(), (size_t)n * 8);
    if (rc == HZSDR_OK) rc = hzsdr_ctx_sync(ctx->h());
    if (rc != HZSDR_OK) return {nullptr, sdr::from_status(rc)};
    // probe the length now so an unsupported size fails at construction like fft.ConvolveFreq does
    rc = hzsdr_convolve_freq(ctx->h(), nullptr, nullptr, fdev->Data(), n, 0);
    if (rc != HZSDR_OK) return {nullptr, sdr::from_status(rc)};
    auto c = ctx;
    auto proc = [c, fdev, n](cuda::DeviceSamples &i, int blocks, cuda::DeviceSamples &o) -> Result {
        int rc2 = hzsdr_convolve_freq(c->h(), i.Data(), o.Data(), fdev->Data(), n, blocks);
        return {blocks * n, sdr::from_status(rc2)};
    };
    const unsigned rate = r->SampleRate();
    const int batch = std::max(1, (1 << 19) / n);
    return {std::make_shared<GpuReadTransformer>(ctx, std::move(r), n, n, SampleFormat::C64, rate, proc, batch), nullptr};
}

// stream.DecimateReader, stream/decimate.go:34-51
inline std::pair<ReaderPtr, Err> DecimateReader(cuda::ContextPtr ctx, ReaderPtr in, unsigned factor) {
    const SampleFormat f = in->Format();
    auto c = ctx;
    auto proc = [c, f, factor](cuda::DeviceSamples &i, int blocks, cuda::DeviceSamples &o) -> Result {
        size_t got = 0;
        int rc = hzsdr_decimate(c->h(), (int)f, i.Data(), (size_t)blocks * kBlock, o.Data(), o.Length(), factor, kBlock, &got);
        return {(int)got, sdr::from_status(rc)};
    };
    const unsigned rate = in->SampleRate() / factor;  // decimate.go:43
    return {std::make_shared<GpuReadTransformer>(ctx, std::move(in), kBlock, kBlock, f, rate, proc), nullptr};
}

// stream.DownsampleReader, stream/downsample.go:47-64 (output always C64)
inline std::pair<ReaderPtr, Err> DownsampleReader(cuda::ContextPtr ctx, ReaderPtr in, unsigned factor) {
    const SampleFormat f = in->Format();
    auto c = ctx;
    auto proc = [c, f, factor](cuda::DeviceSamples &i, int blocks, cuda::DeviceSamples &o) -> Result {
        size_t got = 0;
        int rc = hzsdr_downsample(c->h(), (int)f, i.Data(), (size_t)blocks * kBlock, o.Data(), o.Length(), factor, kBlock, &got);
        return {(int)got, sdr::from_status(rc)};
    };
    const unsigned rate = in->SampleRate() / factor;
    return {std::make_shared<GpuReadTransformer>(ctx, std::move(in), kBlock, kBlock, SampleFormat::C64, rate, proc), nullptr};
}

// stream.Multiply on complex64 (stream/multiply.go:46-89) and stream.Gain (stream/gain.go:30-57):
// in-place on whatever the upstream delivered.
class PointwiseReaderGpu : public DeviceReader {
   public:
    enum Kind { kMultiply, kGain };
    PointwiseReaderGpu(cuda::ContextPtr ctx, ReaderPtr r, Kind k, std::complex<float> m) : ctx_(std::move(ctx)), r_(std::move(r)), kind_(k), m_(m) {}
    void SetMultiplier(std::complex<float> m) { m_ = m; }  // multiply.go:34-36
    cuda::ContextPtr Ctx() const override { return ctx_; }
    SampleFormat Format() const override { return r_->Format(); }
    unsigned SampleRate() const override { return r_->SampleRate(); }
    Result ReadDevice(cuda::DeviceSamples &dst) override {
        Result r;
        if (auto *dr = dynamic_cast<DeviceReader *>(r_.get()))
            r = dr->ReadDevice(dst);
        else {
            if (!host_ || host_->Length() < dst.Length()) host_ = sdr::MakeSamples(SampleFormat::C64, dst.Length());
            auto v = host_->Slice(0, dst.Length());
            r = r_->Read(*v);
            if (r.n > 0) {
                int rc = hzsdr_upload(ctx_->h(), dst.Data(), v->Data(), (size_t)r.n * 8);
                if (rc == HZSDR_OK) rc = hzsdr_ctx_sync(ctx_->h());
                if (rc != HZSDR_OK) return {0, sdr::from_status(rc)};
            }
        }
        if (r.err) return r;
        int rc = HZSDR_OK;
        if (kind_ == kGain)
            rc = hzsdr_scale(ctx_->h(), dst.Data(), r.n, m_.real());
        else if (m_ != std::complex<float>(1.f, 0.f))  // multiply.go:59-62
            rc = hzsdr_rotate(ctx_->h(), dst.Data(), r.n, m_.real(), m_.imag());
        if (rc != HZSDR_OK) return {0, sdr::from_status(rc)};
        return r;
    }

   private:
    cuda::ContextPtr ctx_;
    ReaderPtr r_;
    Kind kind_;
    std::complex<float> m_;
    sdr::SamplesPtr host_;
};
// stream.Multiply on U8 / I8 streams (stream/multiply.go:91-251): every Read is a 65536-entry table
// lookup; SetMultiplier rebuilds the table by Convert -> Multiply -> Convert, on the GPU, exactly the
// way the reference builds it (the u8 reader's x0*255 + x1 index and its collisions included: its
// 65535-entry table is re-indexed into the library's little-endian pair index).
class LutMultiplyReaderGpu : public DeviceReader {
   public:
    LutMultiplyReaderGpu(cuda::ContextPtr ctx, ReaderPtr r, std::complex<float> m) : ctx_(std::move(ctx)), r_(std::move(r)) {
        table_ = std::make_shared<cuda::DeviceSamples>(ctx_, r_->Format(), 65536);
        err_ = SetMultiplier(m);
    }
    Err SetMultiplier(std::complex<float> m) {
        const SampleFormat f = r_->Format();
        const bool u8 = f == SampleFormat::U8;
        const int entries = u8 ? 65535 : 65536;
        std::vector<std::array<uint8_t, 2>> ident(entries);
        if (u8) {  // multiply.go:153-160: real 0..255, imag 0..256 (uint8 wraps), later writes win
            for (int re = 0; re < 256; re++)
                for (int im = 0; im <= 256; im++) ident[re * 255 + (im & 0xff)] = {(uint8_t)re, (uint8_t)im};
        } else {  // LookupTableIdentityI8, iq_lookup_table.go:82-90
            for (int i = 0; i < 65536; i++) ident[i] = {(uint8_t)(i & 0xff), (uint8_t)(i >> 8)};
        }
        cuda::DeviceSamples raw(ctx_, f, entries), c64(ctx_, SampleFormat::C64, entries), back(ctx_, f, entries);
        int rc = hzsdr_upload(ctx_->h(), raw.Data(), ident.data(), (size_t)entries * 2);
        if (rc == HZSDR_OK) rc = hzsdr_ctx_sync(ctx_->h());
        size_t got = 0;
        if (rc == HZSDR_OK) rc = hzsdr_convert(ctx_->h(), (int)f, raw.Data(), entries, HZSDR_FORMAT_C64, c64.Data(), entries, &got);
        if (rc == HZSDR_OK) rc = hzsdr_rotate(ctx_->h(), c64.Data(), entries, m.real(), m.imag());  // cbuf.Multiply(m): no m == 1 shortcut here
        if (rc == HZSDR_OK) rc = hzsdr_convert(ctx_->h(), HZSDR_FORMAT_C64, c64.Data(), entries, (int)f, back.Data(), entries, &got);
        std::vector<std::array<uint8_t, 2>> tab(entries), full(65536);
        if (rc == HZSDR_OK) rc = hzsdr_download(ctx_->h(), tab.data(), back.Data(), (size_t)entries * 2);
        if (rc != HZSDR_OK) return sdr::from_status(rc);
        for (int i = 0; i < 65536; i++) full[i] = u8 ? tab[(i & 0xff) * 255 + (i >> 8)] : tab[i];
        rc = hzsdr_upload(ctx_->h(), table_->Data(), full.data(), 65536 * 2);
        if (rc == HZSDR_OK) rc = hzsdr_ctx_sync(ctx_->h());
        return sdr::from_status(rc);
    }
    cuda::ContextPtr Ctx() const override { return ctx_; }
    SampleFormat Format() const override { return r_->Format(); }
    unsigned SampleRate() const override { return r_->SampleRate(); }
    Result ReadDevice(cuda::DeviceSamples &dst) override {
        if (err_) return {0, err_};
        if (!in_ || in_->Length() < dst.Length()) in_ = std::make_shared<cuda::DeviceSamples>(ctx_, Format(), dst.Length());
        auto v = std::static_pointer_cast<cuda::DeviceSamples>(in_->Slice(0, dst.Length()));
        Result r;
        if (auto *dr = dynamic_cast<DeviceReader *>(r_.get()))
            r = dr->ReadDevice(*v);
        else {
            if (!host_ || host_->Length() < dst.Length()) host_ = sdr::MakeSamples(Format(), dst.Length());
            auto hv = host_->Slice(0, dst.Length());
            r = r_->Read(*hv);
            if (r.n > 0) {
                int rc = hzsdr_upload(ctx_->h(), v->Data(), hv->Data(), (size_t)r.n * 2);
                if (rc == HZSDR_OK) rc = hzsdr_ctx_sync(ctx_->h());
                if (rc != HZSDR_OK) return {0, sdr::from_status(rc)};
            }
        }
        if (r.err) return r;
        int rc = hzsdr_lookup(ctx_->h(), (int)Format(), v->Data(), r.n, (int)Format(), table_->Data(), dst.Data(), dst.Length());
        if (rc != HZSDR_OK) return {0, sdr::from_status(rc)};
        return r;
    }

   private:
    cuda::ContextPtr ctx_;
    ReaderPtr r_;
    std::shared_ptr<cuda::DeviceSamples> table_, in_;
    sdr::SamplesPtr host_;
    Err err_;
};

// stream.Multiply, stream/multiply.go:74-89
inline std::pair<ReaderPtr, Err> Multiply(cuda::ContextPtr ctx, ReaderPtr r, std::complex<float> m) {
    switch (r->Format()) {
        case SampleFormat::C64: return {std::make_shared<PointwiseReaderGpu>(std::move(ctx), std::move(r), PointwiseReaderGpu::kMultiply, m), nullptr};
        case SampleFormat::U8:
        case SampleFormat::I8: return {std::make_shared<LutMultiplyReaderGpu>(std::move(ctx), std::move(r), m), nullptr};
        default: return {nullptr, sdr::ErrSampleFormatUnknown};
    }
}
inline ReaderPtr Gain(cuda::ContextPtr ctx, ReaderPtr r, float v) {
    return std::make_shared<PointwiseReaderGpu>(std::move(ctx), std::move(r), PointwiseReaderGpu::kGain, std::complex<float>(v, 0.f));
}

// stream.Add, stream/add.go:41-185 (complex64): per Read, ReadFull every reader, zero, add in order.
class AddReaderGpu : public DeviceReader {
   public:
    AddReaderGpu(cuda::ContextPtr ctx, std::vector<ReaderPtr> rs) : ctx_(std::move(ctx)), rs_(std::move(rs)) {
        wrong_format_ = sdr::ErrSampleFormatUnknown;  // add.go:129-135
    }
    cuda::ContextPtr Ctx() const override { return ctx_; }
    SampleFormat Format() const override { return rs_[0]->Format(); }
    unsigned SampleRate() const override { return rs_[0]->SampleRate(); }
    Result ReadDevice(cuda::DeviceSamples &dst) override {
        if (err_) return {0, err_};
        const int n = dst.Length();
        if ((int)bufs_.size() != (int)rs_.size() || (bufs_.size() && bufs_[0]->Length() < n)) {
            bufs_.clear();
            for (size_t i = 0; i < rs_.size(); i++) bufs_.push_back(std::make_shared<cuda::DeviceSamples>(ctx_, Format(), n));
        }
        std::vector<const void *> ptrs;
        for (size_t i = 0; i < rs_.size(); i++) {
            auto v = std::static_pointer_cast<cuda::DeviceSamples>(bufs_[i]->Slice(0, n));
            Result r = read_full_to_device(ctx_, *rs_[i], *v, host_);
            if (r.err) {  // add.go:148-158: latch, return 0
                err_ = r.err;
                return {0, err_};
            }
            ptrs.push_back(v->Data());
        }
        int rc = Format() == SampleFormat::C64 ? hzsdr_add(ctx_->h(), dst.Data(), ptrs.data(), (int)ptrs.size(), n)
                                               : hzsdr_add_int(ctx_->h(), (int)Format(), dst.Data(), ptrs.data(), (int)ptrs.size(), n);
        if (rc != HZSDR_OK) return {0, sdr::from_status(rc)};
        return {n, nullptr};
    }

   private:
    cuda::ContextPtr ctx_;
    std::vector<ReaderPtr> rs_;
    std::vector<std::shared_ptr<cuda::DeviceSamples>> bufs_;
    sdr::SamplesPtr host_;
    Err err_;
};
inline std::pair<ReaderPtr, Err> Add(cuda::ContextPtr ctx, std::vector<ReaderPtr> readers) {
    if (readers.empty()) return {nullptr, sdr::make_err("stream.Add: No readers passed")};  // add.go:44-45
    if (readers.size() == 1) return {readers[0], nullptr};                                 // add.go:46-47
    if (readers[0]->Format() == SampleFormat::U8) return {nullptr, sdr::ErrSampleFormatUnknown};  // add.go:56-61: C64, I16, I8 only
    for (auto &r : readers) {
        if (r->Format() != readers[0]->Format()) return {nullptr, sdr::make_err("stream.Add: Readers are not all the same format")};
        if (r->SampleRate() != readers[0]->SampleRate()) return {nullptr, sdr::make_err("stream.Add: Readers are not all the same rate")};
    }
    return {std::make_shared<AddReaderGpu>(std::move(ctx), std::move(readers)), nullptr};
}

// stream.BeamformAngles2D / BeamformAngles, stream/beamform.go:57-128
inline std::vector<std::complex<float>> BeamformAngles2D(double frequency_hz, double angle_deg, std::array<double, 2> center,
                                                         const std::vector<std::array<double, 2>> &antennas) {
    std::vector<std::complex<float>> out(antennas.size());
    if (antennas.empty()) return out;
    hzsdr_beamform_angles_2d(frequency_hz, angle_deg, center.data(), &antennas[0][0], (int)antennas.size(), reinterpret_cast<float *>(out.data()));
    return out;
}
inline std::vector<std::complex<float>> BeamformAngles(double frequency_hz, double angle_deg, const std::vector<double> &distances) {
    std::vector<std::array<double, 2>> ants;
    for (double d : distances) ants.push_back({d, 0.0});
    if (ants.empty()) return {};
    return BeamformAngles2D(frequency_hz, angle_deg, ants[0], ants);
}

// stream.ReadBeamform, stream/beamform.go:148-171.  Raw coherent readers -> one complex64 beam.
// The reference builds N ConvertReaders + N Multiply readers + an Add; here the three collapse into
// one kernel per block (K8) reading every raw channel once.
struct BeamformConfig {
    std::vector<std::complex<float>> Angles;
};
class Beamform : public DeviceReader {
   public:
    Beamform(cuda::ContextPtr ctx, std::vector<ReaderPtr> rs, BeamformConfig cfg) : ctx_(std::move(ctx)), rs_(std::move(rs)), w_(std::move(cfg.Angles)) {
        wrong_format_ = sdr::ErrSampleFormatUnknown;
    }
    Err SetPhaseAngles(const std::vector<std::complex<float>> &angles) {  // beamform.go:131-139
        if (angles.size() != rs_.size()) return sdr::make_err("Beamform.SetPhaseAngles: angles must match the reader length");
        w_ = angles;
        return nullptr;
    }
    cuda::ContextPtr Ctx() const override { return ctx_; }
    SampleFormat Format() const override { return SampleFormat::C64; }
    unsigned SampleRate() const override { return rs_[0]->SampleRate(); }
    Result ReadDevice(cuda::DeviceSamples &dst) override {
        if (err_) return {0, err_};
        // ConvertReader granularity: whole 32768-sample blocks (stream/convert.go:43-44)
        const int n = (dst.Length() / kBlock) * kBlock;
        if (n == 0) return {0, sdr::ErrShortBuffer};
        const SampleFormat f = rs_[0]->Format();
        if (raw_.size() != rs_.size() || raw_[0]->Length() < n) {
            raw_.clear();
            for (size_t i = 0; i < rs_.size(); i++) raw_.push_back(std::make_shared<cuda::DeviceSamples>(ctx_, f, n));
        }
        std::vector<const void *> ptrs;
        for (size_t i = 0; i < rs_.size(); i++) {
            auto v = std::static_pointer_cast<cuda::DeviceSamples>(raw_[i]->Slice(0, n));
            Result r = read_full_to_device(ctx_, *rs_[i], *v, host_);
            if (r.err) {
                err_ = r.err;
                return {0, err_};
            }
            ptrs.push_back(v->Data());
        }
        int rc = hzsdr_beamform(ctx_->h(), (int)f, ptrs.data(), (int)ptrs.size(), reinterpret_cast<const float *>(w_.data()), n, dst.Data());
        if (rc != HZSDR_OK) return {0, sdr::from_status(rc)};
        return {n, nullptr};
    }

   private:
    cuda::ContextPtr ctx_;
    std::vector<ReaderPtr> rs_;
    std::vector<std::complex<float>> w_;
    std::vector<std::shared_ptr<cuda::DeviceSamples>> raw_;
    sdr::SamplesPtr host_;
    Err err_;
};
inline std::pair<std::shared_ptr<Beamform>, Err> ReadBeamform(cuda::ContextPtr ctx, std::vector<ReaderPtr> rs, BeamformConfig cfg) {
    if (rs.empty()) return {nullptr, sdr::make_err("stream.Add: No readers passed")};
    for (auto &r : rs)
        if (r->Format() == SampleFormat::C64 || r->Format() != rs[0]->Format() || r->SampleRate() != rs[0]->SampleRate())
            return {nullptr, sdr::make_err("stream.ReadBeamform: readers must share one raw format and rate")};
    if (cfg.Angles.size() != rs.size()) cfg.Angles.assign(rs.size(), std::complex<float>(1.f, 0.f));  // Multiply(reader, 1), beamform.go:155
    return {std::make_shared<Beamform>(std::move(ctx), std::move(rs), std::move(cfg)), nullptr};
}

// The fused chain as a Reader: what DecimateReader(ConvolutionReader(ShiftReader(ConvertReader(raw))))
// collapses to under `sdr.cuda` when every stage is GPU-backed.  Block rules are the composition of
// the four ReadTransformers: whole 32768-sample blocks of raw input only.
class ChainReader : public DeviceReader {
   public:
    ChainReader(cuda::ContextPtr ctx, ReaderPtr raw, double shift_hz, const std::vector<std::complex<float>> &filter, unsigned factor,
                int blocks_per_launch = 128)
        : ctx_(std::move(ctx)), raw_(std::move(raw)), factor_(factor), batch_(blocks_per_launch) {
        hzsdr_chain_config cfg{};
        cfg.src_format = (int)raw_->Format();
        cfg.sample_rate = raw_->SampleRate();
        cfg.shift_hz = shift_hz;
        cfg.n_fft = filter.size();
        cfg.filter_host = filter.data();
        cfg.decimate = factor;
        int rc = hzsdr_chain_create(ctx_->h(), &cfg, &chain_);
        if (rc != HZSDR_OK) create_err_ = sdr::from_status(rc);
        unit_ = std::max<int>(kBlock, (int)filter.size());
    }
    ~ChainReader() override { hzsdr_chain_destroy(chain_); }
    Err CreateError() const { return create_err_; }
    cuda::ContextPtr Ctx() const override { return ctx_; }
    SampleFormat Format() const override { return SampleFormat::C64; }
    unsigned SampleRate() const override { return raw_->SampleRate() / factor_; }
    double Ts() const {
        double ts = 0;
        hzsdr_chain_get_ts(chain_, &ts);
        return ts;
    }
    Result ReadDevice(cuda::DeviceSamples &dst) override {
        if (avail_ == 0) {
            if (err_) return {0, err_};
            if (!in_) {
                in_ = std::make_shared<cuda::DeviceSamples>(ctx_, raw_->Format(), unit_ * batch_);
                size_t cap = 0;
                hzsdr_chain_out_len(chain_, (size_t)unit_ * batch_, &cap);
                out_ = std::make_shared<cuda::DeviceSamples>(ctx_, SampleFormat::C64, (int)cap);
            }
            Result r = read_full_to_device(ctx_, *raw_, *in_, host_);
            if (r.err) err_ = r.err;
            const int n = (r.n / unit_) * unit_;
            pos_ = 0;
            if (n > 0) {
                size_t got = 0;
                int rc = hzsdr_chain_exec(chain_, in_->Data(), n, out_->Data(), out_->Length(), &got);
                if (rc != HZSDR_OK) {
                    err_ = sdr::from_status(rc);
                    return {0, err_};
                }
                avail_ = (int)got;
            }
            if (avail_ == 0) return {0, err_};
        }
        const int n = std::min(dst.Length(), avail_);
        int rc = hzsdr_copy(ctx_->h(), dst.Data(), (uint8_t *)out_->Data() + (size_t)pos_ * 8, (size_t)n * 8);
        if (rc != HZSDR_OK) return {0, sdr::from_status(rc)};
        pos_ += n;
        avail_ -= n;
        return {n, nullptr};
    }

   private:
    cuda::ContextPtr ctx_;
    ReaderPtr raw_;
    unsigned factor_;
    int batch_, unit_ = kBlock;
    hzsdr_chain *chain_ = nullptr;
    std::shared_ptr<cuda::DeviceSamples> in_, out_;
    sdr::SamplesPtr host_;
    int pos_ = 0, avail_ = 0;
    Err err_, create_err_;
};

}  // namespace stream
