// {{{ Copyright (c) the hzsdr-cuda authors, MIT (same terms as hz.tools/sdr) }}}

//go:build sdr.cuda

// Package cuda is the public face of the B200 backend of hz.tools/sdr: a device context, a
// device-resident SamplesC64 that implements sdr.Samples, pinned host buffers that look like
// ordinary SamplesU8/I8/I16 (so SDR drivers can DMA straight into them), the ring-buffer allocator
// hook, and an fft.Planner.  The stream.* constructors pick these up under `-tags sdr.cuda`.
//
// Written against include/hzsdr_cuda.h; not compilable in the authoring image (no Go toolchain).
package cuda

import (
	"fmt"
	"sync"
	"unsafe"

	"hz.tools/sdr"
	"hz.tools/sdr/fft"
	"hz.tools/sdr/internal/hzcuda"
	"hz.tools/sdr/yikes"
)

// Context is one B200 and one CUDA stream.  Safe to use from any goroutine; calls on one Context
// are serialised by mu because the library's streams are in-order anyway.
type Context struct {
	mu  sync.Mutex
	ctx *hzcuda.Ctx
}

var (
	defaultOnce sync.Once
	defaultCtx  *Context
	defaultErr  error
)

// Default returns the process-wide context on device 0 that the stream.* constructors use when the
// caller did not pick one.  No GPU => the error is returned by every constructor: no CPU fallback.
func Default() (*Context, error) {
	defaultOnce.Do(func() { defaultCtx, defaultErr = New(0) })
	return defaultCtx, defaultErr
}

// New opens device `device`.
func New(device int) (*Context, error) {
	c, err := hzcuda.NewCtx(device)
	if err != nil {
		return nil, translate(err)
	}
	return &Context{ctx: c}, nil
}

// Raw exposes the binding for the stream package twins.
func (c *Context) Raw() *hzcuda.Ctx { return c.ctx }

// Translate is exported for the stream package twins.
func Translate(err error) error { return translate(err) }

// translate maps library status codes onto hz.tools/sdr's sentinel errors so that callers keep
// comparing with == (iq.go:27-39, conv.go:30).
func translate(err error) error {
	if e, ok := err.(*hzcuda.Error); ok {
		switch e.Status {
		case hzcuda.ErrDstTooSmall:
			return sdr.ErrDstTooSmall
		case hzcuda.ErrFormatMismatch:
			return sdr.ErrSampleFormatMismatch
		case hzcuda.ErrFormatUnknown:
			return sdr.ErrSampleFormatUnknown
		case hzcuda.ErrConversionNotImplmented:
			return sdr.ErrConversionNotImplemented
		}
	}
	return err
}

// SamplesC64 is a vector of complex64 samples resident in HBM.  It reports
// sdr.SampleFormatC64, so every Reader format check passes; the `sdr.cuda` twins of CopySamples
// and ConvertBuffer know how to move data in and out of it (copy_cuda.go, conv_cuda.go).
type SamplesC64 struct {
	ctx  *Context
	base *deviceBlock // keeps the allocation alive while slices exist
	ptr  unsafe.Pointer
	n    int
}

type deviceBlock struct {
	ctx *Context
	ptr unsafe.Pointer
}

// MakeSamplesC64 allocates n samples of device memory.
func (c *Context) MakeSamplesC64(n int) (*SamplesC64, error) {
	p, err := c.ctx.Alloc(n * 8)
	if err != nil {
		return nil, translate(err)
	}
	return &SamplesC64{ctx: c, base: &deviceBlock{ctx: c, ptr: p}, ptr: p, n: n}, nil
}

func (s *SamplesC64) Format() sdr.SampleFormat { return sdr.SampleFormatC64 }
func (s *SamplesC64) Size() int                 { return s.n * 8 }
func (s *SamplesC64) Length() int               { return s.n }
func (s *SamplesC64) Slice(start, end int) sdr.Samples {
	return &SamplesC64{ctx: s.ctx, base: s.base, ptr: unsafe.Add(s.ptr, start*8), n: end - start}
}

// DevicePointer is how the root package recognises device samples without importing this package.
func (s *SamplesC64) DevicePointer() (unsafe.Pointer, *hzcuda.Ctx) { return s.ptr, s.ctx.ctx }

// Free releases the allocation (all slices become invalid).
func (s *SamplesC64) Free() error { return s.ctx.ctx.Free(s.base.ptr) }

// ToHost copies into an ordinary sdr.SamplesC64.
func (s *SamplesC64) ToHost(dst sdr.SamplesC64) (int, error) {
	if dst.Length() < s.n {
		return 0, sdr.ErrDstTooSmall
	}
	b, _ := sdr.UnsafeSamplesAsBytes(dst[:s.n])
	return s.n, translate(s.ctx.ctx.Download(b, s.ptr))
}

// PinnedSamples allocates cudaHostAlloc'd memory and presents it as an ordinary host Samples of
// the requested format via yikes.Samples (yikes/bytes.go:50-71) -- the precedent is uhd/rx.go:237,
// which wraps C-malloc'd buffers the same way.  Drivers write into it like any other buffer; the
// library can DMA from it asynchronously.
func PinnedSamples(format sdr.SampleFormat, n int) (sdr.Samples, func() error, error) {
	p, err := hzcuda.PinnedAlloc(n * format.Size())
	if err != nil {
		return nil, nil, translate(err)
	}
	s, err := yikes.Samples(uintptr(p), n*format.Size(), format)
	if err != nil {
		hzcuda.PinnedFree(p)
		return nil, nil, err
	}
	return s, func() error { return hzcuda.PinnedFree(p) }, nil
}

// Planner is an fft.Planner (fft/fft.go:45-48) running on the GPU.  `iq` and `frequency` must be
// device-resident (*SamplesC64 sliced to []complex64 is not possible), so this planner accepts the
// host slices the reference API prescribes and stages them: Transform uploads, runs, downloads.
// Pipelines that stay on the device use stream.ConvolutionReader, which fuses the transforms.
func (c *Context) Planner() fft.Planner {
	return func(iq sdr.SamplesC64, frequency []complex64, direction fft.Direction) (fft.Plan, error) {
		p, err := c.ctx.NewPlan(len(iq), len(frequency), direction == fft.Forward)
		if err != nil {
			return nil, translate(err) // length mismatch -> sdr.ErrDstTooSmall (testutils/fft.go:127-138)
		}
		n := len(iq)
		src, err := c.ctx.Alloc(n * 8)
		if err != nil {
			return nil, translate(err)
		}
		dst, err := c.ctx.Alloc(n * 8)
		if err != nil {
			return nil, translate(err)
		}
		return &plan{c: c, p: p, iq: iq, freq: frequency, fwd: direction == fft.Forward, src: src, dst: dst}, nil
	}
}

type plan struct {
	c        *Context
	p        *hzcuda.Plan
	iq       sdr.SamplesC64
	freq     []complex64
	fwd      bool
	src, dst unsafe.Pointer
}

func (p *plan) Transform() error {
	in, out := []complex64(p.iq), p.freq
	if !p.fwd {
		in, out = p.freq, []complex64(p.iq)
	}
	ib := unsafe.Slice((*byte)(unsafe.Pointer(&in[0])), len(in)*8)
	ob := unsafe.Slice((*byte)(unsafe.Pointer(&out[0])), len(out)*8)
	if err := p.c.ctx.UploadGo(p.src, ib); err != nil {
		return translate(err)
	}
	if err := p.p.Exec(p.src, p.dst, 1); err != nil {
		return translate(err)
	}
	return translate(p.c.ctx.Download(ob, p.dst))
}

func (p *plan) Close() error {
	p.c.ctx.Free(p.src)
	p.c.ctx.Free(p.dst)
	return p.p.Close()
}

// ---- device-resident forms of fft.Convolve / fft.CrossCorrelate and of rtl/kerberos/internal ----
//
// The reference's fft.Convolve (fft/convolution.go:97-139) works unchanged with Planner() above --
// three staged transforms and a host loop.  Pipelines whose buffers already live on the device call
// these instead: one library call, nothing crosses PCIe.

// Convolve writes IFFT(FFT(iq1) * FFT(iq2)) into dst (fft.Convolve), or with conj(FFT(iq2)) when
// crossCorrelate is set (fft.CrossCorrelate).  All three hold `batch` vectors of n samples; the
// transforms are unnormalised, as with any fft.Planner.  dst may be iq1.
func (c *Context) Convolve(dst, iq1, iq2 *SamplesC64, n, batch int, crossCorrelate bool) error {
	if iq1.Length() != iq2.Length() || iq1.Length() != dst.Length() || n*batch != dst.Length() {
		return fmt.Errorf("sdr/fft: IQ/Dest buffer lengths do not match exactly") // fft/convolution.go:36-41
	}
	c.mu.Lock()
	defer c.mu.Unlock()
	scratch, err := c.ctx.Alloc(n * batch * 8)
	if err != nil {
		return translate(err)
	}
	defer c.ctx.Free(scratch)
	if err := c.ctx.FftConvolve(dst.ptr, iq1.ptr, iq2.ptr, n, batch, crossCorrelate, scratch); err != nil {
		return translate(err)
	}
	return translate(c.ctx.Sync()) // scratch is freed on return
}

// FFTShiftAndScale is rtl/kerberos/internal.FFTShiftAndScale (reader.go:57-64) on `batch` vectors of n.
func (c *Context) FFTShiftAndScale(data *SamplesC64, n, batch int, scale float32) error {
	c.mu.Lock()
	defer c.mu.Unlock()
	return translate(c.ctx.FftShiftScale(data.ptr, n, batch, scale))
}

// Graft is the body of GraftReaders' loop (rtl/kerberos/internal/graft.go:96-125): iq holds
// nReaders buffers of fftSize samples, dst and freq nReaders*fftSize each.
func (c *Context) Graft(iq *SamplesC64, nReaders, fftSize int, dst, freq *SamplesC64) error {
	if iq.Length() != nReaders*fftSize || dst.Length() != iq.Length() || freq.Length() != iq.Length() {
		return sdr.ErrDstTooSmall
	}
	c.mu.Lock()
	defer c.mu.Unlock()
	return translate(c.ctx.Graft(iq.ptr, nReaders, fftSize, dst.ptr, freq.ptr))
}

// CorrelationOffsets is checkAlignment's peak search (rtl/kerberos/internal/align.go:125-146) over
// `batch` cross-correlation vectors of n samples: one signed sample offset per vector.
func (c *Context) CorrelationOffsets(cc *SamplesC64, n, batch int) ([]int32, error) {
	c.mu.Lock()
	defer c.mu.Unlock()
	out, err := c.ctx.CorrelatePeak(cc.ptr, n, batch)
	return out, translate(err)
}

// PhaseOffsets is rtl/kerberos/internal.PhaseOffsets (align.go:244-272) over nChan device buffers of
// n samples laid out one after the other.
func (c *Context) PhaseOffsets(bufs *SamplesC64, nChan, n int) ([]complex64, error) {
	c.mu.Lock()
	defer c.mu.Unlock()
	out, err := c.ctx.PhaseOffsets(bufs.ptr, nChan, n)
	return out, translate(err)
}

// DeviceInfo describes one GPU the backend can run on.
type DeviceInfo struct {
	Index            int
	Name             string
	SMMajor, SMMinor int
	SMCount          int
	HBMBytes         uint64
}

// Info is what debug.ReadBuildInfo reports for this backend (debug/build.go:60-75; twin:
// debug/build_cuda.go): the library version and every device it accepts.  Err is set -- and Devices is
// empty -- when there is no usable GPU; nothing falls back to the CPU.
type Info struct {
	Library string
	Devices []DeviceInfo
	Err     error
}

// ReadInfo probes the devices (hzsdr_device_count, hzsdr_ctx_create, hzsdr_ctx_info).
func ReadInfo() Info {
	info := Info{Library: hzcuda.Version()}
	n, err := hzcuda.DeviceCount()
	if err != nil {
		info.Err = translate(err)
		return info
	}
	for d := 0; d < n; d++ {
		ctx, err := hzcuda.NewCtx(d)
		if err != nil { // e.g. not an sm_100 part
			if info.Err == nil {
				info.Err = translate(err)
			}
			continue
		}
		if di, err := ctx.Info(); err == nil {
			info.Devices = append(info.Devices, DeviceInfo{Index: d, Name: di.Name, SMMajor: di.SMMajor, SMMinor: di.SMMinor,
				SMCount: di.SMCount, HBMBytes: di.HBMBytes})
		}
		ctx.Close()
	}
	if len(info.Devices) > 0 {
		info.Err = nil
	}
	return info
}

// BuildInfo is ReadInfo as one line of text.
func BuildInfo() string {
	info := ReadInfo()
	if len(info.Devices) == 0 {
		return fmt.Sprintf("cuda: %s, unavailable (%v)", info.Library, info.Err)
	}
	d := info.Devices[0]
	return fmt.Sprintf("cuda: %s, %d device(s), device %d: %s sm_%d%d, %d SMs, %d GiB", info.Library, len(info.Devices), d.Index, d.Name,
		d.SMMajor, d.SMMinor, d.SMCount, d.HBMBytes>>30)
}
