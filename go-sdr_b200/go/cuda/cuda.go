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

// BuildInfo is what debug.ReadBuildInfo reports for this backend (debug/build.go:60-75).
func BuildInfo() string {
	n, err := hzcuda.DeviceCount()
	if err != nil {
		return fmt.Sprintf("cuda: unavailable (%v)", err)
	}
	return fmt.Sprintf("cuda: libhzsdrcuda sm_100a, %d device(s)", n)
}
