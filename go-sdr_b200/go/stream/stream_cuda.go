// {{{ Copyright (c) the hzsdr-cuda authors, MIT (same terms as hz.tools/sdr) }}}

//go:build sdr.cuda

// stream_cuda.go -- the `sdr.cuda` twins of stream/convert.go, shifter.go, convolution.go,
// decimate.go, downsample.go, multiply.go, gain.go, add.go and beamform.go.  Each of those files
// gains `//go:build !sdr.cuda`; this file provides the same exported constructors with the same
// signatures, returning GPU-backed readers.  There is no CPU fallback: without a B200 every
// constructor returns the context error.
//
// Design (mirrors the compiled + tested C++ twin, go-sdr_b200/host/hzsdr.hpp):
//   - deviceReader: a Reader that can deliver its output into device memory.  Stages discover a
//     GPU upstream by type assertion (the trick SetPhaseAngles already uses, beamform.go:136) and
//     then never touch the host between stages.
//   - gpuReadTransformer: ReadTransformer's block rules (read_transformer.go:118-137) -- ReadFull a
//     block, Proc, hand out, drop a trailing partial block and latch the error -- with Proc being
//     one kernel launch over a batch of blocks.
//   - DecimateReader(ConvolutionReader(ShiftReader(ConvertReader(raw)))) collapses into one fused
//     kernel per buffer (hzsdr_chain_*) when no intermediate stage has been read from yet.
package stream

import (
	"fmt"
	"io"
	"unsafe"

	"hz.tools/rf"
	"hz.tools/sdr"
	"hz.tools/sdr/cuda"
	"hz.tools/sdr/fft"
	"hz.tools/sdr/internal/hzcuda"
)

const gpuBlock = 32 * 1024 // convert.go:43-44, decimate.go:41-42, downsample.go:54-55

// deviceReader is implemented by every reader in this file.
type deviceReader interface {
	sdr.Reader
	// readDevice fills up to dst.Length() samples of device memory in this reader's format.
	readDevice(dst *devBuf) (int, error)
	context() *cuda.Context
}

// devBuf is a typed window of device memory (raw formats too, unlike cuda.SamplesC64).
type devBuf struct {
	ptr    unsafe.Pointer
	n      int
	format sdr.SampleFormat
}

func (d *devBuf) slice(a, b int) *devBuf {
	return &devBuf{ptr: unsafe.Add(d.ptr, a*d.format.Size()), n: b - a, format: d.format}
}

func newDevBuf(c *cuda.Context, f sdr.SampleFormat, n int) (*devBuf, error) {
	p, err := c.Raw().Alloc(n * f.Size())
	if err != nil {
		return nil, cuda.Translate(err)
	}
	return &devBuf{ptr: p, n: n, format: f}, nil
}

// readFullToDevice is sdr.ReadFull with a device destination: GPU upstreams write in place, host
// upstreams are read into a pinned staging buffer and uploaded (the one H2D of the chain).
func readFullToDevice(c *cuda.Context, in sdr.Reader, dst *devBuf, stage *sdr.Samples) (int, error) {
	if dr, ok := in.(deviceReader); ok {
		n := 0
		for n < dst.n {
			nn, err := dr.readDevice(dst.slice(n, dst.n))
			n += nn
			if err != nil {
				if n > 0 && err == io.EOF {
					return n, sdr.ErrUnexpectedEOF // reader.go:109-111
				}
				return n, err
			}
		}
		return n, nil
	}
	if *stage == nil || (*stage).Length() < dst.n {
		s, _, err := cuda.PinnedSamples(in.SampleFormat(), dst.n)
		if err != nil {
			return 0, err
		}
		*stage = s
	}
	view := (*stage).Slice(0, dst.n)
	n, err := sdr.ReadFull(in, view)
	if n > 0 {
		b, _ := sdr.UnsafeSamplesAsBytes(view.Slice(0, n))
		if uerr := c.Raw().Upload(dst.ptr, unsafe.Pointer(&b[0]), len(b)); uerr != nil { // pinned: async is legal
			return 0, cuda.Translate(uerr)
		}
		if uerr := c.Raw().Sync(); uerr != nil { // the staging buffer is reused by the next read
			return 0, cuda.Translate(uerr)
		}
	}
	return n, err
}

// hostRead implements sdr.Reader.Read on top of readDevice: a device destination stays on the
// device, a host destination gets one D2H.
func hostRead(r deviceReader, s sdr.Samples, wrongFormat error, stage **devBuf) (int, error) {
	if s.Format() != r.SampleFormat() {
		return 0, wrongFormat // testutils/reader.go:87-97
	}
	if ds, ok := s.(interface {
		DevicePointer() (unsafe.Pointer, *hzcuda.Ctx)
	}); ok {
		p, _ := ds.DevicePointer()
		return r.readDevice(&devBuf{ptr: p, n: s.Length(), format: s.Format()})
	}
	if *stage == nil || (*stage).n < s.Length() {
		b, err := newDevBuf(r.context(), s.Format(), s.Length())
		if err != nil {
			return 0, err
		}
		*stage = b
	}
	n, err := r.readDevice((*stage).slice(0, s.Length()))
	if n > 0 {
		b, _ := sdr.UnsafeSamplesAsBytes(s.Slice(0, n))
		if derr := r.context().Raw().Download(b, (*stage).ptr); derr != nil {
			return 0, cuda.Translate(derr)
		}
	}
	return n, err
}

// ---- ReadTransformer on the GPU ------------------------------------------------------------------

type gpuProc func(in *devBuf, blocks int, out *devBuf) (int, error)

type gpuReadTransformer struct {
	ctx                    *cuda.Context
	in                     sdr.Reader
	inLen, outLen, batch   int
	format                 sdr.SampleFormat
	rate                   uint
	proc                   gpuProc
	inBuf, outBuf, hostOut *devBuf
	stage                  sdr.Samples
	pos, avail             int
	err                    error
}

func (t *gpuReadTransformer) SampleFormat() sdr.SampleFormat { return t.format }
func (t *gpuReadTransformer) SampleRate() uint               { return t.rate }
func (t *gpuReadTransformer) context() *cuda.Context         { return t.ctx }
func (t *gpuReadTransformer) Read(s sdr.Samples) (int, error) {
	return hostRead(t, s, sdr.ErrSampleFormatMismatch, &t.hostOut)
}

func (t *gpuReadTransformer) readDevice(dst *devBuf) (int, error) {
	if t.avail == 0 {
		if t.err != nil {
			return 0, t.err // latched, read_transformer.go:121-135
		}
		if t.inBuf == nil {
			var err error
			if t.inBuf, err = newDevBuf(t.ctx, t.in.SampleFormat(), t.inLen*t.batch); err != nil {
				return 0, err
			}
			if t.outBuf, err = newDevBuf(t.ctx, t.format, t.outLen*t.batch); err != nil {
				return 0, err
			}
		}
		n, err := readFullToDevice(t.ctx, t.in, t.inBuf, &t.stage)
		if err != nil {
			t.err = err
		}
		blocks := n / t.inLen // the partial block is dropped
		t.pos = 0
		if blocks > 0 {
			got, perr := t.proc(t.inBuf.slice(0, blocks*t.inLen), blocks, t.outBuf)
			if perr != nil {
				t.err = perr
				return 0, perr
			}
			t.avail = got
		}
		if t.avail == 0 {
			return 0, t.err
		}
	}
	n := dst.n
	if t.avail < n {
		n = t.avail
	}
	if err := t.ctx.Raw().Copy(dst.ptr, t.outBuf.slice(t.pos, t.pos+n).ptr, n*t.format.Size()); err != nil {
		return 0, cuda.Translate(err)
	}
	t.pos += n
	t.avail -= n
	return n, nil
}

func ctxFor(in sdr.Reader) (*cuda.Context, error) {
	if dr, ok := in.(deviceReader); ok {
		return dr.context(), nil
	}
	return cuda.Default()
}

// ConvertReader: stream/convert.go:37-51.
func ConvertReader(in sdr.Reader, to sdr.SampleFormat) (sdr.Reader, error) {
	if to != sdr.SampleFormatC64 {
		return nil, sdr.ErrConversionNotImplemented // reverse conversions: SURVEY.md 8(f) rank 3
	}
	c, err := ctxFor(in)
	if err != nil {
		return nil, err
	}
	from := in.SampleFormat()
	return &gpuReadTransformer{ctx: c, in: in, inLen: gpuBlock, outLen: gpuBlock, batch: 16, format: to, rate: in.SampleRate(),
		proc: func(i *devBuf, blocks int, o *devBuf) (int, error) {
			n, err := c.Raw().ConvertToC64(int(from), i.ptr, blocks*gpuBlock, o.ptr, o.n)
			return n, cuda.Translate(err)
		}}, nil
}

// ConvertWriter: stream/convert.go:58-118.  The conversion is sdr.ConvertBuffer -- under this build
// tag the GPU one in conv_cuda.go, every pair of formats -- 32 Ki samples at a time into a buffer of
// out's format, each piece passed on to out.Write.
func ConvertWriter(out sdr.Writer, inputFormat sdr.SampleFormat) (sdr.Writer, error) {
	buf, err := sdr.MakeSamples(out.SampleFormat(), gpuBlock)
	if err != nil {
		return nil, err
	}
	return &convWriter{out: out, inputFormat: inputFormat, buffer: buf}, nil
}

type convWriter struct {
	out         sdr.Writer
	inputFormat sdr.SampleFormat
	buffer      sdr.Samples
}

func (cw *convWriter) SampleFormat() sdr.SampleFormat { return cw.inputFormat }
func (cw *convWriter) SampleRate() uint               { return cw.out.SampleRate() }
func (cw *convWriter) Write(in sdr.Samples) (int, error) {
	if in.Format() != cw.inputFormat {
		return 0, sdr.ErrSampleFormatMismatch
	}
	size, n := cw.buffer.Length(), 0
	for i := 0; i < in.Length(); i += size {
		ie := i + size
		if ie > in.Length() {
			ie = in.Length()
		}
		got, err := sdr.ConvertBuffer(cw.buffer, in.Slice(i, ie))
		if err != nil {
			return n, err
		}
		if got != ie-i {
			return n, fmt.Errorf("ConvertWriter: Conversion mismatch")
		}
		j, err := cw.out.Write(cw.buffer.Slice(0, got))
		n += j
		if err != nil {
			return n, err
		}
	}
	return n, nil
}

// ---- ShiftReader: stream/shifter.go:44-102 -----------------------------------------------------

type shiftReader struct {
	ctx     *cuda.Context
	r       sdr.Reader
	shift   rf.Hz
	nco     hzcuda.Nco
	stage   sdr.Samples
	hostOut *devBuf
}

func (sr *shiftReader) SampleFormat() sdr.SampleFormat { return sr.r.SampleFormat() }
func (sr *shiftReader) SampleRate() uint               { return sr.r.SampleRate() }
func (sr *shiftReader) context() *cuda.Context         { return sr.ctx }
func (sr *shiftReader) Read(s sdr.Samples) (int, error) {
	return hostRead(sr, s, sdr.ErrSampleFormatUnknown, &sr.hostOut) // shifter.go:45-50
}
func (sr *shiftReader) readDevice(dst *devBuf) (int, error) {
	var n int
	var err error
	if dr, ok := sr.r.(deviceReader); ok {
		n, err = dr.readDevice(dst)
	} else {
		// one Read of the upstream, whatever it returns (shifter.go:52-55)
		if sr.stage == nil || sr.stage.Length() < dst.n {
			if sr.stage, _, err = cuda.PinnedSamples(sdr.SampleFormatC64, dst.n); err != nil {
				return 0, err
			}
		}
		view := sr.stage.Slice(0, dst.n)
		n, err = sr.r.Read(view)
		if n > 0 {
			b, _ := sdr.UnsafeSamplesAsBytes(view.Slice(0, n))
			if uerr := sr.ctx.Raw().Upload(dst.ptr, unsafe.Pointer(&b[0]), len(b)); uerr != nil {
				return 0, cuda.Translate(uerr)
			}
			if uerr := sr.ctx.Raw().Sync(); uerr != nil {
				return 0, cuda.Translate(uerr)
			}
		}
	}
	if err != nil {
		return n, err
	}
	// the fp64 time accumulator `ts` continues across reads, bit-equal to shifter.go:73-79
	return n, cuda.Translate(sr.ctx.Raw().Shift(dst.ptr, n, float64(sr.shift), &sr.nco))
}

// ShiftReader will shift the iq samples by the target frequency (stream/shifter.go:89-102).
func ShiftReader(r sdr.Reader, shift rf.Hz) (sdr.Reader, error) {
	if r.SampleFormat() != sdr.SampleFormatC64 {
		return nil, sdr.ErrSampleFormatUnknown
	}
	c, err := ctxFor(r)
	if err != nil {
		return nil, err
	}
	return &shiftReader{ctx: c, r: r, shift: shift, nco: hzcuda.Nco{SampleRate: uint32(r.SampleRate())}}, nil
}

// ShiftBuffer: stream/shifter.go:66-85 -- the buffer-level form ShiftReader is built from.  The closure
// owns the carried fp64 time accumulator exactly like the reference's (bit-equal `ts`, wrapped at 2*pi
// seconds); a host SamplesC64 is staged through the GPU (H2D, hzsdr_shift, D2H).  The reference's
// closure cannot report errors, and neither can this one: like the SIMD gate
// (internal/simd/enabled_amd64.go:35-50) it panics when there is no GPU.
func ShiftBuffer(sampleRate uint) func(rf.Hz, sdr.SamplesC64) {
	nco := hzcuda.Nco{SampleRate: uint32(sampleRate)}
	return func(shift rf.Hz, buf sdr.SamplesC64) {
		if len(buf) == 0 {
			return
		}
		c, err := cuda.Default()
		if err != nil {
			panic(err)
		}
		if err := onDevice(c, buf, func(d *devBuf) error {
			return c.Raw().Shift(d.ptr, d.n, float64(shift), &nco)
		}); err != nil {
			panic(err)
		}
	}
}

// onDevice runs fn over a device copy of a host buffer and copies the result back over it.
func onDevice(c *cuda.Context, buf sdr.Samples, fn func(*devBuf) error) error {
	d, err := newDevBuf(c, buf.Format(), buf.Length())
	if err != nil {
		return err
	}
	defer c.Raw().Free(d.ptr)
	b, err := sdr.UnsafeSamplesAsBytes(buf)
	if err != nil {
		return err
	}
	if err := c.Raw().UploadGo(d.ptr, b); err != nil {
		return cuda.Translate(err)
	}
	if err := fn(d); err != nil {
		return cuda.Translate(err)
	}
	return cuda.Translate(c.Raw().Download(b, d.ptr))
}

// devicePointerOf: the device window behind s, when s is device-resident (cuda.SamplesC64).
func devicePointerOf(s sdr.Samples) (unsafe.Pointer, bool) {
	if ds, ok := s.(interface {
		DevicePointer() (unsafe.Pointer, *hzcuda.Ctx)
	}); ok {
		p, _ := ds.DevicePointer()
		return p, true
	}
	return nil, false
}

// bufferOp is the shared body of DecimateBuffer / DownsampleBuffer: `from` and `to` may each be host or
// device samples; host sides are staged, device sides are used in place.
func bufferOp(to, from sdr.Samples, want int, run func(c *cuda.Context, src, dst unsafe.Pointer) (int, error)) (int, error) {
	c, err := cuda.Default()
	if err != nil {
		return 0, err // no CPU fallback
	}
	src, srcDev := devicePointerOf(from)
	if !srcDev {
		d, err := newDevBuf(c, from.Format(), from.Length())
		if err != nil {
			return 0, err
		}
		defer c.Raw().Free(d.ptr)
		b, err := sdr.UnsafeSamplesAsBytes(from)
		if err != nil {
			return 0, err
		}
		if err := c.Raw().UploadGo(d.ptr, b); err != nil {
			return 0, cuda.Translate(err)
		}
		src = d.ptr
	}
	dst, dstDev := devicePointerOf(to)
	var stage *devBuf
	if !dstDev {
		if stage, err = newDevBuf(c, to.Format(), want); err != nil {
			return 0, err
		}
		defer c.Raw().Free(stage.ptr)
		dst = stage.ptr
	}
	n, err := run(c, src, dst)
	if err != nil {
		return 0, cuda.Translate(err)
	}
	if !dstDev && n > 0 {
		b, err := sdr.UnsafeSamplesAsBytes(to.Slice(0, n))
		if err != nil {
			return 0, err
		}
		if err := c.Raw().Download(b, stage.ptr); err != nil {
			return 0, cuda.Translate(err)
		}
	}
	return n, nil
}

// DecimateBuffer: stream/decimate.go:59-101 -- to[i] = from[factor*i] for i < len(from)/factor.  Formats
// U8, I16 and C64 (I8 is ErrSampleFormatUnknown, as in the reference :85-97); `offset` is accepted and
// ignored exactly as the reference ignores it.
func DecimateBuffer(to, from sdr.Samples, factor uint, offset int) (int, error) {
	if from.Format() != to.Format() {
		return 0, sdr.ErrSampleFormatMismatch
	}
	want := from.Length() / int(factor)
	if to.Length() < want {
		return 0, sdr.ErrDstTooSmall
	}
	switch from.Format() {
	case sdr.SampleFormatU8, sdr.SampleFormatI16, sdr.SampleFormatC64:
	default:
		return 0, sdr.ErrSampleFormatUnknown
	}
	if want == 0 {
		return 0, nil
	}
	return bufferOp(to, from, want, func(c *cuda.Context, src, dst unsafe.Pointer) (int, error) {
		return c.Raw().Decimate(int(from.Format()), src, from.Length(), dst, want, factor, 0)
	})
}

// DownsampleBuffer: stream/downsample.go:68-127 -- to[i] = mean of from[i*factor : (i+1)*factor], summed
// sequentially in fp32; `to` must be C64, `from` U8, I16 or C64.
func DownsampleBuffer(to, from sdr.Samples, factor uint, offset int) (int, error) {
	if to.Format() != sdr.SampleFormatC64 {
		return 0, sdr.ErrSampleFormatMismatch
	}
	want := from.Length() / int(factor)
	if to.Length() < want {
		return 0, sdr.ErrDstTooSmall
	}
	switch from.Format() {
	case sdr.SampleFormatU8, sdr.SampleFormatI16, sdr.SampleFormatC64:
	default:
		return 0, sdr.ErrSampleFormatUnknown
	}
	if want == 0 {
		return 0, nil
	}
	return bufferOp(to, from, want, func(c *cuda.Context, src, dst unsafe.Pointer) (int, error) {
		return c.Raw().Downsample(int(from.Format()), src, from.Length(), dst, want, factor, 0)
	})
}

// ---- ConvolutionReader: stream/convolution.go:36-82 --------------------------------------------

// convolutionReader remembers its parts so DecimateReader can fuse the whole chain.
type convolutionReader struct {
	*gpuReadTransformer
	filter []complex64
	src    sdr.Reader
}

// ConvolutionReader keeps the reference's signature.  The planner argument is accepted for source
// compatibility; the transform pair is fused into the convolution kernel (no Planner round trip),
// with the convention documented in DESIGN.md (forward e^{-2 pi i kn/N}, both unnormalised).
func ConvolutionReader(r sdr.Reader, planner fft.Planner, filter []complex64) (sdr.Reader, error) {
	if r.SampleFormat() != sdr.SampleFormatC64 {
		return nil, sdr.ErrSampleFormatUnknown
	}
	c, err := ctxFor(r)
	if err != nil {
		return nil, err
	}
	n := len(filter)
	fdev, err := newDevBuf(c, sdr.SampleFormatC64, n)
	if err != nil {
		return nil, err
	}
	fb := unsafe.Slice((*byte)(unsafe.Pointer(&filter[0])), n*8)
	if err := c.Raw().UploadGo(fdev.ptr, fb); err != nil {
		return nil, cuda.Translate(err)
	}
	if err := c.Raw().ConvolveFreq(nil, nil, fdev.ptr, n, 0); err != nil { // unsupported length fails here
		return nil, cuda.Translate(err)
	}
	batch := (1 << 19) / n
	if batch < 1 {
		batch = 1
	}
	t := &gpuReadTransformer{ctx: c, in: r, inLen: n, outLen: n, batch: batch, format: sdr.SampleFormatC64, rate: r.SampleRate(),
		proc: func(i *devBuf, blocks int, o *devBuf) (int, error) {
			return blocks * n, cuda.Translate(c.Raw().ConvolveFreq(i.ptr, o.ptr, fdev.ptr, n, blocks))
		}}
	return &convolutionReader{gpuReadTransformer: t, filter: filter, src: r}, nil
}

// ---- DecimateReader / DownsampleReader -----------------------------------------------------------

// DecimateReader: stream/decimate.go:34-51.  When `in` is an untouched
// ConvolutionReader(ShiftReader(ConvertReader(raw))) the four stages become one fused kernel.
func DecimateReader(in sdr.Reader, factor uint) (sdr.Reader, error) {
	if fused := tryFuseChain(in, factor); fused != nil {
		return fused, nil
	}
	c, err := ctxFor(in)
	if err != nil {
		return nil, err
	}
	f := in.SampleFormat()
	return &gpuReadTransformer{ctx: c, in: in, inLen: gpuBlock, outLen: gpuBlock, batch: 16, format: f, rate: in.SampleRate() / factor,
		proc: func(i *devBuf, blocks int, o *devBuf) (int, error) {
			n, err := c.Raw().Decimate(int(f), i.ptr, blocks*gpuBlock, o.ptr, o.n, factor, gpuBlock)
			return n, cuda.Translate(err)
		}}, nil
}

// DownsampleReader: stream/downsample.go:47-64.
func DownsampleReader(in sdr.Reader, factor uint) (sdr.Reader, error) {
	c, err := ctxFor(in)
	if err != nil {
		return nil, err
	}
	f := in.SampleFormat()
	return &gpuReadTransformer{ctx: c, in: in, inLen: gpuBlock, outLen: gpuBlock, batch: 16, format: sdr.SampleFormatC64,
		rate: in.SampleRate() / factor,
		proc: func(i *devBuf, blocks int, o *devBuf) (int, error) {
			n, err := c.Raw().Downsample(int(f), i.ptr, blocks*gpuBlock, o.ptr, o.n, factor, gpuBlock)
			return n, cuda.Translate(err)
		}}, nil
}

// ---- fused chain -------------------------------------------------------------------------------

type chainReader struct {
	ctx        *cuda.Context
	raw        sdr.Reader
	chain      *hzcuda.Chain
	factor     uint
	unit       int
	in, out    *devBuf
	hostOut    *devBuf
	stage      sdr.Samples
	pos, avail int
	err        error
}

func tryFuseChain(in sdr.Reader, factor uint) sdr.Reader {
	cv, ok := in.(*convolutionReader)
	if !ok || cv.inBuf != nil { // already read from: cannot re-anchor the block boundaries
		return nil
	}
	sh, ok := cv.src.(*shiftReader)
	if !ok || sh.nco.Ts != 0 {
		return nil
	}
	ct, ok := sh.r.(*gpuReadTransformer)
	if !ok || ct.inBuf != nil || ct.inLen != gpuBlock || ct.in.SampleFormat() == sdr.SampleFormatC64 {
		return nil
	}
	raw := ct.in
	ch, err := cv.ctx.Raw().NewChain(hzcuda.ChainConfig{SrcFormat: int(raw.SampleFormat()), SampleRate: uint32(raw.SampleRate()),
		ShiftHz: float64(sh.shift), Filter: cv.filter, Decimate: uint32(factor)})
	if err != nil {
		return nil // e.g. unsupported length: fall back to the unfused GPU stages
	}
	unit := gpuBlock
	if len(cv.filter) > unit {
		unit = len(cv.filter)
	}
	return &chainReader{ctx: cv.ctx, raw: raw, chain: ch, factor: factor, unit: unit}
}

func (c *chainReader) SampleFormat() sdr.SampleFormat { return sdr.SampleFormatC64 }
func (c *chainReader) SampleRate() uint               { return c.raw.SampleRate() / c.factor }
func (c *chainReader) context() *cuda.Context         { return c.ctx }
func (c *chainReader) Read(s sdr.Samples) (int, error) {
	return hostRead(c, s, sdr.ErrSampleFormatMismatch, &c.hostOut)
}
func (c *chainReader) readDevice(dst *devBuf) (int, error) {
	const blocksPerLaunch = 128
	if c.avail == 0 {
		if c.err != nil {
			return 0, c.err
		}
		if c.in == nil {
			var err error
			if c.in, err = newDevBuf(c.ctx, c.raw.SampleFormat(), c.unit*blocksPerLaunch); err != nil {
				return 0, err
			}
			if c.out, err = newDevBuf(c.ctx, sdr.SampleFormatC64, c.chain.OutLen(c.unit*blocksPerLaunch)); err != nil {
				return 0, err
			}
		}
		n, err := readFullToDevice(c.ctx, c.raw, c.in, &c.stage)
		if err != nil {
			c.err = err
		}
		n = (n / c.unit) * c.unit // whole blocks only: the composition of the four ReadTransformers
		c.pos = 0
		if n > 0 {
			got, xerr := c.chain.Exec(c.in.ptr, n, c.out.ptr, c.out.n)
			if xerr != nil {
				c.err = cuda.Translate(xerr)
				return 0, c.err
			}
			c.avail = got
		}
		if c.avail == 0 {
			return 0, c.err
		}
	}
	n := dst.n
	if c.avail < n {
		n = c.avail
	}
	if err := c.ctx.Raw().Copy(dst.ptr, c.out.slice(c.pos, c.pos+n).ptr, n*8); err != nil {
		return 0, cuda.Translate(err)
	}
	c.pos += n
	c.avail -= n
	return n, nil
}

// ---- Multiply / Gain / Add ---------------------------------------------------------------------

type multiplyReader struct {
	ctx     *cuda.Context
	r       sdr.Reader
	m       complex64
	isGain  bool
	stage   sdr.Samples
	hostOut *devBuf
}

// SetMultiplier is the reference's undocumented API (multiply.go:34-36); Beamform relies on it.
func (mr *multiplyReader) SetMultiplier(m complex64)      { mr.m = m }
func (mr *multiplyReader) SampleFormat() sdr.SampleFormat { return mr.r.SampleFormat() }
func (mr *multiplyReader) SampleRate() uint               { return mr.r.SampleRate() }
func (mr *multiplyReader) context() *cuda.Context         { return mr.ctx }
func (mr *multiplyReader) Read(s sdr.Samples) (int, error) {
	return hostRead(mr, s, sdr.ErrSampleFormatMismatch, &mr.hostOut) // multiply.go:47-52
}
func (mr *multiplyReader) readDevice(dst *devBuf) (int, error) {
	n, err := readFullToDevice(mr.ctx, mr.r, dst, &mr.stage) // GPU upstreams: in place
	if err != nil && n == 0 {
		return 0, err
	}
	switch {
	case mr.isGain:
		err = cuda.Translate(mr.ctx.Raw().Scale(dst.ptr, n, real(mr.m))) // gain.go:39-57
	case mr.m != 1: // multiply.go:59-62
		err = cuda.Translate(mr.ctx.Raw().Rotate(dst.ptr, n, mr.m))
	}
	return n, err
}

// Multiply: stream/multiply.go:74-89.
func Multiply(r sdr.Reader, m complex64) (sdr.Reader, error) {
	c, err := ctxFor(r)
	if err != nil {
		return nil, err
	}
	switch r.SampleFormat() {
	case sdr.SampleFormatC64:
		return &multiplyReader{ctx: c, r: r, m: m}, nil
	case sdr.SampleFormatU8, sdr.SampleFormatI8:
		ret := &lutMultiplyReader{ctx: c, r: r}
		if err := ret.SetMultiplier(m); err != nil {
			return nil, err
		}
		return ret, nil
	default:
		return nil, sdr.ErrSampleFormatUnknown
	}
}

// lutMultiplyReader is uint8MultiplyReader / int8MultiplyReader (multiply.go:91-251): every Read is
// a 65536-entry table lookup on the GPU; SetMultiplier rebuilds the table by Convert -> Multiply ->
// Convert exactly as the reference does (the u8 reader's x0*255+x1 index, collisions included, is
// reproduced by re-indexing its 65535-entry table into the library's little-endian pair index).
// The C++ twin (host/hzsdr.hpp LutMultiplyReaderGpu) is the compiled + tested version of this code.
type lutMultiplyReader struct {
	ctx     *cuda.Context
	r       sdr.Reader
	table   *devBuf
	in      *devBuf
	stage   sdr.Samples
	hostOut *devBuf
}

func (mr *lutMultiplyReader) SampleFormat() sdr.SampleFormat { return mr.r.SampleFormat() }
func (mr *lutMultiplyReader) SampleRate() uint               { return mr.r.SampleRate() }
func (mr *lutMultiplyReader) context() *cuda.Context         { return mr.ctx }
func (mr *lutMultiplyReader) Read(s sdr.Samples) (int, error) {
	return hostRead(mr, s, sdr.ErrSampleFormatMismatch, &mr.hostOut)
}

func (mr *lutMultiplyReader) SetMultiplier(m complex64) error {
	f := mr.SampleFormat()
	raw := mr.ctx.Raw()
	entries := 65536
	ident := make([]byte, 2*65536)
	if f == sdr.SampleFormatU8 {
		entries = 65535
		for re := 0; re < 256; re++ { // multiply.go:153-160, later writes win
			for im := 0; im <= 256; im++ {
				i := re*255 + (im & 0xff)
				ident[2*i], ident[2*i+1] = byte(re), byte(im)
			}
		}
	} else {
		for i := 0; i < 65536; i++ { // LookupTableIdentityI8, iq_lookup_table.go:82-90
			ident[2*i], ident[2*i+1] = byte(i), byte(i>>8)
		}
	}
	rawBuf, err := newDevBuf(mr.ctx, f, entries)
	if err != nil {
		return err
	}
	defer raw.Free(rawBuf.ptr)
	c64, err := newDevBuf(mr.ctx, sdr.SampleFormatC64, entries)
	if err != nil {
		return err
	}
	defer raw.Free(c64.ptr)
	if err := raw.UploadGo(rawBuf.ptr, ident[:2*entries]); err != nil {
		return cuda.Translate(err)
	}
	if _, err := raw.Convert(int(f), rawBuf.ptr, entries, int(sdr.SampleFormatC64), c64.ptr, entries); err != nil {
		return cuda.Translate(err)
	}
	if err := raw.Rotate(c64.ptr, entries, m); err != nil { // cbuf.Multiply(m)
		return cuda.Translate(err)
	}
	if _, err := raw.Convert(int(sdr.SampleFormatC64), c64.ptr, entries, int(f), rawBuf.ptr, entries); err != nil {
		return cuda.Translate(err)
	}
	tab := make([]byte, 2*entries)
	if err := raw.Download(tab, rawBuf.ptr); err != nil {
		return cuda.Translate(err)
	}
	full := make([]byte, 2*65536)
	for i := 0; i < 65536; i++ {
		j := i
		if f == sdr.SampleFormatU8 {
			j = (i&0xff)*255 + (i >> 8)
		}
		full[2*i], full[2*i+1] = tab[2*j], tab[2*j+1]
	}
	if mr.table == nil {
		if mr.table, err = newDevBuf(mr.ctx, f, 65536); err != nil {
			return err
		}
	}
	return cuda.Translate(raw.UploadGo(mr.table.ptr, full))
}

func (mr *lutMultiplyReader) readDevice(dst *devBuf) (int, error) {
	if mr.in == nil || mr.in.n < dst.n {
		b, err := newDevBuf(mr.ctx, mr.SampleFormat(), dst.n)
		if err != nil {
			return 0, err
		}
		mr.in = b
	}
	var n int
	var err error
	if dr, ok := mr.r.(deviceReader); ok {
		n, err = dr.readDevice(mr.in.slice(0, dst.n))
	} else {
		if mr.stage == nil || mr.stage.Length() < dst.n {
			if mr.stage, _, err = cuda.PinnedSamples(mr.SampleFormat(), dst.n); err != nil {
				return 0, err
			}
		}
		view := mr.stage.Slice(0, dst.n)
		n, err = mr.r.Read(view) // one Read of the upstream, multiply.go:120-123
		if n > 0 {
			b, _ := sdr.UnsafeSamplesAsBytes(view.Slice(0, n))
			if uerr := mr.ctx.Raw().Upload(mr.in.ptr, unsafe.Pointer(&b[0]), len(b)); uerr != nil {
				return 0, cuda.Translate(uerr)
			}
			if uerr := mr.ctx.Raw().Sync(); uerr != nil {
				return 0, cuda.Translate(uerr)
			}
		}
	}
	if err != nil {
		return n, err
	}
	f := int(mr.SampleFormat())
	return n, cuda.Translate(mr.ctx.Raw().Lookup(f, mr.in.ptr, n, f, mr.table.ptr, dst.ptr, dst.n))
}

// Gain: stream/gain.go:30-57.
func Gain(r sdr.Reader, v float32) sdr.Reader {
	c, _ := ctxFor(r)
	return &multiplyReader{ctx: c, r: r, m: complex(v, 0), isGain: true}
}

type addReader struct {
	ctx     *cuda.Context
	readers []sdr.Reader
	bufs    []*devBuf
	stage   sdr.Samples
	hostOut *devBuf
	err     error
}

func (ar *addReader) SampleFormat() sdr.SampleFormat { return ar.readers[0].SampleFormat() }
func (ar *addReader) SampleRate() uint               { return ar.readers[0].SampleRate() }
func (ar *addReader) context() *cuda.Context         { return ar.ctx }
func (ar *addReader) Read(s sdr.Samples) (int, error) {
	return hostRead(ar, s, sdr.ErrSampleFormatUnknown, &ar.hostOut) // add.go:129-135
}
func (ar *addReader) readDevice(dst *devBuf) (int, error) {
	if ar.err != nil {
		return 0, ar.err // add.go:125-127
	}
	ptrs := make([]unsafe.Pointer, len(ar.readers))
	for i, r := range ar.readers {
		if i >= len(ar.bufs) || ar.bufs[i].n < dst.n {
			b, err := newDevBuf(ar.ctx, ar.SampleFormat(), dst.n)
			if err != nil {
				return 0, err
			}
			if i >= len(ar.bufs) {
				ar.bufs = append(ar.bufs, b)
			} else {
				ar.bufs[i] = b
			}
		}
		if _, err := readFullToDevice(ar.ctx, r, ar.bufs[i].slice(0, dst.n), &ar.stage); err != nil {
			ar.err = err // add.go:148-158
			return 0, err
		}
		ptrs[i] = ar.bufs[i].ptr
	}
	if f := ar.SampleFormat(); f != sdr.SampleFormatC64 { // wrapping integer adds, add.go:95-113
		return dst.n, cuda.Translate(ar.ctx.Raw().AddInt(int(f), dst.ptr, ptrs, dst.n))
	}
	// out = ((0 + b0) + b1) + ... in reader order, fp32 (add.go:115-119,165-168)
	return dst.n, cuda.Translate(ar.ctx.Raw().Add(dst.ptr, ptrs, dst.n))
}

// Add: stream/add.go:41-78 (complex64, int16, int8 -- the reference's three cases).
func Add(readers ...sdr.Reader) (sdr.Reader, error) {
	switch len(readers) {
	case 0:
		return nil, fmt.Errorf("stream.Add: No readers passed")
	case 1:
		return readers[0], nil
	}
	switch readers[0].SampleFormat() {
	case sdr.SampleFormatC64, sdr.SampleFormatI16, sdr.SampleFormatI8:
	default:
		return nil, sdr.ErrSampleFormatUnknown // add.go:56-61
	}
	for _, r := range readers {
		if r.SampleFormat() != readers[0].SampleFormat() {
			return nil, fmt.Errorf("stream.Add: Readers are not all the same format")
		}
		if r.SampleRate() != readers[0].SampleRate() {
			return nil, fmt.Errorf("stream.Add: Readers are not all the same rate")
		}
	}
	c, err := ctxFor(readers[0])
	if err != nil {
		return nil, err
	}
	return &addReader{ctx: c, readers: readers}, nil
}

// ---- Beamform: stream/beamform.go ----------------------------------------------------------------

// Beamform combines coherent raw readers into one complex64 beam.  Where the reference stacks N
// ConvertReaders, N Multiply readers and an Add (beamform.go:148-171), this is one kernel per block
// reading every raw channel once (hzsdr_beamform).
type Beamform struct {
	sdr.Reader
	impl *beamReader
}

type beamReader struct {
	ctx     *cuda.Context
	readers sdr.Readers
	w       []complex64
	raw     []*devBuf
	stage   sdr.Samples
	hostOut *devBuf
	err     error
}

// BeamformConfig contains configuration for the combined samples (beamform.go:142-145).
type BeamformConfig struct {
	Angles []complex64
}

func (b *beamReader) SampleFormat() sdr.SampleFormat { return sdr.SampleFormatC64 }
func (b *beamReader) SampleRate() uint               { return b.readers[0].SampleRate() }
func (b *beamReader) context() *cuda.Context         { return b.ctx }
func (b *beamReader) Read(s sdr.Samples) (int, error) {
	return hostRead(b, s, sdr.ErrSampleFormatUnknown, &b.hostOut)
}
func (b *beamReader) readDevice(dst *devBuf) (int, error) {
	if b.err != nil {
		return 0, b.err
	}
	n := (dst.n / gpuBlock) * gpuBlock // ConvertReader granularity (convert.go:43-44)
	if n == 0 {
		return 0, sdr.ErrShortBuffer
	}
	f := b.readers[0].SampleFormat()
	ptrs := make([]unsafe.Pointer, len(b.readers))
	for i, r := range b.readers {
		if i >= len(b.raw) || b.raw[i].n < n {
			buf, err := newDevBuf(b.ctx, f, n)
			if err != nil {
				return 0, err
			}
			if i >= len(b.raw) {
				b.raw = append(b.raw, buf)
			} else {
				b.raw[i] = buf
			}
		}
		if _, err := readFullToDevice(b.ctx, r, b.raw[i].slice(0, n), &b.stage); err != nil {
			b.err = err
			return 0, err
		}
		ptrs[i] = b.raw[i].ptr
	}
	return n, cuda.Translate(b.ctx.Raw().Beamform(int(f), ptrs, b.w, n, dst.ptr))
}

// SetPhaseAngles will set the phase angle to shift every stream by (beamform.go:131-139).
func (b *Beamform) SetPhaseAngles(angles []complex64) error {
	if len(angles) != len(b.impl.readers) {
		return fmt.Errorf("Beamform.SetPhaseAngles: angles must match the reader length")
	}
	b.impl.w = append([]complex64(nil), angles...)
	return nil
}

// ReadBeamform: stream/beamform.go:148-171.
func ReadBeamform(rs sdr.Readers, cfg BeamformConfig) (*Beamform, error) {
	if len(rs) == 0 {
		return nil, fmt.Errorf("stream.Add: No readers passed")
	}
	c, err := ctxFor(rs[0])
	if err != nil {
		return nil, err
	}
	w := make([]complex64, len(rs))
	for i := range w {
		w[i] = 1 // Multiply(reader, 1), beamform.go:155
	}
	impl := &beamReader{ctx: c, readers: rs, w: w}
	b := &Beamform{Reader: impl, impl: impl}
	if len(cfg.Angles) == len(rs) {
		b.SetPhaseAngles(cfg.Angles)
	}
	return b, nil
}

// BeamformAngles2D / BeamformAngles: stream/beamform.go:57-128 (fp64 host math in the library).
func BeamformAngles2D(frequency rf.Hz, angle float64, center [2]float64, antennas [][2]float64) []complex64 {
	return hzcuda.BeamformAngles2D(float64(frequency), angle, center, antennas)
}
func BeamformAngles(frequency rf.Hz, angle float64, distances []float64) []complex64 {
	if len(distances) == 0 {
		return nil
	}
	antennas := make([][2]float64, len(distances))
	for i := range antennas {
		antennas[i] = [2]float64{distances[i], 0}
	}
	return BeamformAngles2D(frequency, angle, antennas[0], antennas)
}

// CudaRingAllocator is a RingBufferOptions.IQBufferAllocator (ring.go:60-64) that backs the whole
// ring with one pinned allocation, so rtl/hackrf/pluto/uhd RX callbacks land raw samples in
// DMA-able memory with no driver change beyond passing this option.
func CudaRingAllocator(format sdr.SampleFormat, opts RingBufferOptions) (sdr.Samples, error) {
	s, _, err := cuda.PinnedSamples(format, opts.Slots*opts.SlotLength)
	return s, err
}
