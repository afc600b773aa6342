// {{{ Copyright (c) the hzsdr-cuda authors, MIT (same terms as hz.tools/sdr) }}}

//go:build sdr.cuda

// conv_cuda.go -- the `sdr.cuda` twin of conv.go and copy.go.  Drop into the root of hz.tools/sdr
// and give conv.go / copy.go the constraint `//go:build !sdr.cuda` (the idiom of iq_u8_amd64.go:21
// vs iq_u8_nosimd.go:21).  ConvertBuffer and CopySamples keep their signatures; what changes is
// that the conversion arithmetic runs on the GPU (hzsdr_convert_to_c64, bit-exact against
// iq_u8.go:111-121 / iq_i8.go:107-119 / iq_i16.go:141-145) and that a 5th, device-resident Samples
// type is understood instead of falling into ErrSampleFormatUnknown (copy.go:36-51).
package sdr

import (
	"fmt"
	"io"
	"unsafe"

	"hz.tools/sdr/internal/hzcuda"
)

// ErrConversionNotImplemented: conv.go:27-31 (that file is excluded under sdr.cuda, so the sentinel
// moves here; same text, and callers keep comparing with ==).
var ErrConversionNotImplemented = fmt.Errorf("sdr: unknown format conversion")

// deviceSamples is implemented by cuda.SamplesC64 (and any future device type).
type deviceSamples interface {
	Samples
	DevicePointer() (unsafe.Pointer, *hzcuda.Ctx)
}

var (
	cudaCtx    *hzcuda.Ctx
	cudaCtxErr error
)

func init() {
	// Like the SIMD gate (internal/simd/enabled_amd64.go:35-50): decide once, fail loudly later.
	cudaCtx, cudaCtxErr = hzcuda.NewCtx(0)
}

func cudaErr(err error) error {
	if e, ok := err.(*hzcuda.Error); ok {
		switch e.Status {
		case hzcuda.ErrDstTooSmall:
			return ErrDstTooSmall
		case hzcuda.ErrFormatMismatch:
			return ErrSampleFormatMismatch
		case hzcuda.ErrFormatUnknown:
			return ErrSampleFormatUnknown
		case hzcuda.ErrConversionNotImplmented:
			return ErrConversionNotImplemented
		}
	}
	return err
}

// ConvertBuffer: conv.go:55-93.  Same-format is CopySamples; src longer than dst is
// ErrDstTooSmall; every other pair of formats runs on the GPU (hzsdr_convert), bit-exact against
// the ToU8 / ToI8 / ToI16 / ToC64 methods of the four sample types.
func ConvertBuffer(dst, src Samples) (int, error) {
	if src.Format() == dst.Format() {
		return CopySamples(dst, src)
	}
	if src.Length() > dst.Length() {
		return 0, ErrDstTooSmall
	}
	if cudaCtxErr != nil {
		return 0, cudaCtxErr // no CPU fallback
	}
	n := src.Length()
	if n == 0 {
		return 0, nil
	}
	ctx := cudaCtx

	// source: already on the device, or staged
	var srcDev unsafe.Pointer
	if ds, ok := src.(deviceSamples); ok {
		srcDev, _ = ds.DevicePointer()
	} else {
		sb, err := UnsafeSamplesAsBytes(src)
		if err != nil {
			return 0, err
		}
		p, err := ctx.Alloc(len(sb))
		if err != nil {
			return 0, cudaErr(err)
		}
		defer ctx.Free(p)
		if err := ctx.UploadGo(p, sb); err != nil {
			return 0, cudaErr(err)
		}
		srcDev = p
	}

	// destination: device-resident stays on the device; a host SamplesC64 gets a D2H
	if dd, ok := dst.(deviceSamples); ok {
		p, _ := dd.DevicePointer()
		got, err := ctx.Convert(int(src.Format()), srcDev, n, int(dst.Format()), p, dst.Length())
		return got, cudaErr(err)
	}
	p, err := ctx.Alloc(n * dst.Format().Size())
	if err != nil {
		return 0, cudaErr(err)
	}
	defer ctx.Free(p)
	got, err := ctx.Convert(int(src.Format()), srcDev, n, int(dst.Format()), p, n)
	if err != nil {
		return 0, cudaErr(err)
	}
	hb, err := UnsafeSamplesAsBytes(dst.Slice(0, got))
	if err != nil {
		return 0, err
	}
	return got, cudaErr(ctx.Download(hb, p))
}

// CopySamples: copy.go:31-52 plus the device cases.
func CopySamples(dst, src Samples) (int, error) {
	if dst.Format() != src.Format() {
		return 0, ErrSampleFormatMismatch
	}
	dd, dstDev := dst.(deviceSamples)
	sd, srcDev := src.(deviceSamples)
	if !dstDev && !srcDev {
		switch dst := dst.(type) { // the reference's four host cases, unchanged
		case SamplesU8:
			return copy(dst, src.(SamplesU8)), nil
		case SamplesI8:
			return copy(dst, src.(SamplesI8)), nil
		case SamplesI16:
			return copy(dst, src.(SamplesI16)), nil
		case SamplesC64:
			return copy(dst, src.(SamplesC64)), nil
		default:
			return 0, ErrSampleFormatUnknown
		}
	}
	if cudaCtxErr != nil {
		return 0, cudaCtxErr
	}
	n := dst.Length()
	if src.Length() < n {
		n = src.Length()
	}
	bytes := n * dst.Format().Size()
	switch {
	case dstDev && srcDev:
		dp, ctx := dd.DevicePointer()
		sp, _ := sd.DevicePointer()
		return n, cudaErr(ctx.Copy(dp, sp, bytes))
	case dstDev:
		dp, ctx := dd.DevicePointer()
		sb, err := UnsafeSamplesAsBytes(src.Slice(0, n))
		if err != nil {
			return 0, err
		}
		return n, cudaErr(ctx.UploadGo(dp, sb))
	default:
		sp, ctx := sd.DevicePointer()
		db, err := UnsafeSamplesAsBytes(dst.Slice(0, n))
		if err != nil {
			return 0, err
		}
		return n, cudaErr(ctx.Download(db, sp))
	}
}

// Copy: copy.go:59-64.  Reader-to-Writer plumbing is host-side and unchanged by sdr.cuda; it lives here
// only because copy.go is excluded (CopySamples above needed the device cases).  rtltcp/server.go:226
// and every other caller compile unchanged.
func Copy(dst Writer, src Reader) (int64, error) {
	if dst.SampleFormat() != src.SampleFormat() {
		return 0, ErrSampleFormatMismatch
	}
	return copyBuffer(dst, src, nil)
}

// CopyBuffer: copy.go:68-76 -- Copy through a caller-provided staging buffer (a cuda.PinnedSamples
// buffer makes every hop DMA-able).
func CopyBuffer(dst Writer, src Reader, buf Samples) (int64, error) {
	if dst.SampleFormat() != src.SampleFormat() || dst.SampleFormat() != buf.Format() {
		return 0, ErrSampleFormatMismatch
	}
	return copyBuffer(dst, src, buf)
}

// copyBuffer: copy.go:80-118.  Read, write what was read, stop at EOF (not an error) or at the first
// failure; a short write is ErrShortWrite.  A nil buf becomes 32 Ki samples of the writer's format.
func copyBuffer(dst Writer, src Reader, buf Samples) (int64, error) {
	if buf == nil {
		b, err := MakeSamples(dst.SampleFormat(), 32*1024)
		if err != nil {
			return 0, err
		}
		buf = b
	}
	var total int64
	for {
		got, rerr := src.Read(buf)
		if got > 0 {
			put, werr := dst.Write(buf.Slice(0, got))
			if put > 0 {
				total += int64(put)
			}
			if werr != nil {
				return total, werr
			}
			if put != got {
				return total, ErrShortWrite
			}
		}
		if rerr == io.EOF {
			return total, nil
		}
		if rerr != nil {
			return total, rerr
		}
	}
}
