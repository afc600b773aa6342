// {{{ Copyright (c) the hzsdr-cuda authors, MIT (same terms as hz.tools/sdr) }}}

//go:build sdr.cuda

// Package hzcuda is the raw cgo binding of libhzsdrcuda.so (include/hzsdr_cuda.h).
//
// It imports nothing from hz.tools/sdr so that both the root `sdr` package (conv_cuda.go,
// copy_cuda.go) and the public `hz.tools/sdr/cuda` package can use it without an import cycle.
// Every exported function is a 1:1 wrapper of one C symbol; status codes are turned into Go
// errors by Err, the same shape as rtl/error.go:31-36 (rvToErr).
//
// NOTE: this file cannot be compiled in the authoring image (no Go toolchain); it is the binding
// a maintainer adds.  The identical calls are exercised through ctypes (tests/) and through the C++
// mirror (go-sdr_b200/host), which are compiled and run on a B200.
package hzcuda

// #cgo LDFLAGS: -lhzsdrcuda
// #cgo static LDFLAGS: -lhzsdrcuda -lcudart_static -ldl -lrt -lpthread -lstdc++
//
// #include <stdlib.h>
// #include <hzsdr_cuda.h>
import "C"

import (
	"fmt"
	"runtime"
	"unsafe"
)

// Status is an hzsdr_status.
type Status int

// Status codes that map onto hz.tools/sdr sentinel errors (iq.go:27-39, conv.go:30); the root
// package translates them, this package only names them.
const (
	OK                         Status = C.HZSDR_OK
	ErrNoDevice                Status = C.HZSDR_ERR_NO_DEVICE
	ErrDstTooSmall             Status = C.HZSDR_ERR_DST_TOO_SMALL
	ErrFormatMismatch          Status = C.HZSDR_ERR_FORMAT_MISMATCH
	ErrFormatUnknown           Status = C.HZSDR_ERR_FORMAT_UNKNOWN
	ErrConversionNotImplmented Status = C.HZSDR_ERR_CONVERSION_NOT_IMPLEMENTED
	ErrRingUnderrun            Status = C.HZSDR_ERR_RING_UNDERRUN
)

// Error carries the library's status and its thread-local message.
type Error struct {
	Status  Status
	Message string
}

func (e *Error) Error() string { return fmt.Sprintf("hzsdrcuda: %s (status %d)", e.Message, e.Status) }

// Err converts a C return value.  The library's message is thread-local, so the failing call and
// hzsdr_last_error must run on the same OS thread: use call, which pins the goroutine for both.
func Err(rc C.int) error {
	if rc == C.HZSDR_OK {
		return nil
	}
	return &Error{Status: Status(rc), Message: C.GoString(C.hzsdr_last_error())}
}

// call runs one C entry point and fetches its error message on the same OS thread (the Go scheduler
// may move a goroutine between two cgo calls, never inside one).
func call(f func() C.int) error {
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	return Err(f())
}

// ptrArray copies device / host pointers into C memory (cgo must not retain Go pointers to Go pointers).
// An empty list gives a one-element array so that callers can always pass a valid address.
func ptrArray(ps []unsafe.Pointer) (*unsafe.Pointer, func()) {
	n := len(ps)
	if n == 0 {
		n = 1
	}
	mem := C.malloc(C.size_t(n) * C.size_t(unsafe.Sizeof(uintptr(0))))
	arr := unsafe.Slice((*unsafe.Pointer)(mem), n)
	copy(arr, ps)
	return (*unsafe.Pointer)(mem), func() { C.free(mem) }
}

// Version is hzsdr_version(); FormatSize is SampleFormat.Size() as the library sees it (iq.go:99-110).
func Version() string            { return C.GoString(C.hzsdr_version()) }
func FormatSize(format int) int  { return int(C.hzsdr_format_size(C.int(format))) }

// Ctx is one GPU + one CUDA stream.
type Ctx struct{ h *C.hzsdr_ctx }

// NewCtx fails loudly when there is no B200: there is no CPU fallback, in the spirit of the SIMD
// CPU-feature gate (internal/simd/enabled_amd64.go:35-50).
func NewCtx(device int) (*Ctx, error) {
	var h *C.hzsdr_ctx
	if err := call(func() C.int { return C.hzsdr_ctx_create(C.int(device), &h) }); err != nil {
		return nil, err
	}
	return &Ctx{h: h}, nil
}

func (c *Ctx) Close() error { return call(func() C.int { return C.hzsdr_ctx_destroy(c.h) }) }
func (c *Ctx) Sync() error  { return call(func() C.int { return C.hzsdr_ctx_sync(c.h) }) }

// Info is what debug.ReadBuildInfo reports for a device (debug/build.go:60-75).
type Info struct {
	Name             string
	SMMajor, SMMinor int
	SMCount          int
	HBMBytes         uint64
}

func (c *Ctx) Info() (Info, error) {
	var name [256]C.char
	var maj, min, sms C.int
	var mem C.size_t
	err := call(func() C.int { return C.hzsdr_ctx_info(c.h, &name[0], C.size_t(len(name)), &maj, &min, &sms, &mem) })
	return Info{Name: C.GoString(&name[0]), SMMajor: int(maj), SMMinor: int(min), SMCount: int(sms), HBMBytes: uint64(mem)}, err
}

// Stream is the context's cudaStream_t, for interop with other CUDA code in the process.
func (c *Ctx) Stream() (unsafe.Pointer, error) {
	var st unsafe.Pointer
	err := call(func() C.int { return C.hzsdr_ctx_stream(c.h, &st) })
	return st, err
}

func (c *Ctx) Memset(dev unsafe.Pointer, value byte, bytes int) error {
	return call(func() C.int { return C.hzsdr_dev_memset(c.h, dev, C.int(value), C.size_t(bytes)) })
}

// I16ShiftLSBToMSB is SamplesI16.ShiftLSBToMSBBits on a device buffer (iq_i16.go:103-111).
func (c *Ctx) I16ShiftLSBToMSB(buf unsafe.Pointer, n, bits int) error {
	return call(func() C.int { return C.hzsdr_i16_shift_lsb_to_msb(c.h, buf, C.size_t(n), C.int(bits)) })
}

func DeviceCount() (int, error) {
	var n C.int
	err := call(func() C.int { return C.hzsdr_device_count(&n) })
	return int(n), err
}

// ---- memory ---------------------------------------------------------------------------------

func (c *Ctx) Alloc(bytes int) (unsafe.Pointer, error) {
	var p unsafe.Pointer
	err := call(func() C.int { return C.hzsdr_dev_alloc(c.h, C.size_t(bytes), &p) })
	return p, err
}
func (c *Ctx) Free(p unsafe.Pointer) error { return call(func() C.int { return C.hzsdr_dev_free(c.h, p) }) }

// PinnedAlloc returns cudaHostAlloc'd memory.  It is C memory: safe to hand to async copies and
// to wrap with yikes.Samples (yikes/bytes.go:50-71).
func PinnedAlloc(bytes int) (unsafe.Pointer, error) {
	var p unsafe.Pointer
	err := call(func() C.int { return C.hzsdr_pinned_alloc(C.size_t(bytes), &p) })
	return p, err
}
func PinnedFree(p unsafe.Pointer) error { return call(func() C.int { return C.hzsdr_pinned_free(p) }) }

// Upload copies from pinned (C) memory only: cgo forbids C retaining a Go pointer after the call
// returns and this copy is asynchronous.  UploadGo is the synchronous form for Go slices.
func (c *Ctx) Upload(dst, srcPinned unsafe.Pointer, bytes int) error {
	return call(func() C.int { return C.hzsdr_upload(c.h, dst, srcPinned, C.size_t(bytes)) })
}
func (c *Ctx) UploadGo(dst unsafe.Pointer, src []byte) error {
	if len(src) == 0 {
		return nil
	}
	if err := call(func() C.int { return C.hzsdr_upload(c.h, dst, unsafe.Pointer(&src[0]), C.size_t(len(src))) }); err != nil {
		return err
	}
	return c.Sync() // the Go slice may move or die once we return
}
func (c *Ctx) Download(dst []byte, src unsafe.Pointer) error {
	if len(dst) == 0 {
		return nil
	}
	return call(func() C.int { return C.hzsdr_download(c.h, unsafe.Pointer(&dst[0]), src, C.size_t(len(dst))) })
}
func (c *Ctx) Copy(dst, src unsafe.Pointer, bytes int) error {
	return call(func() C.int { return C.hzsdr_copy(c.h, dst, src, C.size_t(bytes)) })
}

// ---- kernels ----------------------------------------------------------------------------------

func (c *Ctx) ConvertToC64(format int, src unsafe.Pointer, srcLen int, dst unsafe.Pointer, dstLen int) (int, error) {
	var n C.size_t
	err := call(func() C.int { return C.hzsdr_convert_to_c64(c.h, C.int(format), src, C.size_t(srcLen), dst, C.size_t(dstLen), &n) })
	return int(n), err
}

// Convert is the full ConvertBuffer matrix (conv.go:55-93); dstFormat == C64 is ConvertToC64.
func (c *Ctx) Convert(srcFormat int, src unsafe.Pointer, srcLen int, dstFormat int, dst unsafe.Pointer, dstLen int) (int, error) {
	var n C.size_t
	err := call(func() C.int { return C.hzsdr_convert(c.h, C.int(srcFormat), src, C.size_t(srcLen), C.int(dstFormat), dst, C.size_t(dstLen), &n) })
	return int(n), err
}

// Lookup is LookupTable.Lookup (iq_lookup_table.go:129-147): dst[i] = table[src[i] as uint16].
func (c *Ctx) Lookup(srcFormat int, src unsafe.Pointer, n int, tableFormat int, table, dst unsafe.Pointer, dstLen int) error {
	return call(func() C.int { return C.hzsdr_lookup(c.h, C.int(srcFormat), src, C.size_t(n), C.int(tableFormat), table, dst, C.size_t(dstLen)) })
}

// AddInt is stream.Add on I8 / I16 readers (stream/add.go:95-113).
func (c *Ctx) AddInt(format int, dst unsafe.Pointer, srcs []unsafe.Pointer, n int) error {
	arr, free := ptrArray(srcs)
	defer free()
	return call(func() C.int { return C.hzsdr_add_int(c.h, C.int(format), dst, arr, C.int(len(srcs)), C.size_t(n)) })
}

// FftConvolve is fft.Convolve / fft.CrossCorrelate (fft/convolution.go:97-139) over `batch` vectors.
func (c *Ctx) FftConvolve(dst, iq1, iq2 unsafe.Pointer, n, batch int, crossCorrelate bool, scratch unsafe.Pointer) error {
	xc := C.int(0)
	if crossCorrelate {
		xc = 1
	}
	return call(func() C.int { return C.hzsdr_fft_convolve(c.h, dst, iq1, iq2, C.size_t(n), C.size_t(batch), xc, scratch) })
}

// FftShiftScale is FFTShiftAndScale (rtl/kerberos/internal/reader.go:57-64) over `batch` length-n
// vectors on the device, in place.
func (c *Ctx) FftShiftScale(data unsafe.Pointer, n, batch int, scale float32) error {
	return call(func() C.int { return C.hzsdr_fftshift_scale(c.h, data, C.size_t(n), C.size_t(batch), C.float(scale)) })
}

// Graft is one pass of GraftReaders' loop (rtl/kerberos/internal/graft.go:96-125): nReaders buffers
// of fftSize samples in, one buffer of nReaders*fftSize samples out; freq is device scratch of the
// output's size.
func (c *Ctx) Graft(iq unsafe.Pointer, nReaders, fftSize int, dst, freq unsafe.Pointer) error {
	return call(func() C.int { return C.hzsdr_graft(c.h, iq, C.size_t(nReaders), C.size_t(fftSize), dst, freq) })
}

// CorrelatePeak is checkAlignment's peak search (rtl/kerberos/internal/align.go:125-146) over `batch`
// correlation vectors of length n; it synchronises and returns one signed offset per vector.
func (c *Ctx) CorrelatePeak(cc unsafe.Pointer, n, batch int) ([]int32, error) {
	out := make([]int32, batch)
	if batch == 0 {
		return out, nil
	}
	err := call(func() C.int { return C.hzsdr_correlate_peak(c.h, cc, C.size_t(n), C.size_t(batch), (*C.int32_t)(unsafe.Pointer(&out[0]))) })
	return out, err
}

// PhaseOffsets is rtl/kerberos/internal/align.go:244-272 over nChan device buffers of n samples
// (channel-major); it synchronises and returns one unit phasor per channel.
func (c *Ctx) PhaseOffsets(bufs unsafe.Pointer, nChan, n int) ([]complex64, error) {
	out := make([]complex64, nChan)
	if nChan == 0 {
		return out, nil
	}
	err := call(func() C.int { return C.hzsdr_phase_offsets(c.h, bufs, C.size_t(nChan), C.size_t(n), (*C.float)(unsafe.Pointer(&out[0]))) })
	return out, err
}

// Nco is the ShiftBuffer closure state (stream/shifter.go:67-71).
type Nco struct {
	SampleRate uint32
	Ts         float64
}

func (c *Ctx) Shift(buf unsafe.Pointer, n int, freqHz float64, st *Nco) error {
	cs := C.hzsdr_nco{sample_rate: C.uint32_t(st.SampleRate), ts: C.double(st.Ts)}
	err := call(func() C.int { return C.hzsdr_shift(c.h, buf, C.size_t(n), C.double(freqHz), &cs) })
	st.Ts = float64(cs.ts)
	return err
}

// ConvertShift is ConvertReader + ShiftReader in one pass over HBM (raw -> complex64 -> mixed).
func (c *Ctx) ConvertShift(format int, src unsafe.Pointer, n int, dst unsafe.Pointer, dstLen int, freqHz float64, st *Nco) error {
	cs := C.hzsdr_nco{sample_rate: C.uint32_t(st.SampleRate), ts: C.double(st.Ts)}
	err := call(func() C.int {
		return C.hzsdr_convert_shift(c.h, C.int(format), src, C.size_t(n), dst, C.size_t(dstLen), C.double(freqHz), &cs)
	})
	st.Ts = float64(cs.ts)
	return err
}

// ConvertShiftBatch runs len(srcs) consecutive buffers of one stream (nEach samples each, e.g. drained ring slots)
// through the fused ConvertReader + ShiftReader in one kernel launch per <= 64 buffers.
func (c *Ctx) ConvertShiftBatch(format int, srcs []unsafe.Pointer, nEach int, dsts []unsafe.Pointer, dstLenEach int, freqHz float64, st *Nco) error {
	if len(srcs) != len(dsts) {
		return &Error{Status: Status(C.HZSDR_ERR_INVALID), Message: "ConvertShiftBatch: len(srcs) != len(dsts)"}
	}
	sa, freeS := ptrArray(srcs)
	defer freeS()
	da, freeD := ptrArray(dsts)
	defer freeD()
	cs := C.hzsdr_nco{sample_rate: C.uint32_t(st.SampleRate), ts: C.double(st.Ts)}
	err := call(func() C.int {
		return C.hzsdr_convert_shift_batch(c.h, C.int(format), sa, C.size_t(nEach), da, C.size_t(dstLenEach), C.size_t(len(srcs)), C.double(freqHz), &cs)
	})
	st.Ts = float64(cs.ts)
	return err
}
func (c *Ctx) Rotate(buf unsafe.Pointer, n int, m complex64) error {
	return call(func() C.int { return C.hzsdr_rotate(c.h, buf, C.size_t(n), C.float(real(m)), C.float(imag(m))) })
}
func (c *Ctx) Scale(buf unsafe.Pointer, n int, r float32) error {
	return call(func() C.int { return C.hzsdr_scale(c.h, buf, C.size_t(n), C.float(r)) })
}
func (c *Ctx) Add(dst unsafe.Pointer, srcs []unsafe.Pointer, n int) error {
	arr, free := ptrArray(srcs) // the pointer array itself must be C memory, for the duration of the call only
	defer free()
	return call(func() C.int { return C.hzsdr_add(c.h, dst, arr, C.int(len(srcs)), C.size_t(n)) })
}
func (c *Ctx) Decimate(format int, src unsafe.Pointer, n int, dst unsafe.Pointer, dstLen int, factor uint, block int) (int, error) {
	var out C.size_t
	err := call(func() C.int { return C.hzsdr_decimate(c.h, C.int(format), src, C.size_t(n), dst, C.size_t(dstLen), C.uint(factor), C.size_t(block), &out) })
	return int(out), err
}
func (c *Ctx) Downsample(format int, src unsafe.Pointer, n int, dst unsafe.Pointer, dstLen int, factor uint, block int) (int, error) {
	var out C.size_t
	err := call(func() C.int { return C.hzsdr_downsample(c.h, C.int(format), src, C.size_t(n), dst, C.size_t(dstLen), C.uint(factor), C.size_t(block), &out) })
	return int(out), err
}
func (c *Ctx) ConvolveFreq(src, dst, filter unsafe.Pointer, nFFT, nBlocks int) error {
	return call(func() C.int { return C.hzsdr_convolve_freq(c.h, src, dst, filter, C.size_t(nFFT), C.size_t(nBlocks)) })
}
func (c *Ctx) Beamform(format int, chans []unsafe.Pointer, weights []complex64, n int, dst unsafe.Pointer) error {
	if len(chans) == 0 || len(weights) != len(chans) {
		return &Error{Status: Status(C.HZSDR_ERR_INVALID), Message: "Beamform: one weight per channel, at least one channel"}
	}
	arr, free := ptrArray(chans)
	defer free()
	return call(func() C.int {
		return C.hzsdr_beamform(c.h, C.int(format), arr, C.int(len(chans)), weightsPtr(weights), C.size_t(n), dst)
	})
}

// BeamformSubmitHost is Beamform from pinned HOST channel buffers into a pinned host beam: time
// slices are staged across PCIe, overlapped with the kernel and the return copy.  It only enqueues;
// WaitHost completes it.
func (c *Ctx) BeamformSubmitHost(format int, chans []unsafe.Pointer, weights []complex64, n int, dst unsafe.Pointer) error {
	if len(chans) == 0 || len(weights) != len(chans) {
		return &Error{Status: Status(C.HZSDR_ERR_INVALID), Message: "BeamformSubmitHost: one weight per channel, at least one channel"}
	}
	arr, free := ptrArray(chans)
	defer free()
	return call(func() C.int {
		return C.hzsdr_beamform_submit_host(c.h, C.int(format), arr, C.int(len(chans)), weightsPtr(weights), C.size_t(n), dst)
	})
}

// WaitHost completes everything the SubmitHost calls of this context have enqueued.
func (c *Ctx) WaitHost() error { return call(func() C.int { return C.hzsdr_ctx_wait_host(c.h) }) }

// BeamformAngles2D is stream.BeamformAngles2D's arithmetic (stream/beamform.go:57-107).
func BeamformAngles2D(frequencyHz, angleDeg float64, center [2]float64, antennas [][2]float64) []complex64 {
	if len(antennas) == 0 {
		return nil
	}
	out := make([]complex64, len(antennas))
	C.hzsdr_beamform_angles_2d(C.double(frequencyHz), C.double(angleDeg), (*C.double)(&center[0]),
		(*C.double)(&antennas[0][0]), C.int(len(antennas)), (*C.float)(unsafe.Pointer(&out[0])))
	return out
}

// ---- FFT plan ---------------------------------------------------------------------------------

type Plan struct{ h *C.hzsdr_fft_plan }

func (c *Ctx) NewPlan(iqLen, freqLen int, forward bool) (*Plan, error) {
	dir := C.int(C.HZSDR_FFT_BACKWARD)
	if forward {
		dir = C.int(C.HZSDR_FFT_FORWARD)
	}
	var h *C.hzsdr_fft_plan
	if err := call(func() C.int { return C.hzsdr_fft_plan_create(c.h, C.size_t(iqLen), C.size_t(freqLen), dir, &h) }); err != nil {
		return nil, err
	}
	return &Plan{h: h}, nil
}
func (p *Plan) Exec(src, dst unsafe.Pointer, batch int) error {
	return call(func() C.int { return C.hzsdr_fft_exec(p.h, src, dst, C.size_t(batch)) })
}
func (p *Plan) Close() error { return call(func() C.int { return C.hzsdr_fft_plan_destroy(p.h) }) }

// ---- fused chain ------------------------------------------------------------------------------

type ChainConfig struct {
	SrcFormat     int
	SampleRate    uint32
	ShiftHz       float64
	Filter        []complex64 // frequency domain, len = ConvolutionReader block
	Decimate      uint32
	DecimateBlock uint32
	I16LsbBits    int
	// OverlapSaveTaps > 0: Filter is the spectrum of that many time-domain taps and the chain computes the
	// true linear convolution by overlap-save (extension; 0 = the reference's block-circular reader).
	OverlapSaveTaps uint32
}
type Chain struct{ h *C.hzsdr_chain }

// cFilter copies a filter into C memory: the config struct is Go memory handed to C, so it must not carry a
// Go pointer (cgo's pointer-passing rules); the library copies the filter before hzsdr_chain_create returns.
func cFilter(filter []complex64) (unsafe.Pointer, func(), error) {
	if len(filter) == 0 {
		return nil, func() {}, &Error{Status: Status(C.HZSDR_ERR_INVALID), Message: "empty filter"}
	}
	mem := C.malloc(C.size_t(len(filter)) * 8)
	copy(unsafe.Slice((*complex64)(mem), len(filter)), filter)
	return mem, func() { C.free(mem) }, nil
}

func (c *Ctx) NewChain(cfg ChainConfig) (*Chain, error) {
	filt, freeFilt, err := cFilter(cfg.Filter)
	if err != nil {
		return nil, err
	}
	defer freeFilt()
	cc := C.hzsdr_chain_config{
		src_format: C.int(cfg.SrcFormat), sample_rate: C.uint32_t(cfg.SampleRate), shift_hz: C.double(cfg.ShiftHz),
		n_fft: C.size_t(len(cfg.Filter)), filter_host: filt,
		decimate: C.uint32_t(cfg.Decimate), decimate_block: C.uint32_t(cfg.DecimateBlock), i16_lsb_bits: C.int(cfg.I16LsbBits),
		overlap_save_taps: C.uint32_t(cfg.OverlapSaveTaps),
	}
	var h *C.hzsdr_chain
	if err := call(func() C.int { return C.hzsdr_chain_create(c.h, &cc, &h) }); err != nil {
		return nil, err
	}
	return &Chain{h: h}, nil
}
func (ch *Chain) Close() error { return call(func() C.int { return C.hzsdr_chain_destroy(ch.h) }) }
func (ch *Chain) OutLen(n int) int {
	var out C.size_t
	C.hzsdr_chain_out_len(ch.h, C.size_t(n), &out)
	return int(out)
}
func (ch *Chain) Exec(src unsafe.Pointer, n int, dst unsafe.Pointer, dstLen int) (int, error) {
	var out C.size_t
	err := call(func() C.int { return C.hzsdr_chain_exec(ch.h, src, C.size_t(n), dst, C.size_t(dstLen), &out) })
	return int(out), err
}

// ExecBatch runs len(srcs) consecutive buffers of the stream (device pointers, nEach samples each) in one
// cgo call; every buffer emits the returned number of samples into its dsts entry.
func (ch *Chain) ExecBatch(srcs []unsafe.Pointer, nEach int, dsts []unsafe.Pointer, dstLenEach int) (int, error) {
	if len(srcs) != len(dsts) {
		return 0, &Error{Status: Status(C.HZSDR_ERR_INVALID), Message: "ExecBatch: len(srcs) != len(dsts)"}
	}
	sa, freeS := ptrArray(srcs)
	defer freeS()
	da, freeD := ptrArray(dsts)
	defer freeD()
	var out C.size_t
	err := call(func() C.int {
		return C.hzsdr_chain_exec_batch(ch.h, sa, C.size_t(nEach), da, C.size_t(dstLenEach), C.size_t(len(srcs)), &out)
	})
	return int(out), err
}

// ExecHost is Exec from / to host memory and waits: H2D, the fused kernel, D2H.  The buffers are only
// read / written inside the call, so Go memory is fine here (SubmitHost needs pinned C memory).
func (ch *Chain) ExecHost(src unsafe.Pointer, n int, dst unsafe.Pointer, dstLen int) (int, error) {
	var out C.size_t
	err := call(func() C.int { return C.hzsdr_chain_exec_host(ch.h, src, C.size_t(n), dst, C.size_t(dstLen), &out) })
	return int(out), err
}

// SubmitHost / WaitHost: both buffers MUST be pinned C memory (ring slots / PinnedAlloc).
func (ch *Chain) SubmitHost(srcPinned unsafe.Pointer, n int, dstPinned unsafe.Pointer, dstLen int) (int, error) {
	var out C.size_t
	err := call(func() C.int { return C.hzsdr_chain_submit_host(ch.h, srcPinned, C.size_t(n), dstPinned, C.size_t(dstLen), &out) })
	return int(out), err
}

// SubmitRing sends the ring's next unread slot (raw samples a producer wrote through WritePeek / WritePoke,
// stream/ring.go:344-392) through the chain and copies the result to dstPinned behind the kernel.
// ErrRingUnderrun when no slot is pending.  WaitHost completes it.
func (ch *Chain) SubmitRing(r *Ring, dstPinned unsafe.Pointer, dstLen int) (int, error) {
	var out C.size_t
	err := call(func() C.int { return C.hzsdr_chain_submit_ring(ch.h, r.h, dstPinned, C.size_t(dstLen), &out) })
	return int(out), err
}
func (ch *Chain) WaitHost() error { return call(func() C.int { return C.hzsdr_chain_wait_host(ch.h) }) }
func (ch *Chain) Ts() float64 {
	var ts C.double
	C.hzsdr_chain_get_ts(ch.h, &ts)
	return float64(ts)
}
func (ch *Chain) SetTs(ts float64) { C.hzsdr_chain_set_ts(ch.h, C.double(ts)) }

// ---- pinned ring ------------------------------------------------------------------------------

type Ring struct{ h *C.hzsdr_ring }

func (c *Ctx) NewRing(format, slots, slotLen int) (*Ring, error) {
	var h *C.hzsdr_ring
	if err := call(func() C.int { return C.hzsdr_ring_create(c.h, C.int(format), C.size_t(slots), C.size_t(slotLen), &h) }); err != nil {
		return nil, err
	}
	return &Ring{h: h}, nil
}
func (r *Ring) Close() error { return call(func() C.int { return C.hzsdr_ring_destroy(r.h) }) }
func (r *Ring) WritePeek() (unsafe.Pointer, error) {
	var p unsafe.Pointer
	err := call(func() C.int { return C.hzsdr_ring_write_peek(r.h, &p) })
	return p, err
}
func (r *Ring) WritePoke(n int) error { return call(func() C.int { return C.hzsdr_ring_write_poke(r.h, C.size_t(n)) }) }
func (r *Ring) Read() (unsafe.Pointer, int, error) {
	var p unsafe.Pointer
	var n C.size_t
	err := call(func() C.int { return C.hzsdr_ring_read(r.h, &p, &n) })
	return p, int(n), err
}
func (r *Ring) ReadDone() error { return call(func() C.int { return C.hzsdr_ring_read_done(r.h) }) }

// ---- channelizer: many independent chains, one launch per buffer set ---------------------------

type Channelizer struct {
	h *C.hzsdr_channelizer
	n int
}

func (c *Ctx) NewChannelizer(cfg ChainConfig, shiftHz []float64) (*Channelizer, error) {
	if len(shiftHz) == 0 {
		return nil, &Error{Status: Status(C.HZSDR_ERR_INVALID), Message: "NewChannelizer: no streams"}
	}
	filt, freeFilt, err := cFilter(cfg.Filter)
	if err != nil {
		return nil, err
	}
	defer freeFilt()
	cc := C.hzsdr_chain_config{
		src_format: C.int(cfg.SrcFormat), sample_rate: C.uint32_t(cfg.SampleRate),
		n_fft: C.size_t(len(cfg.Filter)), filter_host: filt,
		decimate: C.uint32_t(cfg.Decimate), decimate_block: C.uint32_t(cfg.DecimateBlock), i16_lsb_bits: C.int(cfg.I16LsbBits),
	}
	var h *C.hzsdr_channelizer
	if err := call(func() C.int { return C.hzsdr_channelizer_create(c.h, &cc, (*C.double)(&shiftHz[0]), C.size_t(len(shiftHz)), &h) }); err != nil {
		return nil, err
	}
	return &Channelizer{h: h, n: len(shiftHz)}, nil
}

// Exec: srcs / dsts are device pointers, one per stream; every stream consumes n samples.
func (z *Channelizer) Exec(srcs []unsafe.Pointer, n int, dsts []unsafe.Pointer, dstLen int) (int, error) {
	if len(srcs) != z.n || len(dsts) != z.n {
		return 0, &Error{Status: Status(C.HZSDR_ERR_INVALID), Message: "Channelizer.Exec: one source and one destination per stream"}
	}
	sa, freeS := ptrArray(srcs)
	defer freeS()
	da, freeD := ptrArray(dsts)
	defer freeD()
	var out C.size_t
	err := call(func() C.int { return C.hzsdr_channelizer_exec(z.h, sa, C.size_t(n), da, C.size_t(dstLen), &out) })
	return int(out), err
}

// SubmitHost is Exec from pinned HOST buffers (cuda.PinnedSamples / ring slots): the streams cross
// PCIe in groups, overlapped with the kernel and the return copies.  It only enqueues; Ctx.WaitHost
// completes it, and the buffers must stay untouched until then.
func (z *Channelizer) SubmitHost(srcs []unsafe.Pointer, n int, dsts []unsafe.Pointer, dstLen int) (int, error) {
	if len(srcs) != z.n || len(dsts) != z.n {
		return 0, &Error{Status: Status(C.HZSDR_ERR_INVALID), Message: "Channelizer.SubmitHost: one source and one destination per stream"}
	}
	sa, freeS := ptrArray(srcs)
	defer freeS()
	da, freeD := ptrArray(dsts)
	defer freeD()
	var out C.size_t
	err := call(func() C.int { return C.hzsdr_channelizer_submit_host(z.h, sa, C.size_t(n), da, C.size_t(dstLen), &out) })
	return int(out), err
}

// Ts / SetTs: the carried NCO time of every stream (checkpoint / resume, stream/shifter.go:68).
func (z *Channelizer) Ts() ([]float64, error) {
	ts := make([]float64, z.n)
	err := call(func() C.int { return C.hzsdr_channelizer_get_ts(z.h, (*C.double)(unsafe.Pointer(&ts[0]))) })
	return ts, err
}
func (z *Channelizer) SetTs(ts []float64) error {
	if len(ts) != z.n {
		return &Error{Status: Status(C.HZSDR_ERR_INVALID), Message: "Channelizer.SetTs: one value per stream"}
	}
	return call(func() C.int { return C.hzsdr_channelizer_set_ts(z.h, (*C.double)(unsafe.Pointer(&ts[0]))) })
}
func (z *Channelizer) Close() error { return call(func() C.int { return C.hzsdr_channelizer_destroy(z.h) }) }

// ---- FIR extension (no reference counterpart) --------------------------------------------------

type Fir struct{ h *C.hzsdr_fir }

func (c *Ctx) NewFir(taps []complex64, decimate uint, method int) (*Fir, error) {
	if len(taps) == 0 {
		return nil, &Error{Status: Status(C.HZSDR_ERR_INVALID), Message: "NewFir: no taps"}
	}
	var h *C.hzsdr_fir
	if err := call(func() C.int { return C.hzsdr_fir_create(c.h, (*C.float)(unsafe.Pointer(&taps[0])), C.size_t(len(taps)), C.uint(decimate), C.int(method), &h) }); err != nil {
		return nil, err
	}
	return &Fir{h: h}, nil
}
func (f *Fir) Exec(src unsafe.Pointer, n int, dst unsafe.Pointer, dstLen int) (int, error) {
	var out C.size_t
	err := call(func() C.int { return C.hzsdr_fir_exec(f.h, src, C.size_t(n), dst, C.size_t(dstLen), &out) })
	return int(out), err
}
func (f *Fir) Reset() error { return call(func() C.int { return C.hzsdr_fir_reset(f.h) }) }
func (f *Fir) Close() error { return call(func() C.int { return C.hzsdr_fir_destroy(f.h) }) }

// ---- fused polyphase decimator on a raw stream (extension) ------------------------------------

// Polyphase is Convert -> Shift -> real-tap FIR -> keep every D-th sample as one kernel
// (hzsdr_polyphase_*): true linear convolution, history and the NCO time carried between calls.
type Polyphase struct{ h *C.hzsdr_polyphase }

func (c *Ctx) NewPolyphase(format int, sampleRate uint32, shiftHz float64, taps []float32, decimate uint, i16LsbBits int) (*Polyphase, error) {
	if len(taps) == 0 {
		return nil, &Error{Status: Status(C.HZSDR_ERR_INVALID), Message: "NewPolyphase: no taps"}
	}
	ct := (*C.float)(C.malloc(C.size_t(len(taps)) * 4)) // taps cross the boundary through C memory
	defer C.free(unsafe.Pointer(ct))
	copy(unsafe.Slice((*float32)(unsafe.Pointer(ct)), len(taps)), taps)
	var h *C.hzsdr_polyphase
	if err := call(func() C.int {
		return C.hzsdr_polyphase_create(c.h, C.int(format), C.uint32_t(sampleRate), C.double(shiftHz), ct, C.size_t(len(taps)), C.uint(decimate), C.int(i16LsbBits), &h)
	}); err != nil {
		return nil, err
	}
	return &Polyphase{h: h}, nil
}

// OutLen is the number of samples the next n input samples will produce.
func (p *Polyphase) OutLen(n int) (int, error) {
	var out C.size_t
	err := call(func() C.int { return C.hzsdr_polyphase_out_len(p.h, C.size_t(n), &out) })
	return int(out), err
}
func (p *Polyphase) Exec(src unsafe.Pointer, n int, dst unsafe.Pointer, dstLen int) (int, error) {
	var out C.size_t
	err := call(func() C.int { return C.hzsdr_polyphase_exec(p.h, src, C.size_t(n), dst, C.size_t(dstLen), &out) })
	return int(out), err
}
func (p *Polyphase) Ts() (float64, error) {
	var ts C.double
	err := call(func() C.int { return C.hzsdr_polyphase_get_ts(p.h, &ts) })
	return float64(ts), err
}
func (p *Polyphase) SetTs(ts float64) error {
	return call(func() C.int { return C.hzsdr_polyphase_set_ts(p.h, C.double(ts)) })
}
func (p *Polyphase) Close() error { return call(func() C.int { return C.hzsdr_polyphase_destroy(p.h) }) }

// ---- multi-GPU Beamform -----------------------------------------------------------------------

// Comm is the NCCL communicator (one process or goroutine-group per GPU).
type Comm struct{ h *C.hzsdr_comm }

func CommUniqueID() ([C.HZSDR_NCCL_UNIQUE_ID_BYTES]byte, error) {
	var id [C.HZSDR_NCCL_UNIQUE_ID_BYTES]byte
	err := call(func() C.int { return C.hzsdr_comm_unique_id(unsafe.Pointer(&id[0])) })
	return id, err
}
func (c *Ctx) NewComm(nranks, rank int, id [C.HZSDR_NCCL_UNIQUE_ID_BYTES]byte) (*Comm, error) {
	var h *C.hzsdr_comm
	if err := call(func() C.int { return C.hzsdr_comm_create(c.h, C.int(nranks), C.int(rank), unsafe.Pointer(&id[0]), &h) }); err != nil {
		return nil, err
	}
	return &Comm{h: h}, nil
}
func (m *Comm) ReduceC64(buf unsafe.Pointer, n, root int) error {
	return call(func() C.int { return C.hzsdr_comm_reduce_c64(m.h, buf, C.size_t(n), C.int(root)) })
}
func (m *Comm) AllReduceC64(buf unsafe.Pointer, n int) error {
	return call(func() C.int { return C.hzsdr_comm_allreduce_c64(m.h, buf, C.size_t(n)) })
}
func (m *Comm) Close() error { return call(func() C.int { return C.hzsdr_comm_destroy(m.h) }) }

// BeamGroup is the Beamform whose reduce-scatter is fused into the kernel over NVLink peer memory.
type BeamGroup struct{ h *C.hzsdr_beam_group }

// NewBeamGroup returns the group and the 64-byte IPC handle every other rank needs (Connect).  n: samples
// per buffer; maxBatch: buffers per exchange the staging area is sized for (1..64).
func (c *Ctx) NewBeamGroup(nranks, rank, n, maxBatch int) (*BeamGroup, [C.HZSDR_IPC_HANDLE_BYTES]byte, error) {
	var h *C.hzsdr_beam_group
	var handle [C.HZSDR_IPC_HANDLE_BYTES]byte
	err := call(func() C.int {
		return C.hzsdr_beam_group_create(c.h, C.int(nranks), C.int(rank), C.size_t(n), C.size_t(maxBatch), unsafe.Pointer(&handle[0]), &h)
	})
	return &BeamGroup{h: h}, handle, err
}
func (g *BeamGroup) Connect(allHandles []byte) error {
	if len(allHandles) == 0 {
		return &Error{Status: Status(C.HZSDR_ERR_INVALID), Message: "BeamGroup.Connect: no handles"}
	}
	return call(func() C.int { return C.hzsdr_beam_group_connect(g.h, unsafe.Pointer(&allHandles[0])) })
}

// weightsPtr: the address of the first weight, or nil for a rank that owns no channel.
func weightsPtr(w []complex64) *C.float {
	if len(w) == 0 {
		return nil
	}
	return (*C.float)(unsafe.Pointer(&w[0]))
}

func (g *BeamGroup) Exec(format int, chans []unsafe.Pointer, weights []complex64, dstSlice unsafe.Pointer) error {
	arr, free := ptrArray(chans)
	defer free()
	return call(func() C.int {
		return C.hzsdr_beam_group_exec(g.h, C.int(format), arr, C.int(len(chans)), weightsPtr(weights), dstSlice)
	})
}

// ExecBatch is ONE exchange over len(dstSlices) buffers: chans[k*nchan+c] = channel c of buffer k.
func (g *BeamGroup) ExecBatch(format int, chans []unsafe.Pointer, nchan int, weights []complex64, dstSlices []unsafe.Pointer) error {
	if len(dstSlices) == 0 || len(chans) != nchan*len(dstSlices) || len(weights) != nchan {
		return &Error{Status: Status(C.HZSDR_ERR_INVALID), Message: "BeamGroup.ExecBatch: nchan pointers per buffer, nchan weights"}
	}
	arr, free := ptrArray(chans)
	defer free()
	dst, freeD := ptrArray(dstSlices)
	defer freeD()
	return call(func() C.int {
		return C.hzsdr_beam_group_exec_batch(g.h, C.int(format), arr, C.int(nchan), weightsPtr(weights), C.size_t(len(dstSlices)), dst)
	})
}
func (g *BeamGroup) Join() error  { return call(func() C.int { return C.hzsdr_beam_group_join(g.h) }) }
func (g *BeamGroup) Close() error { return call(func() C.int { return C.hzsdr_beam_group_destroy(g.h) }) }
