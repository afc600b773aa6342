// {{{ Copyright (c) the hzsdr-cuda authors, MIT (same terms as hz.tools/sdr) }}}

//go:build sdr.cuda

// build_cuda.go -- the `sdr.cuda` twin of debug/build.go (which gains `//go:build !sdr.cuda`).
// ReadBuildInfo keeps its signature and every field of BuildInfo (debug/build.go:41-56); it gains one:
// CUDA, the backend the sample chain runs on under this tag.
package debug

import (
	"encoding/binary"

	"hz.tools/sdr"
	"hz.tools/sdr/cuda"
	"hz.tools/sdr/internal"
	"hz.tools/sdr/internal/simd"
)

// SIMDInfo: debug/build.go:31-39, unchanged (the SIMD kernels are still compiled; host-side helpers use them).
type SIMDInfo struct {
	// Enabled is True if using SIMD ASM instructions in the backend, False if using the pure-go implementation.
	Enabled bool

	// Backends is a list of the SIMD backends in use.
	Backends []string
}

// CUDAInfo describes the GPU backend behind ConvertBuffer and the stream.* readers.
type CUDAInfo struct {
	// Enabled is true when at least one usable (sm_100) device was found.  When it is false the readers'
	// constructors return Error: there is no CPU fallback under sdr.cuda.
	Enabled bool

	// Library is libhzsdrcuda's version string (hzsdr_version).
	Library string

	// Devices lists the GPUs the library accepts (hzsdr_ctx_info).
	Devices []cuda.DeviceInfo

	// Error is why the backend is unavailable, when it is.
	Error error
}

// BuildInfo: debug/build.go:41-56 plus CUDA.
type BuildInfo struct {
	// SampleFormats will return all known sdr.SampleFormats understood by this compiled version of hz.tools/sdr
	SampleFormats []sdr.SampleFormat

	// RadioDrivers is a string separated list of known radio drivers.
	RadioDrivers []string

	// SIMD will return the compile-time SIMD support.
	SIMD SIMDInfo

	// CUDA will return the GPU backend's status.
	CUDA CUDAInfo

	// HostEndianness will return the detected host ByteOrder.
	HostEndianness binary.ByteOrder
}

// ReadBuildInfo: debug/build.go:60-75.
func ReadBuildInfo() BuildInfo {
	gpu := cuda.ReadInfo()
	return BuildInfo{
		SampleFormats: []sdr.SampleFormat{sdr.SampleFormatC64, sdr.SampleFormatI16, sdr.SampleFormatU8, sdr.SampleFormatI8},
		RadioDrivers:  radioDrivers,
		SIMD:          SIMDInfo{Backends: simd.Backends, Enabled: simd.Enabled},
		CUDA:          CUDAInfo{Enabled: len(gpu.Devices) > 0, Library: gpu.Library, Devices: gpu.Devices, Error: gpu.Err},
		HostEndianness: internal.NativeEndian,
	}
}
