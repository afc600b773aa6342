// fft_inst.cu -- one translation unit per FFT length: nvcc ... -DHZ_FFT_N=<n>
#ifndef HZ_FFT_N
#error "compile with -DHZ_FFT_N=<power of two in 2..16384>"
#endif
#include "fft_kernels.cuh"
