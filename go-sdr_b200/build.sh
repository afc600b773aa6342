#!/bin/bash
# Builds libhzsdrcuda.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
# The FFT kernels are one translation unit per length so they compile in parallel.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${HZSDR_OUT:-$HERE/lib}"   # experiments build variants elsewhere (HZSDR_LIB selects one at run time)
OBJ="${HZSDR_OBJ:-$HERE/build}"
mkdir -p "$OUT" "$OBJ"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
JOBS="${JOBS:-$(nproc)}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr
       -Xcompiler -fPIC -Xcompiler -Wall -Xcompiler -fvisibility=hidden -DHZSDR_BUILD ${HZSDR_NVCC_EXTRA:-})
FFT_LENGTHS=(16384 8192 1024 4096 2048 512 256 128 64 32 16 8 4 2)

newer() {  # newer <obj> <src...>: true when obj is up to date against every source and header
  local obj="$1"; shift
  [[ -f "$obj" ]] || return 1
  for s in "$@" "$HERE"/csrc/*.cuh "$HERE"/csrc/*.h "$HERE"/../include/hzsdr_cuda.h "$HERE/build.sh"; do
    [[ "$obj" -nt "$s" ]] || return 1
  done
}

cmds=()
objs=()
for f in api elementwise fft chain1024 chain16k chaink fir polyphase beamgroup convert_more bigfft kerberos; do
  objs+=("$OBJ/$f.o")
  newer "$OBJ/$f.o" "$HERE/csrc/$f.cu" || cmds+=("$NVCC ${FLAGS[*]} -c $HERE/csrc/$f.cu -o $OBJ/$f.o")
done
for n in "${FFT_LENGTHS[@]}"; do
  objs+=("$OBJ/fft_$n.o")
  newer "$OBJ/fft_$n.o" "$HERE/csrc/fft_inst.cu" || cmds+=("$NVCC ${FLAGS[*]} -DHZ_FFT_N=$n -c $HERE/csrc/fft_inst.cu -o $OBJ/fft_$n.o")
done
if ((${#cmds[@]})); then
  printf '%s\n' "${cmds[@]}" | xargs -P "$JOBS" -I{} bash -c '{}'
fi
"$NVCC" -shared -o "$OUT/libhzsdrcuda.so" "${objs[@]}" \
  -gencode arch=compute_100a,code=sm_100a -lcudart_static -ldl -lpthread -lrt
echo "built $OUT/libhzsdrcuda.so"
