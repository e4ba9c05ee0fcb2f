#!/bin/bash
# Builds libq3tts_b200.so (sm_100a only) next to the Python package.  nvcc cross-compiles without a GPU.
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
out="$here/../qwen3_tts_b200/libq3tts_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
ARCH=(-gencode arch=compute_100a,code=sm_100a)
FLAGS=("${ARCH[@]}" -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -diag-suppress 177)
objs=(); pids=()
mkdir -p "$here/build"
for f in glue w8_gemv w8_gemm_tc attn_decode attn_prefill sampler engine codec tapgemm_tc encoders frame_ll; do
  src="$here/$f.cu"; obj="$here/build/$f.o"
  if [ ! -f "$obj" ] || [ "$src" -nt "$obj" ] || [ "$here/common.cuh" -nt "$obj" ] || [ "$here/sampler.cuh" -nt "$obj" ] || [ "$here/../../include/q3tts_b200.h" -nt "$obj" ]; then
    "$NVCC" "${FLAGS[@]}" -c "$src" -o "$obj" &
    pids+=($!)
  fi
  objs+=("$obj")
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
"$NVCC" "${ARCH[@]}" -shared --cudart static -o "$out" "${objs[@]}"
echo "built $out"
# profiling build of the persistent kernel (id-coded clock64 stamps compiled in; tools/ll_timing.py loads it through Q3T_LIB)
if [ "${Q3T_BUILD_PROF:-1}" = "1" ]; then
  pobj="$here/build/frame_ll_prof.o"
  if [ ! -f "$pobj" ] || [ "$here/frame_ll.cu" -nt "$pobj" ] || [ "$here/common.cuh" -nt "$pobj" ] || [ "$here/sampler.cuh" -nt "$pobj" ]; then
    "$NVCC" "${FLAGS[@]}" -DLL_STAMPS -c "$here/frame_ll.cu" -o "$pobj"
  fi
  pobjs=(); for o in "${objs[@]}"; do [ "$o" = "$here/build/frame_ll.o" ] && pobjs+=("$pobj") || pobjs+=("$o"); done
  "$NVCC" "${ARCH[@]}" -shared --cudart static -o "$here/../qwen3_tts_b200/libq3tts_b200_prof.so" "${pobjs[@]}"
  echo "built libq3tts_b200_prof.so"
fi
