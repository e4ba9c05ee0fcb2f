#!/bin/bash
# Builds libq3tts_b200_<tag>.so with extra nvcc defines for frame_ll.cu (kernel-tuning experiments; load it with Q3T_LIB=...).
# usage: tools/build_variant.sh <tag> -DLL_CWARPS_N=8 ...
set -euo pipefail
tag="$1"; shift
here="$(cd "$(dirname "$0")/../qwen3-tts-apple-silicon_b200/csrc" && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
ARCH=(-gencode arch=compute_100a,code=sm_100a)
"$NVCC" "${ARCH[@]}" -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -diag-suppress 177 "$@" -c "$here/frame_ll.cu" -o "$here/build/frame_ll_$tag.o"
objs=()
for f in glue w8_gemv w8_gemm_tc attn_decode attn_prefill sampler engine codec tapgemm_tc; do objs+=("$here/build/$f.o"); done
"$NVCC" "${ARCH[@]}" -shared --cudart static -o "$here/../qwen3_tts_b200/libq3tts_b200_$tag.so" "${objs[@]}" "$here/build/frame_ll_$tag.o"
echo "built libq3tts_b200_$tag.so"
