#!/bin/bash
# Builds libq3tts_b200_<tag>.so with extra nvcc defines for w8_gemm_tc.cu (e.g. -DTC_TIMING for tools/gemm_stamps.py; load with Q3T_LIB=...).
set -euo pipefail
tag="$1"; shift
here="$(cd "$(dirname "$0")/../qwen3-tts-apple-silicon_b200/csrc" && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
ARCH=(-gencode arch=compute_100a,code=sm_100a)
"$NVCC" "${ARCH[@]}" -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -diag-suppress 177 "$@" -c "$here/w8_gemm_tc.cu" -o "$here/build/w8_gemm_tc_$tag.o"
objs=()
for f in glue w8_gemv attn_decode attn_prefill sampler engine codec tapgemm_tc encoders frame_ll; do objs+=("$here/build/$f.o"); done
"$NVCC" "${ARCH[@]}" -shared --cudart static -o "$here/../qwen3_tts_b200/libq3tts_b200_$tag.so" "${objs[@]}" "$here/build/w8_gemm_tc_$tag.o"
echo "built libq3tts_b200_$tag.so"
