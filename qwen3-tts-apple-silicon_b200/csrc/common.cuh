// Shared device/host helpers for the sm_100a kernels of the Qwen3-TTS hot path.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace q3t {

// ---- error plumbing (no exception crosses the C ABI) -----------------------------------------
extern thread_local char g_err[512];
extern unsigned long long g_launches;
extern int g_use_pdl;   // programmatic dependent launch between consecutive kernels (env Q3T_PDL=0 disables)

#define Q3T_CHECK_LAUNCH(name)                                                        \
    do {                                                                              \
        cudaError_t e__ = cudaGetLastError();                                         \
        if (e__ != cudaSuccess) {                                                     \
            snprintf(q3t::g_err, sizeof(q3t::g_err), "%s: %s", name, cudaGetErrorString(e__)); \
            return 1;                                                                 \
        }                                                                             \
        ++q3t::g_launches;                                                            \
    } while (0)

#define Q3T_REQUIRE(cond, msg)                                                        \
    do {                                                                              \
        if (!(cond)) { snprintf(q3t::g_err, sizeof(q3t::g_err), "%s:%d: %s", __FILE__, __LINE__, msg); return 2; } \
    } while (0)

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------
// Every kernel of the decode chain (a) lets its successor start early and (b) waits for its predecessor only
// right before it touches activations, so the successor's weight prefetch overlaps the predecessor's tail.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = g_use_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ---- warp / block reductions --------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum; `red` is a shared float[32]; result broadcast to every thread.  Deterministic.
__device__ __forceinline__ float block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    float r = (lane < nw) ? red[lane] : 0.f;
    r = warp_sum(r);
    return r;
}
__device__ __forceinline__ float block_max(float v, float* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    float r = (lane < nw) ? red[lane] : -INFINITY;
    r = warp_max(r);
    return r;
}

// ---- streaming loads: weights are read exactly once per step -> do not allocate in L1 --------------
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ float bf16lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

__device__ __forceinline__ float silu_f(float x) { return x / (1.f + expf(-x)); }
// pointwise activations of the reference-clip encoders (Q3T_ACT_ELU .. Q3T_ACT_TANH); ELU with alpha = 1 (torch default)
__device__ __forceinline__ float act_simple(float x, int act) {
    if (act == 5) return x > 0.f ? x : expm1f(x);
    if (act == 6) return fmaxf(x, 0.f);
    if (act == 7) return 1.f / (1.f + expf(-x));
    if (act == 8) return tanhf(x);
    return x;
}

}  // namespace q3t
