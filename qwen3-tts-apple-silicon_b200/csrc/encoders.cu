// Reference-clip side of voice cloning (reference call site sessions/clone.py:218-224): the operators of
// the speech-tokenizer ENCODER (Mimi: conv stem, transformer, residual VQ encode) and of the ECAPA-TDNN speaker encoder that are
// not convolutions / linears (those run on q3t_tapgemm with force_fp32: a nearest-codebook search follows, so the operands are
// not rounded to TF32).  Runs once per reference clip (3 s of audio), so these kernels are sized for clarity, not for a roofline:
//   rvq_encode_kernel      mimi:1197-1203 (nearest codebook entry), :1262-1281 (residual levels) - integer results
//   time_stats_kernel      (attention-)weighted mean / std over time per channel   (qwen2_5_omni:2654-2679)
//   softmax_time_kernel    softmax over the time axis per channel                   (qwen2_5_omni:2674)
//   eltwise_kernel         add / tanh / channel gate + residual                     (Res2Net sum, squeeze-excitation)
//   mel_kernel             |STFT| -> slaney mel filter bank -> log(clamp)           (BigVGAN-style front end)
//   layernorm_kernel       LayerNorm over channels with bias                        (mimi:926-995)
// Cousin arithmetic: transformers mimi/modeling_mimi.py, qwen2_5_omni/modeling_qwen2_5_omni.py:2499-2790.
#include "common.cuh"
#include "../../include/q3tts_b200.h"

namespace q3t {

// ---- residual VQ encode: one CTA per vector walks all levels (levels are sequential per vector, vectors independent) --------
struct RvqEncTables { const float* t[32]; };
constexpr int RE_THREADS = 256;

__global__ void __launch_bounds__(RE_THREADS) rvq_encode_kernel(const float* __restrict__ x, RvqEncTables tabs, int n_levels, int size,
                                                                 int dim, long long idx_level_stride, int* __restrict__ idx_out,
                                                                 float* __restrict__ resid_out) {
    extern __shared__ float re_sh[];                 // residual [dim]
    __shared__ float red_v[RE_THREADS / 32];
    __shared__ int red_i[RE_THREADS / 32];
    __shared__ int best_sh;
    const int v = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int d = tid; d < dim; d += RE_THREADS) re_sh[d] = x[(size_t)v * dim + d];
    __syncthreads();
    for (int lv = 0; lv < n_levels; ++lv) {
        const float* tab = tabs.t[lv];
        // squared Euclidean distance in the direct form sum (x - c)^2 (cdist's argmin; ties -> lowest index like torch.argmin)
        float bv = INFINITY; int bi = 0x7fffffff;
        for (int e = warp; e < size; e += RE_THREADS / 32) {
            const float* c = tab + (size_t)e * dim;
            float acc = 0.f;
            for (int d = lane; d < dim; d += 32) { const float df = re_sh[d] - c[d]; acc = fmaf(df, df, acc); }
            acc = warp_sum(acc);
            if (acc < bv) { bv = acc; bi = e; }      // e ascends inside a warp: strict < keeps the lowest index
        }
        if (lane == 0) { red_v[warp] = bv; red_i[warp] = bi; }
        __syncthreads();
        if (tid == 0) {
            float b = red_v[0]; int i = red_i[0];
            for (int w = 1; w < RE_THREADS / 32; ++w)
                if (red_v[w] < b || (red_v[w] == b && red_i[w] < i)) { b = red_v[w]; i = red_i[w]; }
            best_sh = i;
            idx_out[(size_t)lv * idx_level_stride + v] = i;
        }
        __syncthreads();
        const float* c = tab + (size_t)best_sh * dim;
        for (int d = tid; d < dim; d += RE_THREADS) re_sh[d] -= c[d];
        __syncthreads();
    }
    if (resid_out)
        for (int d = tid; d < dim; d += RE_THREADS) resid_out[(size_t)v * dim + d] = re_sh[d];
}

// ---- weighted mean / std over time: x [B, T, C] (+ weights w [B, T, C] or uniform 1/T) -> mean [B, C], std [B, C] -------------
__global__ void time_stats_kernel(const float* __restrict__ x, const float* __restrict__ w, int T, int C, float eps,
                                  float* __restrict__ mean, float* __restrict__ stdv) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (c >= C) return;
    const float* xb = x + (size_t)b * T * C + c;
    const float* wb = w ? w + (size_t)b * T * C + c : nullptr;
    const float u = 1.f / (float)T;
    float m = 0.f;
    for (int t = 0; t < T; ++t) m = fmaf(wb ? wb[(size_t)t * C] : u, xb[(size_t)t * C], m);
    float s = 0.f;
    for (int t = 0; t < T; ++t) { const float d = xb[(size_t)t * C] - m; s = fmaf(wb ? wb[(size_t)t * C] : u, d * d, s); }
    mean[(size_t)b * C + c] = m;
    stdv[(size_t)b * C + c] = sqrtf(fmaxf(s, eps));
}

__global__ void softmax_time_kernel(const float* __restrict__ x, int T, int C, float* __restrict__ y) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (c >= C) return;
    const float* xb = x + (size_t)b * T * C + c;
    float* yb = y + (size_t)b * T * C + c;
    float mx = -INFINITY;
    for (int t = 0; t < T; ++t) mx = fmaxf(mx, xb[(size_t)t * C]);
    float l = 0.f;
    for (int t = 0; t < T; ++t) l += expf(xb[(size_t)t * C] - mx);
    const float il = 1.f / l;
    for (int t = 0; t < T; ++t) yb[(size_t)t * C] = expf(xb[(size_t)t * C] - mx) * il;
}

// ---- elementwise: op 0: out = a + b;  op 1: out = tanh(a);  op 2: out = a * gate[b, c] + r   (a, r [B, T, C]; gate [B, C]) ------
__global__ void eltwise_kernel(int op, const float* __restrict__ a, const float* __restrict__ b2, const float* __restrict__ r,
                               long long n, int C, long long per_item, float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = a[i];
    if (op == 0) v += b2[i];
    else if (op == 1) v = tanhf(v);
    else { const long long item = i / per_item; v = fmaf(v, b2[item * C + (i % C)], r ? r[i] : 0.f); }
    out[i] = v;
}

// ---- mel front end: spec [R, ld] holds re (cols 0..NF-1) and im (cols NF..2NF-1) of every frame ---------------------------------
__global__ void __launch_bounds__(128) mel_kernel(const float* __restrict__ spec, int ld, int NF, const float* __restrict__ fb, int n_mels,
                                                  float* __restrict__ out) {
    extern __shared__ float mag[];                   // [NF]
    const int r = blockIdx.x;
    const float* s = spec + (size_t)r * ld;
    for (int f = threadIdx.x; f < NF; f += blockDim.x) mag[f] = sqrtf(s[f] * s[f] + s[NF + f] * s[NF + f] + 1e-9f);
    __syncthreads();
    for (int m = threadIdx.x; m < n_mels; m += blockDim.x) {
        float acc = 0.f;
        for (int f = 0; f < NF; ++f) acc = fmaf(fb[(size_t)f * n_mels + m], mag[f], acc);
        out[(size_t)r * n_mels + m] = logf(fmaxf(acc, 1e-5f));
    }
}

// ---- LayerNorm over the channel axis (weight + bias), one CTA per row -------------------------------------------------------
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                                        int C, float eps, float* __restrict__ y) {
    __shared__ float red[32];
    const float* xr = x + (size_t)blockIdx.x * C;
    float s = 0.f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) s += xr[c];
    const float mean = block_sum(s, red) / (float)C;
    float q = 0.f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) { const float d = xr[c] - mean; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(block_sum(q, red) / (float)C + eps);
    for (int c = threadIdx.x; c < C; c += blockDim.x) y[(size_t)blockIdx.x * C + c] = (xr[c] - mean) * rstd * w[c] + b[c];
}

}  // namespace q3t

extern "C" int q3t_rvq_encode(const float* x, const float* const* tables_host, int n_vectors, int n_levels, int codebook_size,
                              int dim, long long idx_level_stride, int* idx_out, float* resid_out, void* stream) {
    Q3T_REQUIRE(n_levels >= 1 && n_levels <= 32 && dim >= 1 && dim <= 4096, "rvq_encode: 1..32 levels, dim <= 4096");
    if (n_vectors == 0) return 0;
    q3t::RvqEncTables tabs;
    for (int g = 0; g < 32; ++g) tabs.t[g] = g < n_levels ? tables_host[g] : nullptr;
    q3t::rvq_encode_kernel<<<n_vectors, q3t::RE_THREADS, (size_t)dim * sizeof(float), (cudaStream_t)stream>>>(
        x, tabs, n_levels, codebook_size, dim, idx_level_stride, idx_out, resid_out);
    Q3T_CHECK_LAUNCH("rvq_encode");
    return 0;
}

extern "C" int q3t_time_stats(const float* x, const float* w, int B, int T, int C, float eps, float* mean, float* stdv, void* stream) {
    if ((long long)B * T * C == 0) return 0;
    q3t::time_stats_kernel<<<dim3((unsigned)((C + 127) / 128), (unsigned)B), 128, 0, (cudaStream_t)stream>>>(x, w, T, C, eps, mean, stdv);
    Q3T_CHECK_LAUNCH("time_stats");
    return 0;
}

extern "C" int q3t_softmax_time(const float* x, int B, int T, int C, float* y, void* stream) {
    if ((long long)B * T * C == 0) return 0;
    q3t::softmax_time_kernel<<<dim3((unsigned)((C + 127) / 128), (unsigned)B), 128, 0, (cudaStream_t)stream>>>(x, T, C, y);
    Q3T_CHECK_LAUNCH("softmax_time");
    return 0;
}

extern "C" int q3t_eltwise(int op, const float* a, const float* b, const float* r, long long n, int C, long long per_item, float* out,
                           void* stream) {
    Q3T_REQUIRE(op >= 0 && op <= 2, "eltwise: op in {0 add, 1 tanh, 2 gate}");
    if (n == 0) return 0;
    q3t::eltwise_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(op, a, b, r, n, C, per_item, out);
    Q3T_CHECK_LAUNCH("eltwise");
    return 0;
}

extern "C" int q3t_mel(const float* spec, long long rows, int ld, int n_freq, const float* fb, int n_mels, float* out, void* stream) {
    Q3T_REQUIRE(2 * n_freq <= ld && n_freq <= 8192, "mel: spectrum row too short / too many bins");
    if (rows == 0) return 0;
    q3t::mel_kernel<<<(unsigned)rows, 128, (size_t)n_freq * sizeof(float), (cudaStream_t)stream>>>(spec, ld, n_freq, fb, n_mels, out);
    Q3T_CHECK_LAUNCH("mel");
    return 0;
}

extern "C" int q3t_layernorm(const float* x, const float* w, const float* b, long long rows, int C, float eps, float* y, void* stream) {
    if (rows == 0) return 0;
    q3t::layernorm_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(x, w, b, C, eps, y);
    Q3T_CHECK_LAUNCH("layernorm");
    return 0;
}
