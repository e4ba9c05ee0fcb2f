// Speech-tokenizer decoder operators (codes -> 24 kHz waveform), time-major fp32 activations.
//
// Replaces mx.take / add (RVQ), mx.conv1d, mx.conv_transpose1d, LayerNorm/GELU/SnakeBeta elementwise ops and
// the masked sdpa of the codec transformer in the reference stack (cousin arithmetic: transformers
// qwen3_omni_moe/modeling_qwen3_omni_moe.py:3283-3790, mimi/modeling_mimi.py:1176-1349).
//
// All convolutions are expressed as ONE causal tap-GEMM: rows = time steps, K = Cin per tap, N = Cout
// (x upsampling phases for transposed convs, whose output [T, s*Cout] IS the time-major [T*s, Cout] tensor).
// Round-1 implementation runs the contraction on the FP32 pipe with fused bias / LayerScale / residual /
// SnakeBeta / GELU / SwiGLU epilogues; the tcgen05 implicit-GEMM version replaces `tapgemm_kernel` only.
#include <stdlib.h>
#include <cuda_fp16.h>
#include "common.cuh"
#include "../../include/q3tts_b200.h"

namespace q3t {

// ---------------------------------------------------------------------------------- RVQ gather + sum
struct RvqTables { const float* t[32]; };

// A code outside [0, codebook_size) - the talker's EOS / control ids in the rows a finished sequence keeps producing in a
// lock-step batch - contributes a zero vector instead of reading past the table (the caller trims those frames).
__global__ void __launch_bounds__(128) rvq_gather_sum_kernel(const int* __restrict__ codes, RvqTables tabs, int G, int T,
                                                             int g_lo, int g_hi, int dim, int size, float* __restrict__ out) {
    const int bt = blockIdx.x, b = bt / T, t = bt % T;
    for (int d = threadIdx.x; d < dim; d += blockDim.x) {
        float acc = 0.f;
        for (int g = g_lo; g < g_hi; ++g) {           // left-to-right fp32 sum == the reference order (bit-exact)
            const int c = codes[((size_t)b * G + g) * T + t];
            const float v = (unsigned)c < (unsigned)size ? tabs.t[g][(size_t)c * dim + d] : 0.f;
            acc = (g == g_lo) ? v : acc + v;
        }
        out[(size_t)bt * dim + d] = acc;
    }
}

// ---------------------------------------------------------------------------------- tap-GEMM
constexpr int TG_BM = 64, TG_BN = 64, TG_BK = 16, TG_THREADS = 256;

struct TapGemmParams {
    const float* A; int B, T_in, Cin;
    const float* W; const float* bias;
    int taps; int shift[8]; int N; int Cout; int rows;   // rows per batch item
    const float* scale; const float* resid; float* out_raw; float* out_act; int act;
    const float* act_a; const float* act_b;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }

__global__ void __launch_bounds__(TG_THREADS) tapgemm_kernel(const TapGemmParams p) {
    __shared__ float As[TG_BK][TG_BM + 4];
    __shared__ float Ws[TG_BK][TG_BN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * TG_BM, n0 = blockIdx.y * TG_BN;
    const long long Mtot = (long long)p.B * p.rows;
    const int tx = tid & 15, ty = tid >> 4;            // 16 x 16 threads, 4x4 outputs each
    const int lr = tid >> 2, lk = (tid & 3) * 4;       // loader: row 0..63, k offset 0,4,8,12
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const long long gm = m0 + lr;
    const int lb = gm < Mtot ? (int)(gm / p.rows) : 0, lt = gm < Mtot ? (int)(gm % p.rows) : 0;
    const int gn = n0 + lr;
    for (int tap = 0; tap < p.taps; ++tap) {
        const int ts = lt + p.shift[tap];
        const bool arow_ok = gm < Mtot && ts >= 0 && ts < p.T_in;
        const float* arow = p.A + ((size_t)lb * p.T_in + (arow_ok ? ts : 0)) * p.Cin;
        const float* wrow = p.W + ((size_t)tap * p.N + (gn < p.N ? gn : 0)) * p.Cin;
        for (int k0 = 0; k0 < p.Cin; k0 += TG_BK) {
            float4 av = make_float4(0.f, 0.f, 0.f, 0.f), wv = av;
            if (arow_ok && k0 + lk < p.Cin) av = *reinterpret_cast<const float4*>(arow + k0 + lk);
            if (gn < p.N && k0 + lk < p.Cin) wv = *reinterpret_cast<const float4*>(wrow + k0 + lk);
            __syncthreads();
            As[lk + 0][lr] = av.x; As[lk + 1][lr] = av.y; As[lk + 2][lr] = av.z; As[lk + 3][lr] = av.w;
            Ws[lk + 0][lr] = wv.x; Ws[lk + 1][lr] = wv.y; Ws[lk + 2][lr] = wv.z; Ws[lk + 3][lr] = wv.w;
            __syncthreads();
#pragma unroll
            for (int k = 0; k < TG_BK; ++k) {
                const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
                const float4 w4 = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
                const float a[4] = {a4.x, a4.y, a4.z, a4.w}, w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
            }
        }
    }
    // ---- epilogue
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long m = m0 + ty * 4 + i;
        if (m >= Mtot) continue;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            float x = acc[i][j];
            if (n < p.N) {
                const int c = n % p.Cout;
                if (p.bias) x += p.bias[c];
                if (p.scale) x *= p.scale[c];
                if (p.resid) x += p.resid[m * p.N + n];
                if (p.out_raw) p.out_raw[m * p.N + n] = x;
            }
            v[j] = x;
        }
        if (!p.out_act) continue;
        if (p.act == Q3T_ACT_SWIGLU_PAIR) {
#pragma unroll
            for (int j = 0; j < 4; j += 2) {
                const int n = n0 + tx * 4 + j;
                if (n + 1 < p.N) p.out_act[m * (p.N / 2) + (n >> 1)] = silu_f(v[j]) * v[j + 1];
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + tx * 4 + j;
                if (n >= p.N) continue;
                const int c = n % p.Cout;
                float x = v[j];
                if (p.act == Q3T_ACT_SNAKE) { const float sn = sinf(x * p.act_a[c]); x = x + p.act_b[c] * (sn * sn); }
                else if (p.act == Q3T_ACT_GELU) x = gelu_erf(x);
                else if (p.act == Q3T_ACT_SILU) x = silu_f(x);
                else if (p.act >= Q3T_ACT_ELU) x = act_simple(x, p.act);
                p.out_act[m * p.N + n] = x;
            }
        }
    }
}

// ---------------------------------------------------------------------------------- ConvNeXt front half
__global__ void __launch_bounds__(256) dwconv_ln_kernel(const float* __restrict__ x, const float* __restrict__ dw_w,
                                                        const float* __restrict__ dw_b, const float* __restrict__ ln_w,
                                                        const float* __restrict__ ln_b, float eps, int T, int C, int ks,
                                                        float* __restrict__ out) {
    extern __shared__ float vals[];   // [C]
    __shared__ float red[32];
    const int bt = blockIdx.x, b = bt / T, t = bt % T;
    float s = 0.f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float a = dw_b[c];
        for (int j = 0; j < ks; ++j) {
            const int ts = t - (ks - 1) + j;
            if (ts >= 0) a = fmaf(dw_w[c * ks + j], x[((size_t)b * T + ts) * C + c], a);
        }
        vals[c] = a; s += a;
    }
    const float mean = block_sum(s, red) / (float)C;
    float q = 0.f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) { const float d = vals[c] - mean; q += d * d; }
    const float rstd = rsqrtf(block_sum(q, red) / (float)C + eps);
    for (int c = threadIdx.x; c < C; c += blockDim.x)
        out[(size_t)bt * C + c] = (vals[c] - mean) * rstd * ln_w[c] + ln_b[c];
}

// ---------------------------------------------------------------------------------- codec transformer attention
__global__ void __launch_bounds__(256) rope_qk_kernel(float* qkv, const float* __restrict__ inv_freq, int T, int H, int D,
                                                      long long total) {
    // one thread per (b, t, which in {q,k}, head, i < D/2)
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int half = D / 2;
    const int i = (int)(idx % half);
    long long r = idx / half;
    const int h = (int)(r % H); r /= H;
    const int which = (int)(r % 2); r /= 2;
    const int t = (int)(r % T);
    const long long bt = r;   // b*T + t
    float* v = qkv + (bt * 3 + which) * (long long)H * D + (long long)h * D;
    float sn, cs;
    sincosf((float)t * inv_freq[i], &sn, &cs);
    const float a = v[i], c = v[i + half];
    v[i] = a * cs - c * sn;
    v[i + half] = c * cs + a * sn;
}

template <int D>
__global__ void __launch_bounds__(128) window_attn_kernel(const float* __restrict__ qkv, int T, int H, int window,
                                                          float* __restrict__ out) {
    // one warp per (b, t, h); lane owns D/32 dims
    constexpr int E = D / 32;
    const int warp_global = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int h = warp_global % H;
    const long long bt = warp_global / H;
    const int t = (int)(bt % T);
    const long long b = bt / T;
    const size_t rowstride = (size_t)3 * H * D;
    const float* q = qkv + bt * rowstride + (size_t)h * D;
    float qr[E], acc[E];
    const float sc = rsqrtf((float)D);
#pragma unroll
    for (int e = 0; e < E; ++e) { qr[e] = q[lane * E + e] * sc; acc[e] = 0.f; }
    float m = -INFINITY, l = 0.f;
    const int j0 = max(0, t - window + 1);
    for (int j = j0; j <= t; ++j) {
        const float* kr = qkv + (b * T + j) * rowstride + (size_t)(H + h) * D;
        const float* vr = kr + (size_t)H * D;
        float s = 0.f;
#pragma unroll
        for (int e = 0; e < E; ++e) s = fmaf(qr[e], kr[lane * E + e], s);
        s = warp_sum(s);
        const float mn = fmaxf(m, s), corr = __expf(m - mn), pj = __expf(s - mn);
        l = l * corr + pj;
#pragma unroll
        for (int e = 0; e < E; ++e) acc[e] = fmaf(pj, vr[lane * E + e], acc[e] * corr);
        m = mn;
    }
    float* o = out + bt * (size_t)H * D + (size_t)h * D;
#pragma unroll
    for (int e = 0; e < E; ++e) o[lane * E + e] = acc[e] / l;
}

__global__ void snake_kernel(const float* __restrict__ x, const float* __restrict__ a, const float* __restrict__ b,
                             long long n, int C, float* __restrict__ y) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = (int)(i % C);
    const float v = x[i], sn = sinf(v * a[c]);
    y[i] = v + b[c] * (sn * sn);
}

__global__ void clamp_pcm_kernel(const float* __restrict__ x, long long n, float* y, int16_t* pcm) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = fminf(1.f, fmaxf(-1.f, x[i]));
    if (y) y[i] = v;
    if (pcm) pcm[i] = (int16_t)__float2int_rn(v * 32767.f);
}

// ---- output convolution of the vocoder: causal conv1d C -> 1 (k taps, dilation 1) + clamp(-1, 1) --------------------------
// 672 MACs per sample against 384 bytes of input per time step: HBM-bound, not GEMM-shaped (one output channel).  A CTA
// stages CO_TT + taps - 1 activation rows in shared memory (row stride C + 1 words: thread t reads row t + j, so the 32 lanes
// of a warp hit 32 different banks) and every thread produces one sample; the weights are broadcast reads.
constexpr int CO_TT = 256;

template <bool F16>      // F16: act is fp16 (written by a tap-GEMM with act_f16), staged to fp32 in shared memory
__global__ void __launch_bounds__(CO_TT) conv_out_clamp_kernel(const void* __restrict__ act_v, int T, int C, const float* __restrict__ W,
                                                                const float* __restrict__ bias, int taps, float* __restrict__ wav) {
    extern __shared__ float co_s[];
    const int ld = C + 1, nrow = CO_TT + taps - 1;
    float* w_s = co_s + (size_t)nrow * ld;
    const int b = blockIdx.y, t0 = blockIdx.x * CO_TT, tid = threadIdx.x;
    const int c4n = C >> 2;
    for (int i = tid; i < nrow * c4n; i += CO_TT) {
        const int r = i / c4n, c4 = i - r * c4n, t = t0 - (taps - 1) + r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);                    // rows before 0 are the causal left padding
        if (t >= 0 && t < T) {
            if (F16) {
                const uint2 h = __ldcs(reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(act_v) + ((size_t)b * T + t) * C) + c4);
                const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&h.x)), hi = __half22float2(*reinterpret_cast<const __half2*>(&h.y));
                v = make_float4(lo.x, lo.y, hi.x, hi.y);
            } else {
                v = __ldcs(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(act_v) + ((size_t)b * T + t) * C) + c4);
            }
        }
        float* d = co_s + (size_t)r * ld + 4 * c4;
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
    for (int i = tid; i < taps * C; i += CO_TT) w_s[i] = W[i];
    __syncthreads();
    const int t = t0 + tid;
    if (t >= T) return;
    float acc0 = bias ? bias[0] : 0.f, acc1 = 0.f;
    for (int j = 0; j < taps; ++j) {
        const float* a = co_s + (size_t)(tid + j) * ld;
        const float* w = w_s + j * C;
#pragma unroll 8
        for (int c = 0; c < C; c += 2) { acc0 = fmaf(a[c], w[c], acc0); acc1 = fmaf(a[c + 1], w[c + 1], acc1); }
    }
    wav[(size_t)b * T + t] = fminf(1.f, fmaxf(-1.f, acc0 + acc1));
}

}  // namespace q3t

using namespace q3t;

extern "C" int q3t_rvq_gather_sum(const int* codes, const float* const* tables_host, int B, int G, int T, int g_lo,
                                  int g_hi, int dim, int codebook_size, float* out, void* stream) {
    Q3T_REQUIRE(G <= 32 && g_lo >= 0 && g_hi <= G && g_lo < g_hi && codebook_size > 0, "rvq_gather_sum: bad group range");
    RvqTables tabs;
    for (int g = 0; g < 32; ++g) tabs.t[g] = g < G ? tables_host[g] : nullptr;
    if (B * T == 0) return 0;
    rvq_gather_sum_kernel<<<B * T, 128, 0, (cudaStream_t)stream>>>(codes, tabs, G, T, g_lo, g_hi, dim, codebook_size, out);
    Q3T_CHECK_LAUNCH("rvq_gather_sum");
    return 0;
}

namespace q3t {
int launch_tapgemm_tc(const q3t_tapgemm_args* a, cudaStream_t stream);
bool tapgemm_tc_eligible(const q3t_tapgemm_args* a);
// tensor-core implicit GEMM (tcgen05) for every eligible layer; Q3T_CODEC_TC=0 forces the FP32-pipe kernel
static bool tc_enabled() {
    static int use_tc = -1;
    if (use_tc < 0) { const char* e = getenv("Q3T_CODEC_TC"); use_tc = (e && e[0] == '0') ? 0 : 1; }
    return use_tc != 0;
}
// which kernel served q3t_tapgemm: [0] tcgen05 tap-GEMM, [1] FP32-pipe kernel although Cin % 32 == 0 (too few rows / odd N /
// Q3T_CODEC_TC=0), [2] FP32-pipe kernel because Cin % 32 != 0.  Tests assert [1] == 0 at the BASELINE shapes.
static unsigned long long g_tap_stats[3] = {0, 0, 0};
}

extern "C" void q3t_tapgemm_stats(unsigned long long* out3, int reset) {
    for (int i = 0; i < 3; ++i) { if (out3) out3[i] = q3t::g_tap_stats[i]; if (reset) q3t::g_tap_stats[i] = 0; }
}

extern "C" int q3t_tapgemm(const q3t_tapgemm_args* a, void* stream) {
    Q3T_REQUIRE(a->taps >= 1 && a->taps <= 8, "tapgemm: taps in [1,8]");
    Q3T_REQUIRE(a->Cin % 4 == 0, "tapgemm: Cin % 4");
    Q3T_REQUIRE(a->act != Q3T_ACT_SWIGLU_PAIR || (a->up * a->Cout) % 2 == 0, "tapgemm: SWIGLU_PAIR needs even N");
    if (q3t::tc_enabled() && !a->force_fp32 && (long long)a->B * a->T_out_rows > 0) {
        const int rc = q3t::launch_tapgemm_tc(a, (cudaStream_t)stream);
        if (rc >= 0) { q3t::g_tap_stats[0]++; return rc; }
    }
    Q3T_REQUIRE(!a->a_f16 && !a->act_f16, "tapgemm: fp16 operands only exist on the tcgen05 path (q3t_tapgemm_tc_eligible)");
    q3t::g_tap_stats[(a->Cin % 32 == 0 && !a->force_fp32) ? 1 : 2]++;
    TapGemmParams p;
    p.A = a->A; p.B = a->B; p.T_in = a->T_in; p.Cin = a->Cin; p.W = a->W; p.bias = a->bias; p.taps = a->taps;
    for (int i = 0; i < 8; ++i) p.shift[i] = a->shift[i];
    p.N = a->up * a->Cout; p.Cout = a->Cout; p.rows = a->T_out_rows; p.scale = a->scale; p.resid = a->resid;
    p.out_raw = a->out_raw; p.out_act = a->out_act; p.act = a->act; p.act_a = a->act_a; p.act_b = a->act_b;
    const long long M = (long long)a->B * a->T_out_rows;
    if (M == 0) return 0;
    dim3 grid((unsigned)((M + TG_BM - 1) / TG_BM), (unsigned)((p.N + TG_BN - 1) / TG_BN));
    tapgemm_kernel<<<grid, TG_THREADS, 0, (cudaStream_t)stream>>>(p);
    Q3T_CHECK_LAUNCH("tapgemm");
    return 0;
}

extern "C" int q3t_dwconv_ln(const float* x, const float* dw_w, const float* dw_b, const float* ln_w, const float* ln_b,
                             float ln_eps, int B, int T, int C, int ksize, float* out, void* stream) {
    if (B * T == 0) return 0;
    dwconv_ln_kernel<<<B * T, 256, C * sizeof(float), (cudaStream_t)stream>>>(x, dw_w, dw_b, ln_w, ln_b, ln_eps, T, C,
                                                                             ksize, out);
    Q3T_CHECK_LAUNCH("dwconv_ln");
    return 0;
}

extern "C" int q3t_window_attn(float* qkv, const float* inv_freq, int B, int T, int H, int D, int window, float* out,
                               void* stream) {
    if (B * T == 0) return 0;
    const long long total = (long long)B * T * 2 * H * (D / 2);
    rope_qk_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(qkv, inv_freq, T, H, D, total);
    Q3T_CHECK_LAUNCH("rope_qk");
    const long long warps = (long long)B * T * H;
    Q3T_REQUIRE(warps % 4 == 0, "window_attn: B*T*H must be a multiple of 4");
    if (D == 64) window_attn_kernel<64><<<(unsigned)(warps / 4), 128, 0, (cudaStream_t)stream>>>(qkv, T, H, window, out);
    else if (D == 32) window_attn_kernel<32><<<(unsigned)(warps / 4), 128, 0, (cudaStream_t)stream>>>(qkv, T, H, window, out);
    else if (D == 128) window_attn_kernel<128><<<(unsigned)(warps / 4), 128, 0, (cudaStream_t)stream>>>(qkv, T, H, window, out);
    else Q3T_REQUIRE(false, "window_attn: head_dim must be 32, 64 or 128");
    Q3T_CHECK_LAUNCH("window_attn");
    return 0;
}

extern "C" int q3t_snake(const float* x, const float* a, const float* b, long long rows, int C, float* y, void* stream) {
    const long long n = rows * C;
    if (n == 0) return 0;
    snake_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, a, b, n, C, y);
    Q3T_CHECK_LAUNCH("snake");
    return 0;
}

static int conv_out_clamp_any(const void* act, bool f16, int B, int T, int C, const float* W, const float* bias, int taps, float* wav,
                              void* stream) {
    Q3T_REQUIRE(C % 4 == 0 && taps >= 1 && taps <= 16, "conv_out_clamp: C % 4, 1 <= taps <= 16");
    if (B * T == 0) return 0;
    const size_t smem = ((size_t)(CO_TT + taps - 1) * (C + 1) + (size_t)taps * C) * sizeof(float);
    Q3T_REQUIRE(smem <= 227 * 1024, "conv_out_clamp: too many channels for one CTA");
    static size_t smem_set[2] = {0, 0};
    if (smem > smem_set[f16]) {
        if (f16) cudaFuncSetAttribute(conv_out_clamp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        else cudaFuncSetAttribute(conv_out_clamp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        smem_set[f16] = smem;
    }
    const dim3 grid((unsigned)((T + CO_TT - 1) / CO_TT), (unsigned)B);
    if (f16) conv_out_clamp_kernel<true><<<grid, CO_TT, smem, (cudaStream_t)stream>>>(act, T, C, W, bias, taps, wav);
    else conv_out_clamp_kernel<false><<<grid, CO_TT, smem, (cudaStream_t)stream>>>(act, T, C, W, bias, taps, wav);
    Q3T_CHECK_LAUNCH("conv_out_clamp");
    return 0;
}
extern "C" int q3t_conv_out_clamp(const float* act, int B, int T, int C, const float* W, const float* bias, int taps, float* wav,
                                  void* stream) {
    return conv_out_clamp_any(act, false, B, T, C, W, bias, taps, wav, stream);
}
extern "C" int q3t_conv_out_clamp_h(const void* act_f16, int B, int T, int C, const float* W, const float* bias, int taps, float* wav,
                                    void* stream) {
    return conv_out_clamp_any(act_f16, true, B, T, C, W, bias, taps, wav, stream);
}
extern "C" int q3t_tapgemm_tc_eligible(const q3t_tapgemm_args* a) {
    return (q3t::tc_enabled() && (long long)a->B * a->T_out_rows > 0 && q3t::tapgemm_tc_eligible(a)) ? 1 : 0;
}

extern "C" int q3t_clamp_pcm16(const float* x, long long n, float* y, int16_t* pcm, void* stream) {
    if (n == 0) return 0;
    clamp_pcm_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, n, y, pcm);
    Q3T_CHECK_LAUNCH("clamp_pcm16");
    return 0;
}
