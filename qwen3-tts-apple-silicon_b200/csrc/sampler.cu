// On-device sampler: repetition penalty, min_new_tokens, suppress range, temperature, top-k, top-p,
// argmax / categorical draw -- one CTA per sequence, no host round trip per token.
//
// Replaces mx.argmax / mx.random.categorical plus the host-side logits processors the reference stack
// runs once per sampled token (one device sync per token).  Semantics follow the HF
// processors: logits_process.py:302 (repetition penalty over the SET of generated
// ids), :164 (min_new_tokens), :1865 (suppress), :236 (temperature), :536 (top-k, ties at the threshold
// kept), :469 (top-p on the ascending sort, min_tokens_to_keep = 1), then argmax (lowest index wins)
// or an inverse-CDF draw in index order.
#include "common.cuh"
#include "../../include/q3tts_b200.h"

namespace q3t {

constexpr int SAMPLE_THREADS = 1024;
constexpr int SAMPLE_MAXV = 4096;

struct SampleParams {
    const float* logits; int V; long long logits_stride;
    q3t_sampling sp;
    unsigned int* seen;
    const int* step;
    int rng_stream;
    const float* uniforms;
    int* out; long long out_stride; long long fo_stride, fo_step_stride;
    const int* forced; int* own;
    int* done;
};

__device__ __forceinline__ uint32_t ordered_key(float v) {
    const uint32_t u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ float hash_uniform(unsigned long long seed, int step, int stream, int b) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(step * 1315423911u + stream * 2654435761u + b * 97u + 1u);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (float)(z >> 40) * (1.0f / 16777216.0f);
}

__global__ void __launch_bounds__(SAMPLE_THREADS) sample_kernel(const SampleParams p) {
    __shared__ float sc[SAMPLE_MAXV];
    __shared__ float pe[SAMPLE_MAXV];
    __shared__ unsigned short cand[SAMPLE_MAXV];
    __shared__ unsigned int hist[256];
    __shared__ float redf[32];
    __shared__ int redi[32];
    __shared__ int sh_i[4];
    __shared__ float sh_f[2];

    pdl_wait();
    pdl_launch_dependents();
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int V = p.V;
    const int step = p.step ? *p.step : 0;
    const float* lg = p.logits + (size_t)b * p.logits_stride;
    const unsigned int* seen = p.seen ? p.seen + (size_t)b * ((V + 31) / 32) : nullptr;
    const q3t_sampling sp = p.sp;

    // ---- A: load + penalty + masks -----------------------------------------------------------
    for (int i = tid; i < V; i += SAMPLE_THREADS) {
        float v = lg[i];
        if (seen && sp.repetition_penalty != 1.0f && ((seen[i >> 5] >> (i & 31)) & 1u))
            v = v < 0.f ? v * sp.repetition_penalty : v / sp.repetition_penalty;
        if (i == sp.eos_id && step < sp.min_new_tokens) v = -INFINITY;
        if (i >= sp.suppress_lo && i < sp.suppress_hi && i != sp.eos_id) v = -INFINITY;
        if (sp.do_sample && sp.temperature != 1.0f) v = v / sp.temperature;
        sc[i] = v;
    }
    __syncthreads();

    int choice = 0;
    if (!sp.do_sample) {
        // ---- greedy: max value, lowest index on ties --------------------------------------------
        float bv = -INFINITY; int bi = 0x7fffffff;
        for (int i = tid; i < V; i += SAMPLE_THREADS) {
            const float v = sc[i];
            if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { redf[wid] = bv; redi[wid] = bi; }
        __syncthreads();
        if (wid == 0) {
            bv = redf[lane]; bi = redi[lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (lane == 0) sh_i[0] = (bi == 0x7fffffff) ? 0 : bi;
        }
        __syncthreads();
        choice = sh_i[0];
    } else {
        // ---- C: top-k threshold by 4-pass radix select on order-preserving keys ------------------
        if (sp.top_k > 0 && sp.top_k < V) {
            uint32_t prefix = 0, mask = 0;
            int k = sp.top_k;
            for (int shift = 24; shift >= 0; shift -= 8) {
                if (tid < 256) hist[tid] = 0;
                __syncthreads();
                for (int i = tid; i < V; i += SAMPLE_THREADS) {
                    const uint32_t key = ordered_key(sc[i]);
                    if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
                }
                __syncthreads();
                if (tid == 0) {
                    int cum = 0, bin = 255;
                    for (; bin > 0; --bin) {
                        if (cum + (int)hist[bin] >= k) break;
                        cum += hist[bin];
                    }
                    sh_i[1] = bin; sh_i[2] = k - cum;
                }
                __syncthreads();
                prefix |= (uint32_t)sh_i[1] << shift;
                mask |= 0xffu << shift;
                k = sh_i[2];
                __syncthreads();
            }
            for (int i = tid; i < V; i += SAMPLE_THREADS)
                if (ordered_key(sc[i]) < prefix) sc[i] = -INFINITY;
            __syncthreads();
        }
        // ---- D: compact survivors in index order ----------------------------------------------------
        // each thread owns a contiguous run of ids so that the compaction preserves index order
        const int per = (V + SAMPLE_THREADS - 1) / SAMPLE_THREADS;
        const int i0 = tid * per, i1 = min(V, i0 + per);
        int cnt = 0;
        for (int i = i0; i < i1; ++i) cnt += (sc[i] > -INFINITY);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
        }
        if (lane == 31) redi[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int w = redi[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int n = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += n;
            }
            redi[lane] = wi - w;            // exclusive warp offsets
            if (lane == 31) sh_i[3] = wi;   // total survivors
        }
        __syncthreads();
        int off = redi[wid] + incl - cnt;
        for (int i = i0; i < i1; ++i)
            if (sc[i] > -INFINITY) cand[off++] = (unsigned short)i;
        __syncthreads();
        const int n = sh_i[3];
        // ---- E: softmax numerators ------------------------------------------------------------------
        float mx = -INFINITY;
        for (int j = tid; j < n; j += SAMPLE_THREADS) mx = fmaxf(mx, sc[cand[j]]);
        mx = block_max(mx, redf);
        float part = 0.f;
        for (int j = tid; j < n; j += SAMPLE_THREADS) { const float e = expf(sc[cand[j]] - mx); pe[j] = e; part += e; }
        float total = block_sum(part, redf);
        __syncthreads();
        // ---- F: top-p on the ascending stable sort -------------------------------------------------
        if (sp.top_p < 1.0f && n > 1) {
            const float lim = 1.0f - sp.top_p;
            float keep_e[(SAMPLE_MAXV + SAMPLE_THREADS - 1) / SAMPLE_THREADS];
            int q = 0;
            for (int j = tid; j < n; j += SAMPLE_THREADS, ++q) {
                const float sj = sc[cand[j]];
                float cum = 0.f; bool is_last = true;
                for (int j2 = 0; j2 < n; ++j2) {
                    const float s2 = sc[cand[j2]];
                    const bool before = (s2 < sj) || (s2 == sj && j2 <= j);
                    cum += before ? pe[j2] : 0.f;
                    is_last = is_last && before;
                }
                keep_e[q] = (!is_last && (cum / total) <= lim) ? 0.f : pe[j];
            }
            __syncthreads();
            q = 0; part = 0.f;
            for (int j = tid; j < n; j += SAMPLE_THREADS, ++q) { pe[j] = keep_e[q]; part += keep_e[q]; }
            total = block_sum(part, redf);
            __syncthreads();
        }
        // ---- G: inverse-CDF draw in index order -------------------------------------------------------
        if (tid == 0) {
            const float u = p.uniforms ? p.uniforms[b] : hash_uniform(sp.seed, step, p.rng_stream, b);
            const float target = u * total;
            float cum = 0.f; int pick = -1, last_nz = 0;
            for (int j = 0; j < n; ++j) {
                if (pe[j] > 0.f) last_nz = j;
                cum += pe[j];
                if (cum > target && pe[j] > 0.f) { pick = j; break; }
            }
            if (pick < 0) pick = last_nz;
            sh_i[0] = n > 0 ? cand[pick] : 0;
        }
        __syncthreads();
        choice = sh_i[0];
    }

    if (tid == 0) {
        const long long o = (long long)b * p.fo_stride + (long long)step * p.fo_step_stride;
        if (p.own) p.own[o] = choice;
        if (p.forced) choice = p.forced[o];
        p.out[(long long)b * p.out_stride] = choice;
        if (p.done && choice == sp.eos_id) p.done[b] = 1;
        if (p.seen) atomicOr(p.seen + (size_t)b * ((V + 31) / 32) + (choice >> 5), 1u << (choice & 31));
    }
}

int launch_sample(const q3t_sample_args* a, cudaStream_t stream) {
    Q3T_REQUIRE(a->V > 0 && a->V <= SAMPLE_MAXV, "sample: vocabulary larger than the shared-memory buffer (4096)");
    SampleParams p;
    p.logits = a->logits; p.V = a->V; p.logits_stride = a->logits_stride; p.sp = a->sp; p.seen = a->seen;
    p.step = a->step; p.rng_stream = a->rng_stream; p.uniforms = a->uniforms; p.out = a->out;
    p.out_stride = a->out_stride; p.fo_stride = a->fo_stride; p.fo_step_stride = a->fo_step_stride;
    p.forced = a->forced; p.own = a->own;
    p.done = a->done;
    launch_pdl(sample_kernel, dim3(a->B), dim3(SAMPLE_THREADS), 0, stream, p);
    Q3T_CHECK_LAUNCH("sample");
    return 0;
}

}  // namespace q3t

extern "C" int q3t_sample(const q3t_sample_args* a, void* stream) {
    return q3t::launch_sample(a, (cudaStream_t)stream);
}
