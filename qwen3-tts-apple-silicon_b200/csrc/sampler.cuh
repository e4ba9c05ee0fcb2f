// On-device sampler core shared by the stand-alone sampler kernel (sampler.cu) and the persistent frame kernel
// (frame_ll.cu).  Semantics follow the HF processors: logits_process.py:302 (repetition penalty
// over the SET of generated ids), :164 (min_new_tokens), :1865 (suppress), :236 (temperature), :536 (top-k, ties at
// the threshold kept), :469 (top-p on the ascending sort, min_tokens_to_keep = 1), then argmax (lowest index wins) or
// an inverse-CDF draw in index order.
#pragma once
#include "common.cuh"
#include "../../include/q3tts_b200.h"

namespace q3t {

constexpr int SAMPLE_MAXV = 4096;

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

// scratch of one sampling CTA (shared or global memory; only the owning CTA touches it)
struct SampleScratch {
    float* sc;              // [SAMPLE_MAXV]
    float* pe;              // [SAMPLE_MAXV]
    unsigned short* cand;   // [SAMPLE_MAXV]
    unsigned int* hist;     // [256]
    float* redf;            // [32]
    int* redi;              // [32]
    int* sh_i;              // [4]
};

// NT threads (tid in [0, NT)) of one CTA; BAR() is a barrier over exactly those threads.  `sc` must already hold the
// processed scores (penalty, masks, temperature).  Returns the chosen id in every thread.
template <int NT, typename BarT>
__device__ __forceinline__ int sample_core(const SampleScratch& s, int V, const q3t_sampling& sp, float uniform, int tid,
                                           BarT BAR) {
    const int lane = tid & 31, wid = tid >> 5;
    constexpr int NW = NT / 32;
    float* sc = s.sc; float* pe = s.pe; unsigned short* cand = s.cand; unsigned int* hist = s.hist;
    float* redf = s.redf; int* redi = s.redi; int* sh_i = s.sh_i;
    auto bsum = [&](float v) {
        v = warp_sum(v);
        BAR();
        if (lane == 0) redf[wid] = v;
        BAR();
        float r = (lane < NW) ? redf[lane] : 0.f;
        return warp_sum(r);
    };
    auto bmax = [&](float v) {
        v = warp_max(v);
        BAR();
        if (lane == 0) redf[wid] = v;
        BAR();
        float r = (lane < NW) ? redf[lane] : -INFINITY;
        return warp_max(r);
    };
    if (!sp.do_sample) {
        float bv = -INFINITY; int bi = 0x7fffffff;
        for (int i = tid; i < V; i += NT) {
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
        BAR();
        if (wid == 0) {
            bv = lane < NW ? redf[lane] : -INFINITY; bi = lane < NW ? redi[lane] : 0x7fffffff;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (lane == 0) sh_i[0] = (bi == 0x7fffffff) ? 0 : bi;
        }
        BAR();
        const int choice = sh_i[0];
        BAR();
        return choice;
    }
    // ---- top-k threshold by 4-pass radix select on order-preserving keys ------------------
    if (sp.top_k > 0 && sp.top_k < V) {
        uint32_t prefix = 0, mask = 0;
        int k = sp.top_k;
        for (int shift = 24; shift >= 0; shift -= 8) {
            for (int i = tid; i < 256; i += NT) hist[i] = 0;
            BAR();
            for (int i = tid; i < V; i += NT) {
                const uint32_t key = ordered_key(sc[i]);
                if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
            }
            BAR();
            if (tid == 0) {
                int cum = 0, bin = 255;
                for (; bin > 0; --bin) {
                    if (cum + (int)hist[bin] >= k) break;
                    cum += hist[bin];
                }
                sh_i[1] = bin; sh_i[2] = k - cum;
            }
            BAR();
            prefix |= (uint32_t)sh_i[1] << shift;
            mask |= 0xffu << shift;
            k = sh_i[2];
            BAR();
        }
        for (int i = tid; i < V; i += NT)
            if (ordered_key(sc[i]) < prefix) sc[i] = -INFINITY;
        BAR();
    }
    // ---- compact survivors in index order: each thread owns a contiguous run of ids ------------
    const int per = (V + NT - 1) / NT;
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
    BAR();
    if (wid == 0) {
        int w = lane < NW ? redi[lane] : 0, wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += n;
        }
        if (lane < NW) redi[lane] = wi - w;   // exclusive warp offsets
        if (lane == 31) sh_i[3] = wi;         // total survivors
    }
    BAR();
    int off = redi[wid] + incl - cnt;
    for (int i = i0; i < i1; ++i)
        if (sc[i] > -INFINITY) cand[off++] = (unsigned short)i;
    BAR();
    const int n = sh_i[3];
    // ---- softmax numerators ------------------------------------------------------------------
    float mx = -INFINITY;
    for (int j = tid; j < n; j += NT) mx = fmaxf(mx, sc[cand[j]]);
    mx = bmax(mx);
    float part = 0.f;
    for (int j = tid; j < n; j += NT) { const float e = expf(sc[cand[j]] - mx); pe[j] = e; part += e; }
    float total = bsum(part);
    BAR();
    // ---- top-p on the ascending stable sort -------------------------------------------------
    if (sp.top_p < 1.0f && n > 1) {
        const float lim = 1.0f - sp.top_p;
        float keep_e[(SAMPLE_MAXV + NT - 1) / NT];
        int q = 0;
        for (int j = tid; j < n; j += NT, ++q) {
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
        BAR();
        q = 0; part = 0.f;
        for (int j = tid; j < n; j += NT, ++q) { pe[j] = keep_e[q]; part += keep_e[q]; }
        total = bsum(part);
        BAR();
    }
    // ---- inverse-CDF draw in index order -------------------------------------------------------
    if (tid == 0) {
        const float target = uniform * total;
        float cum = 0.f; int pick = -1, last_nz = 0;
        for (int j = 0; j < n; ++j) {
            if (pe[j] > 0.f) last_nz = j;
            cum += pe[j];
            if (cum > target && pe[j] > 0.f) { pick = j; break; }
        }
        if (pick < 0) pick = last_nz;
        sh_i[0] = n > 0 ? cand[pick] : 0;
    }
    BAR();
    const int choice = sh_i[0];
    BAR();
    return choice;
}

// score of logit i after the HF processors that precede the top-k/top-p filters
__device__ __forceinline__ float sample_score(float v, int i, const q3t_sampling& sp, const unsigned int* seen, int step) {
    if (seen && sp.repetition_penalty != 1.0f && ((seen[i >> 5] >> (i & 31)) & 1u))
        v = v < 0.f ? v * sp.repetition_penalty : v / sp.repetition_penalty;
    if (i == sp.eos_id && step < sp.min_new_tokens) v = -INFINITY;
    if (i >= sp.suppress_lo && i < sp.suppress_hi && i != sp.eos_id) v = -INFINITY;
    if (sp.do_sample && sp.temperature != 1.0f) v = v / sp.temperature;
    return v;
}

}  // namespace q3t
