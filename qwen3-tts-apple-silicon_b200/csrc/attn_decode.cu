// Paged-KV GQA decode attention with fused per-head q/k RMSNorm + RoPE + KV-page write (sm_100a).
//
// Replaces, for one new token per sequence: mx.fast.rms_norm (q_norm, k_norm), mx.fast.rope, the KV
// cache update and mx.fast.scaled_dot_product_attention of the reference stack
// (cousin arithmetic: transformers qwen3/modeling_qwen3.py:222-291).
//
// grid = (kv_head, split, batch); 128 threads.  A CTA owns one KV head (its `REP` query heads share every
// K/V byte it streams) and one contiguous slice of the context, so the KV pages are read exactly once
// per step.  Each half-warp streams two tokens of a page per iteration with 128-bit loads (16 lanes x 8 bf16 =
// one 256 B row), the next page's loads in flight; scores are reduced with 4 xor-shuffles; softmax is online in
// fp32.  The slices of one (kv head, sequence) are a thread-block cluster merged by slice 0 through distributed
// shared memory (more than 8 slices: L2 workspace + last CTA to arrive); fixed slice order => deterministic,
// no float atomics.
#include "common.cuh"
#include "../../include/q3tts_b200.h"

namespace q3t {

struct AttnParams {
    const float* qkv; const float* q_norm_w; const float* k_norm_w; float eps;
    const float* inv_freq;
    __nv_bfloat16* kv_pool;
    const int* block_tbl; int max_pages;
    const int* pos;
    float* out; __nv_bfloat16* out_bf16; float* work; int* counters;
    int B, H, Hkv, nsplit;
    int mode;                 // 0 fused decode, 1 write K/V of the row only, 2 attention only (K/V already in the cache)
    const int* seq_of_row;    // optional: block-table row of launch row b (prefill: many rows share one sequence)
    int cluster;              // the nsplit CTAs of one (kv head, sequence) were launched as one thread-block cluster
};

template <int EPL> struct KvVec;
template <> struct KvVec<8> { using T = uint4; };
template <> struct KvVec<4> { using T = uint2; };
template <> struct KvVec<2> { using T = uint32_t; };

template <int D, int REP>
__global__ void __launch_bounds__(128) attn_decode_kernel(const AttnParams p) {
    constexpr int E = D / 32;     // elements per lane in the prep stage
    constexpr int EPL = D / 16;   // elements per sub-lane in the streaming stage
    constexpr int PSTR = D + 2;   // partial record: acc[D], m, l
    __shared__ float q_s[REP][D];
    __shared__ float part_s[8][REP][PSTR];
    __shared__ float rec_s[REP][PSTR];   // this slice's merged record
    __shared__ int is_last;

    const int kvh = blockIdx.x, split = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int* btbl = p.block_tbl + (size_t)(p.seq_of_row ? p.seq_of_row[b] : b) * p.max_pages;
    const size_t page_elems = (size_t)2 * p.Hkv * Q3T_KV_PAGE * D;
    const size_t head_off = (size_t)kvh * Q3T_KV_PAGE * D;
    const size_t v_off = (size_t)p.Hkv * Q3T_KV_PAGE * D;
    pdl_launch_dependents();
    if (p.mode != 1) {   // before waiting for the QKV GEMV: pull this CTA's K/V pages towards L2.  `pos` may still be one step
        // stale here (it only steers a prefetch); the authoritative read happens after pdl_wait().
        const int ctx_h = __ldcg(p.pos + b) + 1;
        int ch = (ctx_h + p.nsplit - 1) / p.nsplit;
        ch = (ch + Q3T_KV_PAGE - 1) / Q3T_KV_PAGE * Q3T_KV_PAGE;
        const int h0 = split * ch, h1 = min(ctx_h, h0 + ch);
        const int npg = h1 > h0 ? (h1 - 1) / Q3T_KV_PAGE - h0 / Q3T_KV_PAGE + 1 : 0;
        for (int i = tid; i < 2 * npg; i += 128) {
            const int pg = h0 / Q3T_KV_PAGE + (i >> 1);
            if (pg < p.max_pages) {
                const __nv_bfloat16* src = p.kv_pool + (size_t)btbl[pg] * page_elems + head_off + ((i & 1) ? v_off : 0);
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(Q3T_KV_PAGE * D * 2) : "memory");
            }
        }
    }
    pdl_wait();
    const int pos = __ldcg(p.pos + b), ctx = pos + 1;
    int chunk = (ctx + p.nsplit - 1) / p.nsplit;
    chunk = (chunk + Q3T_KV_PAGE - 1) / Q3T_KV_PAGE * Q3T_KV_PAGE;
    const int s0 = split * chunk, s1 = min(ctx, s0 + chunk);
    const bool owner = (pos >= s0 && pos < s1);
    const int qkv_dim = (p.H + 2 * p.Hkv) * D;
    const float* row = p.qkv + (size_t)b * qkv_dim;

    // ---- stage 1: q (REP heads), k, v of the new token: RMSNorm + RoPE ------------------------
    if (s0 < s1) {
        for (int vi = warp; vi < REP + 2; vi += 4) {
            const bool is_q = vi < REP, is_k = vi == REP;
            if (!is_q && (!owner || p.mode == 2)) continue;
            if (is_q && p.mode == 1) continue;
            const float* src = is_q ? row + (size_t)(kvh * REP + vi) * D
                                    : (is_k ? row + (size_t)(p.H + kvh) * D : row + (size_t)(p.H + p.Hkv + kvh) * D);
            float x[E];
#pragma unroll
            for (int e = 0; e < E; ++e) x[e] = src[lane * E + e];
            if (is_q || is_k) {
                float ss = 0.f;
#pragma unroll
                for (int e = 0; e < E; ++e) ss += x[e] * x[e];
                ss = warp_sum(ss);
                const float rstd = rsqrtf(ss / (float)D + p.eps);
                const float* nw = is_q ? p.q_norm_w : p.k_norm_w;
#pragma unroll
                for (int e = 0; e < E; ++e) x[e] = nw[lane * E + e] * (x[e] * rstd);
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const float other = __shfl_xor_sync(0xffffffffu, x[e], 16);
                    const float ang = (float)pos * p.inv_freq[(lane & 15) * E + e];
                    float sn, cs;
                    sincosf(ang, &sn, &cs);
                    x[e] = (lane < 16) ? (x[e] * cs - other * sn) : (x[e] * cs + other * sn);
                }
            }
            if (is_q) {
                const float sc = rsqrtf((float)D);
#pragma unroll
                for (int e = 0; e < E; ++e) q_s[vi][lane * E + e] = x[e] * sc;
            } else {
                __nv_bfloat16* dst = p.kv_pool + (size_t)btbl[pos / Q3T_KV_PAGE] * page_elems + head_off +
                                     (is_k ? 0 : v_off) + (size_t)(pos % Q3T_KV_PAGE) * D + lane * E;
#pragma unroll
                for (int e = 0; e < E; ++e) dst[e] = __float2bfloat16_rn(x[e]);
            }
        }
    }
    if (p.mode == 1) return;
    __syncthreads();

    // ---- stage 2: stream the slice --------------------------------------------------------------
    const int hw = lane >> 4, sl = lane & 15, hwid = warp * 2 + hw;
    float m_run[REP], l_run[REP], acc[REP][EPL], qr[REP][EPL];
#pragma unroll
    for (int r = 0; r < REP; ++r) {
        m_run[r] = -INFINITY; l_run[r] = 0.f;
#pragma unroll
        for (int e = 0; e < EPL; ++e) { acc[r][e] = 0.f; qr[r][e] = (s0 < s1) ? q_s[r][sl * EPL + e] : 0.f; }
    }
    // trip count is warp-uniform (both half-warps shuffle together); out-of-range tokens are masked.  One iteration = the 16
    // tokens of ONE page (s0 is page aligned): half-warp h takes rows h and h + 8.  The loop is software pipelined two deep -
    // the four 16-byte loads of the next page are in flight while this page's scores / softmax / PV run, and the block-table
    // entry is fetched one page further ahead - because a CTA's slice is a short dependent chain of L2 / HBM round trips
    // (ctx 300 in 4 slices: 5 pages): un-pipelined, 39 us per layer at batch 64 against 13 us of K/V bytes.
    using Raw = typename KvVec<EPL>::T;
    auto issue = [&](int base, int page, Raw (&raw)[4]) {
        const int tok = base + hwid, tok2 = tok + 8;
        const int r1 = (tok < s1 ? tok : base) % Q3T_KV_PAGE, r2 = (tok2 < s1 ? tok2 : base) % Q3T_KV_PAGE;   // masked rows re-read a valid one
        const __nv_bfloat16* pg = p.kv_pool + (size_t)page * page_elems + head_off + sl * EPL;
        raw[0] = *reinterpret_cast<const Raw*>(pg + (size_t)r1 * D);
        raw[1] = *reinterpret_cast<const Raw*>(pg + (size_t)r1 * D + v_off);
        raw[2] = *reinterpret_cast<const Raw*>(pg + (size_t)r2 * D);
        raw[3] = *reinterpret_cast<const Raw*>(pg + (size_t)r2 * D + v_off);
    };
    auto unpack = [&](const Raw& raw, float (&o)[EPL]) {
        const uint32_t* u = reinterpret_cast<const uint32_t*>(&raw);
#pragma unroll
        for (int i = 0; i < EPL / 2; ++i) { o[2 * i] = bf16lo(u[i]); o[2 * i + 1] = bf16hi(u[i]); }
    };
    Raw cur[4];
    int pg_next = 0;
    if (s0 < s1) {
        issue(s0, btbl[s0 / Q3T_KV_PAGE], cur);
        if (s0 + Q3T_KV_PAGE < s1) pg_next = btbl[s0 / Q3T_KV_PAGE + 1];
    }
    for (int base = s0; base < s1; base += Q3T_KV_PAGE) {
        const int tok = base + hwid, tok2 = tok + 8;
        const bool has1 = tok < s1, has2 = tok2 < s1;
        Raw nxt[4];
        int pg_next2 = 0;
        if (base + Q3T_KV_PAGE < s1) {
            issue(base + Q3T_KV_PAGE, pg_next, nxt);
            if (base + 2 * Q3T_KV_PAGE < s1) pg_next2 = btbl[base / Q3T_KV_PAGE + 2];
        }
        float k0[EPL], v0[EPL], k1[EPL], v1[EPL];
        unpack(cur[0], k0); unpack(cur[1], v0); unpack(cur[2], k1); unpack(cur[3], v1);
#pragma unroll
        for (int r = 0; r < REP; ++r) {
            float sa = 0.f, sb = 0.f;
#pragma unroll
            for (int e = 0; e < EPL; ++e) { sa = fmaf(qr[r][e], k0[e], sa); sb = fmaf(qr[r][e], k1[e], sb); }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                sa += __shfl_xor_sync(0xffffffffu, sa, o);
                sb += __shfl_xor_sync(0xffffffffu, sb, o);
            }
            if (!has1) sa = -INFINITY;
            if (!has2) sb = -INFINITY;
            const float mn = fmaxf(m_run[r], fmaxf(sa, sb));
            if (mn == -INFINITY) continue;             // nothing seen yet by this half-warp
            const float corr = __expf(m_run[r] - mn);   // exp(-inf) = 0 on the first token
            const float pa = __expf(sa - mn), pb = __expf(sb - mn);
            l_run[r] = l_run[r] * corr + pa + pb;
#pragma unroll
            for (int e = 0; e < EPL; ++e) acc[r][e] = fmaf(pb, v1[e], fmaf(pa, v0[e], acc[r][e] * corr));
            m_run[r] = mn;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) cur[i] = nxt[i];
        pg_next = pg_next2;
    }
#pragma unroll
    for (int r = 0; r < REP; ++r) {
#pragma unroll
        for (int e = 0; e < EPL; ++e) part_s[hwid][r][sl * EPL + e] = acc[r][e];
        if (sl == 0) { part_s[hwid][r][D] = m_run[r]; part_s[hwid][r][D + 1] = l_run[r]; }
    }
    __syncthreads();

    // ---- stage 3: merge the 8 half-warps into this slice's record {acc[D], m, l} per query head ---------------------------
    for (int i = tid; i < REP * D; i += 128) {
        const int r = i / D, d = i % D;
        float M = -INFINITY;
#pragma unroll
        for (int h = 0; h < 8; ++h) M = fmaxf(M, part_s[h][r][D]);
        float L = 0.f, A = 0.f;
#pragma unroll
        for (int h = 0; h < 8; ++h) {
            const float mh = part_s[h][r][D];
            const float w = (mh == -INFINITY) ? 0.f : __expf(mh - M);
            L = fmaf(part_s[h][r][D + 1], w, L);
            A = fmaf(part_s[h][r][d], w, A);
        }
        rec_s[r][d] = A;
        if (d == 0) { rec_s[r][D] = M; rec_s[r][D + 1] = L; }
    }
    auto store_out = [&](int r, int d, float v) {
        const size_t oi = (size_t)b * p.H * D + (size_t)(kvh * REP + r) * D + d;
        if (p.out_bf16) {          // feeds the O-projection GEMM without an fp32 round trip: split row [hi(H*D) | lo(H*D)]
            const __nv_bfloat16 hi = __float2bfloat16_rn(v);
            const size_t si = oi + (size_t)b * p.H * D;
            p.out_bf16[si] = hi;
            p.out_bf16[si + (size_t)p.H * D] = __float2bfloat16_rn(v - __bfloat162float(hi));
        } else p.out[oi] = v;
    };
    // ---- stage 4: merge the slices in slice order (deterministic, no float atomics) -----------------------------------------
    // One slice: the record is the result.  Up to 8 slices: the CTAs of one (kv head, sequence) are a thread-block cluster and
    // slice 0 reads the records of its peers through distributed shared memory - no global round trip (write, fence, atomic,
    // re-read: ~3 us of a ~12 us CTA).  More slices: records through the L2 workspace, last CTA to arrive merges.
    if (p.nsplit == 1) {
        __syncthreads();
        for (int i = tid; i < REP * D; i += 128) { const int r = i / D, d = i % D; store_out(r, d, rec_s[r][d] / rec_s[r][D + 1]); }
        return;
    }
    if (p.cluster) {
        asm volatile("barrier.cluster.arrive.release;\nbarrier.cluster.wait.acquire;" ::: "memory");
        if (split == 0) {
            const uint32_t rec_u32 = (uint32_t)__cvta_generic_to_shared(&rec_s[0][0]);
            auto ld_rec = [&](int s, int off) -> float {
                uint32_t ra; float v;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(rec_u32 + (uint32_t)off * 4u), "r"(s));
                asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
                return v;
            };
            for (int i = tid; i < REP * D; i += 128) {
                const int r = i / D, d = i % D;
                float M = -INFINITY;
                for (int s = 0; s < p.nsplit; ++s) M = fmaxf(M, ld_rec(s, r * PSTR + D));
                float L = 0.f, A = 0.f;
                for (int s = 0; s < p.nsplit; ++s) {
                    const float ms = ld_rec(s, r * PSTR + D);
                    const float w = (ms == -INFINITY) ? 0.f : __expf(ms - M);
                    L = fmaf(ld_rec(s, r * PSTR + D + 1), w, L);
                    A = fmaf(ld_rec(s, r * PSTR + d), w, A);
                }
                store_out(r, d, A / L);
            }
        }
        asm volatile("barrier.cluster.arrive.release;\nbarrier.cluster.wait.acquire;" ::: "memory");   // records stay alive until slice 0 has read them
        return;
    }
    __syncthreads();
    float* wbase = p.work + ((size_t)(b * p.Hkv + kvh) * p.nsplit) * REP * PSTR;
    for (int i = tid; i < REP * PSTR; i += 128) wbase[(size_t)split * REP * PSTR + i] = (&rec_s[0][0])[i];
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const int prev = atomicAdd(p.counters + b * p.Hkv + kvh, 1);
        is_last = (prev == p.nsplit - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    for (int i = tid; i < REP * D; i += 128) {
        const int r = i / D, d = i % D;
        float M = -INFINITY;
        for (int s = 0; s < p.nsplit; ++s) M = fmaxf(M, __ldcg(wbase + ((size_t)s * REP + r) * PSTR + D));
        float L = 0.f, A = 0.f;
        for (int s = 0; s < p.nsplit; ++s) {
            const float* rec = wbase + ((size_t)s * REP + r) * PSTR;
            const float ms = __ldcg(rec + D);
            const float w = (ms == -INFINITY) ? 0.f : __expf(ms - M);
            L = fmaf(__ldcg(rec + D + 1), w, L);
            A = fmaf(__ldcg(rec + d), w, A);
        }
        store_out(r, d, A / L);
    }
    if (tid == 0) p.counters[b * p.Hkv + kvh] = 0;   // re-arm for the next launch / graph replay
}

template <int D, int REP>
static int launch_attn_t(const AttnParams& p, cudaStream_t stream) {
    dim3 grid(p.Hkv, p.nsplit, p.B);
    AttnParams q = p;
    q.cluster = (p.nsplit > 1 && p.nsplit <= 8 && p.mode != 1) ? 1 : 0;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = dim3(128); cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (g_use_pdl) { attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[na].val.programmaticStreamSerializationAllowed = 1; ++na; }
    if (q.cluster) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = 1; attr[na].val.clusterDim.y = (unsigned)p.nsplit; attr[na].val.clusterDim.z = 1; ++na;
    }
    cfg.attrs = attr; cfg.numAttrs = na;
    cudaLaunchKernelEx(&cfg, attn_decode_kernel<D, REP>, q);
    Q3T_CHECK_LAUNCH("attn_decode");
    return 0;
}

int launch_attn_decode(const q3t_attn_args* a, cudaStream_t stream) {
    Q3T_REQUIRE(a->H % a->Hkv == 0, "attn_decode: H % Hkv");
    Q3T_REQUIRE(a->nsplit >= 1 && a->nsplit <= 64, "attn_decode: nsplit in [1,64]");
    AttnParams p;
    p.qkv = a->qkv; p.q_norm_w = a->q_norm_w; p.k_norm_w = a->k_norm_w; p.eps = a->eps; p.inv_freq = a->inv_freq;
    p.kv_pool = (__nv_bfloat16*)a->kv_pool; p.block_tbl = a->block_tbl; p.max_pages = a->max_pages; p.pos = a->pos;
    p.out = a->out; p.out_bf16 = (__nv_bfloat16*)a->out_bf16; p.work = a->work; p.counters = a->counters; p.B = a->B; p.H = a->H; p.Hkv = a->Hkv;
    p.nsplit = a->nsplit; p.mode = a->mode; p.seq_of_row = a->seq_of_row;
    const int rep = a->H / a->Hkv;
    if (a->D == 128 && rep == 2) return launch_attn_t<128, 2>(p, stream);
    if (a->D == 128 && rep == 1) return launch_attn_t<128, 1>(p, stream);
    if (a->D == 128 && rep == 4) return launch_attn_t<128, 4>(p, stream);
    if (a->D == 64 && rep == 2) return launch_attn_t<64, 2>(p, stream);
    if (a->D == 64 && rep == 1) return launch_attn_t<64, 1>(p, stream);
    Q3T_REQUIRE(false, "attn_decode: unsupported (head_dim, H/Hkv); built for D in {64,128}, rep in {1,2,4}");
    return 2;
}

}  // namespace q3t

extern "C" int q3t_attn_decode(const q3t_attn_args* a, void* stream) {
    return q3t::launch_attn_decode(a, (cudaStream_t)stream);
}
