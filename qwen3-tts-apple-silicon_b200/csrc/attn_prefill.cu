// Causal GQA attention of the PROMPT rows (prefill pass 2) on the tensor cores: one CTA = up to 32 consecutive rows of
// one sequence x one kv head (2 query heads), K/V tiles of 64 tokens staged once in shared memory and shared by all 64
// (row, head) queries of the CTA.  Replaces mx.fast.scaled_dot_product_attention (+ q RMSNorm + RoPE) on the prompt of the
// reference stack (cousin transformers qwen3/modeling_qwen3.py:294-334).
//
// Why: running the decode kernel once per prompt row re-reads every row's whole K/V prefix - 19 200 rows x 8 kv heads at
// batch 64 x 300 tokens = 11.8 GB of L2 reads per layer and 88 of the 177 ms of the prefill (profiles/README.md).  Here a
// K/V tile is read once per 32 rows: 0.37 GB per layer.
//
//   * q: fp32 row of the fused QKV output -> RMSNorm(q_norm) -> RoPE(pos) -> x 1/sqrt(D) -> bf16 in shared memory
//     (K and V are already in the paged cache as bf16: pass 1 of the prefill wrote them);
//   * S = Q K^T and O += P V with mma.sync.m16n8k16 (bf16 x bf16 -> fp32), operands through ldmatrix (V transposed by
//     ldmatrix.trans), rows padded to 272 bytes so that the eight rows of an ldmatrix hit different banks;
//   * online softmax in registers (FlashAttention-2 layout: one warp = 16 queries, the quad of a row reduces max / sum);
//   * K/V tiles arrive by cp.async (16-byte chunks gathered through the block table), double buffered.
#include "common.cuh"
#include "../../include/q3tts_b200.h"

namespace q3t {

constexpr int PF_QB = 32;                 // query rows per CTA
constexpr int PF_TK = 64;                 // tokens per K/V tile
constexpr int PF_D = 128;
constexpr int PF_LD = PF_D + 8;           // padded row (bf16 elements): 272 bytes
constexpr int PF_THREADS = 128;
constexpr int PF_MAXPAGES = 512;          // pages of one sequence kept in shared memory (8192 tokens)
constexpr int PF_SMEM = (2 * PF_QB + 2 * 2 * PF_TK) * PF_LD * 2 + PF_MAXPAGES * 4;

struct PrefillAttnParams {
    const float* qkv; const float* q_norm_w; float eps; const float* inv_freq;
    const __nv_bfloat16* kv_pool; const int* block_tbl; int max_pages;
    const int* pos; const int* seq_of_row; const int* blocks;      // blocks [n][2] = (first row, row count <= PF_QB)
    float* out; __nv_bfloat16* out_bf16;
    int H, Hkv;
};

__device__ __forceinline__ uint32_t pf_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void pf_ldsm4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void pf_ldsm4t(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void pf_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pf_pack(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
// two adjacent outputs into a split bf16 row [hi(width) | lo(width)] (what the tcgen05 W8 GEMM reads, w8_gemm_tc.cu)
__device__ __forceinline__ void pf_store_split(__nv_bfloat16* dst, size_t width, float a, float b) {
    const uint32_t hi = pf_pack(a, b);
    *reinterpret_cast<uint32_t*>(dst) = hi;
    *reinterpret_cast<uint32_t*>(dst + width) = pf_pack(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xffff0000u));
}

__global__ void __launch_bounds__(PF_THREADS) attn_prefill_kernel(const PrefillAttnParams p) {
    extern __shared__ __align__(16) unsigned char pf_raw[];
    __nv_bfloat16* q_s = reinterpret_cast<__nv_bfloat16*>(pf_raw);                   // [2 heads][PF_QB][PF_LD]
    __nv_bfloat16* kv_s = q_s + 2 * PF_QB * PF_LD;                                   // [2 stages][K|V][PF_TK][PF_LD]
    int* pages = reinterpret_cast<int*>(kv_s + 2 * 2 * PF_TK * PF_LD);               // [PF_MAXPAGES]
    constexpr int D = PF_D, REP = 2;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int kvh = blockIdx.x, blk = blockIdx.y;
    pdl_launch_dependents();
    pdl_wait();
    const int row0 = p.blocks[2 * blk], nrows = p.blocks[2 * blk + 1];
    const int seq = p.seq_of_row[row0], pos0 = p.pos[row0];
    const int maxpos = pos0 + nrows - 1, ntile = maxpos / PF_TK + 1;
    const int* btbl = p.block_tbl + (size_t)seq * p.max_pages;
    const size_t page_elems = (size_t)2 * p.Hkv * Q3T_KV_PAGE * D;
    const size_t head_off = (size_t)kvh * Q3T_KV_PAGE * D, v_off = (size_t)p.Hkv * Q3T_KV_PAGE * D;
    for (int i = tid; i <= maxpos / Q3T_KV_PAGE && i < PF_MAXPAGES; i += PF_THREADS) pages[i] = btbl[i];
    __syncthreads();

    auto load_tile = [&](int t, int stage) {
        // 64 tokens x (K, V) x 16 chunks of 16 bytes; tokens past the block's last position are zero-filled
        __nv_bfloat16* dst0 = kv_s + (size_t)stage * 2 * PF_TK * PF_LD;
        for (int i = tid; i < PF_TK * 2 * 16; i += PF_THREADS) {
            const int tok_l = i >> 5, kv = (i >> 4) & 1, ch = i & 15, tok = t * PF_TK + tok_l;
            const bool ok = tok <= maxpos;
            const __nv_bfloat16* src = p.kv_pool + (size_t)pages[ok ? tok / Q3T_KV_PAGE : 0] * page_elems + head_off +
                                       (kv ? v_off : (size_t)0) + (size_t)(tok % Q3T_KV_PAGE) * D + ch * 8;
            const uint32_t d = pf_smem(dst0 + ((size_t)kv * PF_TK + tok_l) * PF_LD + ch * 8);
            const int sz = ok ? 16 : 0;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    load_tile(0, 0);

    // ---- q: RMSNorm + RoPE + scale -> bf16.  Two threads per (row, head) query; thread `hf` owns dims [32 hf, 32 hf + 32)
    // and their rotation partners [64 + 32 hf, ...), so every rotate_half pair lives in one thread.
    {
        const int qi = tid >> 1, hf = tid & 1;                 // query 0..63 = head * 32 + row
        const int head = qi >> 5, r = qi & 31;
        __nv_bfloat16* dst = q_s + ((size_t)head * PF_QB + r) * PF_LD;
        const int qkv_dim = (p.H + 2 * p.Hkv) * D;
        const float* src = p.qkv + (size_t)(row0 + (r < nrows ? r : 0)) * qkv_dim + (size_t)(kvh * REP + head) * D;
        float ss = 0.f;
        if (r < nrows) {
#pragma unroll 8
            for (int j = 0; j < 32; j += 4) {
                const float4 a = *reinterpret_cast<const float4*>(src + hf * 32 + j), b = *reinterpret_cast<const float4*>(src + 64 + hf * 32 + j);
                ss += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w + b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
            }
        }
        ss += __shfl_xor_sync(0xffffffffu, ss, 1);        // the two threads of a query (whole warp takes part)
        if (r < nrows) {
            const float rstd = rsqrtf(ss / (float)D + p.eps), sc = rsqrtf((float)D);
            const float fpos = (float)(pos0 + r);
            for (int j = 0; j < 32; ++j) {
                const int d = hf * 32 + j;
                const float lo = p.q_norm_w[d] * (src[d] * rstd), hi = p.q_norm_w[d + 64] * (src[d + 64] * rstd);
                float sn, cs;
                sincosf(fpos * p.inv_freq[d], &sn, &cs);
                dst[d] = __float2bfloat16_rn((lo * cs - hi * sn) * sc);
                dst[d + 64] = __float2bfloat16_rn((hi * cs + lo * sn) * sc);
            }
        } else {
            for (int j = 0; j < 32; ++j) { dst[hf * 32 + j] = __float2bfloat16_rn(0.f); dst[64 + hf * 32 + j] = __float2bfloat16_rn(0.f); }
        }
    }
    __syncthreads();

    // ---- this warp: 16 queries = head (warp >> 1), rows 16 (warp & 1) .. +15
    const int head = warp >> 1, rbase = (warp & 1) * 16;
    const int g = lane >> 2, t4 = lane & 3;
    uint32_t qf[8][4];
    {
        const __nv_bfloat16* qb = q_s + ((size_t)head * PF_QB + rbase) * PF_LD;
        const int lr = (lane & 7) + ((lane >> 3) & 1) * 8, lc = (lane >> 4) * 8;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) pf_ldsm4(qf[ks], pf_smem(qb + (size_t)lr * PF_LD + ks * 16 + lc));
    }
    const int pos_a = pos0 + rbase + g, pos_b = pos_a + 8;     // positions of this thread's two query rows
    float o[16][4];
#pragma unroll
    for (int n = 0; n < 16; ++n) { o[n][0] = 0.f; o[n][1] = 0.f; o[n][2] = 0.f; o[n][3] = 0.f; }
    float m_a = -INFINITY, m_b = -INFINITY, l_a = 0.f, l_b = 0.f;

    for (int t = 0; t < ntile; ++t) {
        const int stage = t & 1;
        if (t + 1 < ntile) { load_tile(t + 1, stage ^ 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        const __nv_bfloat16* ks_ = kv_s + (size_t)stage * 2 * PF_TK * PF_LD;
        const __nv_bfloat16* vs_ = ks_ + (size_t)PF_TK * PF_LD;
        // S = Q K^T : 8 n-tiles of 8 tokens
        float s[8][4];
#pragma unroll
        for (int n = 0; n < 8; ++n) { s[n][0] = 0.f; s[n][1] = 0.f; s[n][2] = 0.f; s[n][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
            for (int np = 0; np < 4; ++np) {        // two n-tiles per ldmatrix.x4
                uint32_t b[4];
                const int tok = np * 16 + (lane & 7) + ((lane >> 4) & 1) * 8, col = ks * 16 + ((lane >> 3) & 1) * 8;
                pf_ldsm4(b, pf_smem(ks_ + (size_t)tok * PF_LD + col));
                pf_mma(s[2 * np], qf[ks], b[0], b[1]);
                pf_mma(s[2 * np + 1], qf[ks], b[2], b[3]);
            }
        }
        // causal mask + online softmax (rows g and g + 8 of the warp's 16)
        const int tok0 = t * PF_TK;
        float mx_a = -INFINITY, mx_b = -INFINITY;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const int c0 = tok0 + n * 8 + 2 * t4;
            if (c0 > pos_a) s[n][0] = -INFINITY;
            if (c0 + 1 > pos_a) s[n][1] = -INFINITY;
            if (c0 > pos_b) s[n][2] = -INFINITY;
            if (c0 + 1 > pos_b) s[n][3] = -INFINITY;
            mx_a = fmaxf(mx_a, fmaxf(s[n][0], s[n][1])); mx_b = fmaxf(mx_b, fmaxf(s[n][2], s[n][3]));
        }
        mx_a = fmaxf(mx_a, __shfl_xor_sync(0xffffffffu, mx_a, 1)); mx_a = fmaxf(mx_a, __shfl_xor_sync(0xffffffffu, mx_a, 2));
        mx_b = fmaxf(mx_b, __shfl_xor_sync(0xffffffffu, mx_b, 1)); mx_b = fmaxf(mx_b, __shfl_xor_sync(0xffffffffu, mx_b, 2));
        const float mn_a = fmaxf(m_a, mx_a), mn_b = fmaxf(m_b, mx_b);
        // a row whose every key so far is masked keeps m = -inf: use 0 as the reference so that exp() stays finite
        const float ra = (mn_a == -INFINITY) ? 0.f : mn_a, rb = (mn_b == -INFINITY) ? 0.f : mn_b;
        const float ca = __expf(m_a - ra), cb = __expf(m_b - rb);
        float sum_a = 0.f, sum_b = 0.f;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            s[n][0] = __expf(s[n][0] - ra); s[n][1] = __expf(s[n][1] - ra);
            s[n][2] = __expf(s[n][2] - rb); s[n][3] = __expf(s[n][3] - rb);
            sum_a += s[n][0] + s[n][1]; sum_b += s[n][2] + s[n][3];
        }
        sum_a += __shfl_xor_sync(0xffffffffu, sum_a, 1); sum_a += __shfl_xor_sync(0xffffffffu, sum_a, 2);
        sum_b += __shfl_xor_sync(0xffffffffu, sum_b, 1); sum_b += __shfl_xor_sync(0xffffffffu, sum_b, 2);
        l_a = l_a * ca + sum_a; l_b = l_b * cb + sum_b; m_a = mn_a; m_b = mn_b;
#pragma unroll
        for (int n = 0; n < 16; ++n) { o[n][0] *= ca; o[n][1] *= ca; o[n][2] *= cb; o[n][3] *= cb; }
        // O += P V : k = tokens (4 steps of 16), n = dims (16 tiles of 8)
#pragma unroll
        for (int kt = 0; kt < 4; ++kt) {
            uint32_t a[4];
            a[0] = pf_pack(s[2 * kt][0], s[2 * kt][1]); a[1] = pf_pack(s[2 * kt][2], s[2 * kt][3]);
            a[2] = pf_pack(s[2 * kt + 1][0], s[2 * kt + 1][1]); a[3] = pf_pack(s[2 * kt + 1][2], s[2 * kt + 1][3]);
#pragma unroll
            for (int np = 0; np < 8; ++np) {        // two dim-tiles per ldmatrix.x4.trans
                uint32_t b[4];
                const int tok = kt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, col = np * 16 + ((lane >> 4) & 1) * 8;
                pf_ldsm4t(b, pf_smem(vs_ + (size_t)tok * PF_LD + col));
                pf_mma(o[2 * np], a, b[0], b[1]);
                pf_mma(o[2 * np + 1], a, b[2], b[3]);
            }
        }
        __syncthreads();       // every warp is done with this stage before the next load overwrites it
    }

    // ---- O / l -> out[row, (kvh*REP + head) * D + d]
    const float ia = l_a > 0.f ? 1.f / l_a : 0.f, ib = l_b > 0.f ? 1.f / l_b : 0.f;
    const int ra_ = rbase + g, rb_ = ra_ + 8;
    const size_t ostride = (size_t)p.H * D, hoff = (size_t)(kvh * REP + head) * D;
#pragma unroll
    for (int n = 0; n < 16; ++n) {
        const int d = n * 8 + 2 * t4;
        if (ra_ < nrows) {
            const size_t oi = (size_t)(row0 + ra_) * ostride + hoff + d;
            if (p.out_bf16) pf_store_split(p.out_bf16 + oi + (size_t)(row0 + ra_) * ostride, ostride, o[n][0] * ia, o[n][1] * ia);
            else *reinterpret_cast<float2*>(p.out + oi) = make_float2(o[n][0] * ia, o[n][1] * ia);
        }
        if (rb_ < nrows) {
            const size_t oi = (size_t)(row0 + rb_) * ostride + hoff + d;
            if (p.out_bf16) pf_store_split(p.out_bf16 + oi + (size_t)(row0 + rb_) * ostride, ostride, o[n][2] * ib, o[n][3] * ib);
            else *reinterpret_cast<float2*>(p.out + oi) = make_float2(o[n][2] * ib, o[n][3] * ib);
        }
    }
}

// ---- pass 1: K (RMSNorm + RoPE) and V of every prompt row -> bf16 rows of the paged cache.  One warp per (row, kv head),
// four dims per lane; the arithmetic (and therefore every stored bit) is the decode kernel's (attn_decode.cu, stage 1).
__global__ void __launch_bounds__(128) kv_write_kernel(const float* __restrict__ qkv, const float* __restrict__ k_norm_w, float eps,
                                                       const float* __restrict__ inv_freq, __nv_bfloat16* kv_pool,
                                                       const int* __restrict__ block_tbl, int max_pages, const int* __restrict__ pos,
                                                       const int* __restrict__ seq_of_row, int M, int H, int Hkv) {
    constexpr int D = PF_D, E = 4;
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31, m = blockIdx.x * 4 + (threadIdx.x >> 5), kvh = blockIdx.y;
    if (m >= M) return;
    const int p = pos[m];
    const float* row = qkv + (size_t)m * (H + 2 * Hkv) * D;
    const size_t page_elems = (size_t)2 * Hkv * Q3T_KV_PAGE * D;
    __nv_bfloat16* dst = kv_pool + (size_t)block_tbl[(size_t)seq_of_row[m] * max_pages + p / Q3T_KV_PAGE] * page_elems +
                         (size_t)kvh * Q3T_KV_PAGE * D + (size_t)(p % Q3T_KV_PAGE) * D + lane * E;
    {   // k
        const float* src = row + (size_t)(H + kvh) * D;
        float x[E];
#pragma unroll
        for (int e = 0; e < E; ++e) x[e] = src[lane * E + e];
        float ss = 0.f;
#pragma unroll
        for (int e = 0; e < E; ++e) ss += x[e] * x[e];
        ss = warp_sum(ss);
        const float rstd = rsqrtf(ss / (float)D + eps);
#pragma unroll
        for (int e = 0; e < E; ++e) x[e] = k_norm_w[lane * E + e] * (x[e] * rstd);
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const float other = __shfl_xor_sync(0xffffffffu, x[e], 16);
            float sn, cs;
            sincosf((float)p * inv_freq[(lane & 15) * E + e], &sn, &cs);
            x[e] = (lane < 16) ? (x[e] * cs - other * sn) : (x[e] * cs + other * sn);
        }
#pragma unroll
        for (int e = 0; e < E; ++e) dst[e] = __float2bfloat16_rn(x[e]);
    }
    {   // v
        const float* src = row + (size_t)(H + Hkv + kvh) * D;
        __nv_bfloat16* dv = dst + (size_t)Hkv * Q3T_KV_PAGE * D;
#pragma unroll
        for (int e = 0; e < E; ++e) dv[e] = __float2bfloat16_rn(src[lane * E + e]);
    }
}

int launch_attn_prefill(const q3t_attn_prefill_args* a, cudaStream_t stream) {
    Q3T_REQUIRE(a->D == PF_D && a->H == 2 * a->Hkv, "attn_prefill: built for head_dim 128 and two query heads per kv head");
    Q3T_REQUIRE(a->n_blocks >= 1 && a->blocks && a->pos && a->seq_of_row && a->qkv && (a->out || a->out_bf16), "attn_prefill: arguments");
    Q3T_REQUIRE(a->max_pages <= PF_MAXPAGES, "attn_prefill: context too long for the per-CTA page table");
    PrefillAttnParams p;
    p.qkv = a->qkv; p.q_norm_w = a->q_norm_w; p.eps = a->eps; p.inv_freq = a->inv_freq;
    p.kv_pool = (const __nv_bfloat16*)a->kv_pool; p.block_tbl = a->block_tbl; p.max_pages = a->max_pages;
    p.pos = a->pos; p.seq_of_row = a->seq_of_row; p.blocks = a->blocks; p.out = a->out; p.out_bf16 = (__nv_bfloat16*)a->out_bf16;
    p.H = a->H; p.Hkv = a->Hkv;
    if (a->k_norm_w) {     // pass 1 (optional): every row's K/V into the cache before any row attends
        Q3T_REQUIRE(a->M >= 1, "attn_prefill: M rows for the K/V write");
        launch_pdl(kv_write_kernel, dim3((a->M + 3) / 4, a->Hkv), dim3(128), 0, stream, a->qkv, a->k_norm_w, a->eps, a->inv_freq,
                   (__nv_bfloat16*)a->kv_pool, a->block_tbl, a->max_pages, a->pos, a->seq_of_row, a->M, a->H, a->Hkv);
        Q3T_CHECK_LAUNCH("kv_write");
    }
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(attn_prefill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PF_SMEM);
        attr_set = true;
    }
    launch_pdl(attn_prefill_kernel, dim3(a->Hkv, a->n_blocks), dim3(PF_THREADS), (size_t)PF_SMEM, stream, p);
    Q3T_CHECK_LAUNCH("attn_prefill");
    return 0;
}

}  // namespace q3t

extern "C" int q3t_attn_prefill(const q3t_attn_prefill_args* a, void* stream) {
    return q3t::launch_attn_prefill(a, (cudaStream_t)stream);
}
