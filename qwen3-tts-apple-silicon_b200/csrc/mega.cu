// Persistent "stack pass" kernel: ONE cooperative launch runs a token through every layer of a dense Qwen3 stack
// (talker: 28 layers, code predictor: 5 layers), the final norm and the head GEMV, for batch 1.
//
// Why: at batch 1 a layer streams only 53 MB (8 us of HBM time on a B200) split over five dependent contractions; with
// one kernel per contraction the ~4 us launch + prologue + tail floor of each kernel, not HBM, sets the step time
// (profiles/r01_launches_multikernel_pdl.csv).  Here the whole step is one grid of #SM CTAs:
//   * a dedicated producer warp per CTA streams this CTA's share of EVERY weight matrix of the step, in order, through
//     a shared-memory ring with cp.async.bulk (TMA, `UBLKCP`) + mbarrier full/empty pairs.  It never takes part in the
//     phase barriers, so HBM keeps streaming while the consumers exchange activations between phases;
//   * 16 consumer warps run the IMMA dequant-dot of w8_gemv.cu out of the ring (a 4352-byte tile per warp) against the
//     digit planes of the phase input held in shared memory;
//   * tiles of a matrix are dealt out contiguously in [row_tile][k_chunk] order, balanced to +-1 tile per CTA; a row
//     tile that straddles CTAs is completed by summing per-CTA partial rows ("slots") in fixed order in the consumer
//     phase's prologue - no atomics, bit-reproducible;
//   * every CTA keeps its own fp32 copy of the residual stream in shared memory (identical arithmetic everywhere), so the
//     residual adds never touch global memory;
//   * phases are separated by a grid-wide barrier (one atomic + one acquire poll per CTA).
// Per layer: QKV | attention (split-KV, last-arriver merge) | O | gate-up | down  = 5 barriers.
#include <cooperative_groups.h>
#include "common.cuh"
#include "../../include/q3tts_b200.h"

namespace q3t {

constexpr int MG_CWARPS = 16;                    // consumer warps
constexpr int MG_CTHREADS = MG_CWARPS * 32;      // 512
constexpr int MG_THREADS = MG_CTHREADS + 32;     // + producer warp
constexpr int MG_NSLOT = 32;                     // ring slots of one tile each (32 * 4352 = 139 264 B)
constexpr int MG_MAXSLOT = 4;                    // partial-row slots per row tile
constexpr int MG_MAXT = 64;                      // max tiles of one matrix per CTA
constexpr int MG_MAXK = 8192;

struct MegaParams {
    const q3t_layer* layers;                     // DEVICE array [n_layers]
    int n_layers, hidden, n_heads, n_kv, head_dim, inter;
    float eps;
    const float* final_norm; const float* inv_freq;
    __nv_bfloat16* kv_pool; long long kv_layer_stride;   // elements
    const int* block_tbl; int max_pages; const int* pos;
    q3t_w8 head; int has_head;
    const float* x_in; float* hidden_out; float* logits_out;
    float* part_qkv; float* part_o; float* part_gu; float* part_down; float* part_head;   // [MG_MAXSLOT][N]
    float* attn_out; float* attn_work; int* attn_counters; int attn_nsplit;
    unsigned int* bar;
    unsigned long long* timing;                  // optional: [gridDim][MG_NSTAMP] globaltimer stamps of thread 0
};
constexpr int MG_NSTAMP = 1024;

__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define MG_STAMP() do { if (p.timing && threadIdx.x == 0 && st.nstamp < MG_NSTAMP) p.timing[(size_t)blockIdx.x * MG_NSTAMP + st.nstamp++] = gtimer(); } while (0)

// ---- tile geometry ------------------------------------------------------------------------------------------------
struct Geo {
    int N, K, nrt, nkc, T, G;   // G = active CTAs for this matrix
    __device__ __forceinline__ void init(int n, int k, int grid) {
        N = n; K = k; nrt = n >> 4; nkc = k >> 8; T = nrt * nkc; G = T < grid ? T : grid;
    }
    __device__ __forceinline__ int t0(int c) const { return c >= G ? T : (int)(((long long)T * c) / G); }
    __device__ __forceinline__ int owner(int tile) const { return (int)((((long long)tile + 1) * G - 1) / T); }
    __device__ __forceinline__ int nslots(int rt) const { return owner(rt * nkc + nkc - 1) - owner(rt * nkc) + 1; }
};

// ---- mbarrier / TMA bulk PTX ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MBAR_DONE;\n"
        "bra MBAR_WAIT;\n"
        "MBAR_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void cbar() { asm volatile("bar.sync 1, %0;" ::"n"(MG_CTHREADS) : "memory"); }

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void imma_16832_m(int (&c)[4], const uint4 a, const uint32_t b0, const uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
        : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}

// ---- shared memory ----------------------------------------------------------------------------------------------------
struct MegaSmem {
    uint8_t* ring;          // [MG_NSLOT][4352]
    uint4* xfrag;           // [MG_MAXK/64][32]   (also reused as attention scratch)
    float* xsum;            // [MG_MAXK/64]
    float* xscl;            // [MG_MAXK/64]
    float* resid;           // [hidden]
    float* tile_out;        // [MG_MAXT][16]
    float* red;             // [64]
    uint64_t* full;         // [MG_NSLOT]
    uint64_t* empty;        // [MG_NSLOT]
};

__device__ __forceinline__ float cblock_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    cbar();
    if (lane == 0) red[wid] = v;
    cbar();
    float r = (lane < MG_CWARPS) ? red[lane] : 0.f;
    return warp_sum(r);
}

// Sum of the per-CTA partial rows of a row tile, fixed slot order.
__device__ __forceinline__ float4 sum_slots4(const float* part, int N, int n, int ns) {
    float4 a = __ldcg(reinterpret_cast<const float4*>(part + n));
    for (int s = 1; s < ns; ++s) {
        const float4 b = __ldcg(reinterpret_cast<const float4*>(part + (size_t)s * N + n));
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    return a;
}

// v (4 consecutive inputs starting at k = 4*k4) -> digit planes + per-group sum/scale.  Whole warps call this together.
__device__ __forceinline__ void emit_digits(const MegaSmem& s, float4 v, int k4, int lane) {
    float amax = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    const float inv = amax > 0.f ? 1073741824.f / amax : 0.f;
    const float xscale = amax * (1.f / 1073741824.f);
    int e[4] = {__float2int_rn(v.x * inv), __float2int_rn(v.y * inv), __float2int_rn(v.z * inv), __float2int_rn(v.w * inv)};
    long long gs = (long long)e[0] + e[1] + e[2] + e[3];
    uint32_t wd[4] = {0, 0, 0, 0};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        int v0 = e[q];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const int dg = (int)(signed char)(v0 & 0xff);
            wd[d] |= (uint32_t)(dg & 0xff) << (8 * q);
            v0 = (v0 - dg) >> 8;
        }
        wd[3] |= (uint32_t)(v0 & 0xff) << (8 * q);
    }
    const int k = k4 << 2, G = k >> 6, kk = k & 63;
    const int r = ((kk >> 5) << 1) | ((kk >> 4) & 1), t = (kk >> 2) & 3;
    uint32_t* base = reinterpret_cast<uint32_t*>(s.xfrag + G * 32);
#pragma unroll
    for (int d = 0; d < 4; ++d) base[(d * 4 + t) * 4 + r] = wd[d];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) gs += __shfl_xor_sync(0xffffffffu, gs, o);
    if ((lane & 15) == 0) { s.xsum[G] = (float)gs * xscale; s.xscl[G] = xscale; }
}

// One 4352-byte tile out of the ring against the digit planes: 16 output rows (partial over this tile's 256 inputs).
__device__ __forceinline__ void tile_dot(const MegaSmem& s, const uint8_t* tile, int kc, int lane, float& out_lo, float& out_hi) {
    const int g = lane >> 2, t = lane & 3;
    const uint4 mlo = *reinterpret_cast<const uint4*>(tile + 4096 + g * 16);
    const uint4 mhi = *reinterpret_cast<const uint4*>(tile + 4096 + (g + 8) * 16);
    const uint32_t slo_w[2] = {mlo.x, mlo.y}, shi_w[2] = {mhi.x, mhi.y}, blo_w[2] = {mlo.z, mlo.w}, bhi_w[2] = {mhi.z, mhi.w};
    float f[4] = {0.f, 0.f, 0.f, 0.f}, bacc_lo = 0.f, bacc_hi = 0.f;
#pragma unroll
    for (int j4 = 0; j4 < 4; ++j4) {
        const int G = kc * 4 + j4;
        const uint4 a0 = *reinterpret_cast<const uint4*>(tile + (j4 * 2 + 0) * 512 + lane * 16);
        const uint4 a1 = *reinterpret_cast<const uint4*>(tile + (j4 * 2 + 1) * 512 + lane * 16);
        const uint4 b = s.xfrag[G * 32 + lane];
        int acc[4] = {0, 0, 0, 0};
        imma_16832_m(acc, a0, b.x, b.y);
        imma_16832_m(acc, a1, b.z, b.w);
        const float xg = s.xscl[G], xs = s.xsum[G];
        const uint32_t sw_lo = slo_w[j4 >> 1], sw_hi = shi_w[j4 >> 1], bw_lo = blo_w[j4 >> 1], bw_hi = bhi_w[j4 >> 1];
        const float slo = ((j4 & 1) ? bf16hi(sw_lo) : bf16lo(sw_lo)) * xg;
        const float shi = ((j4 & 1) ? bf16hi(sw_hi) : bf16lo(sw_hi)) * xg;
        f[0] = fmaf(slo, (float)acc[0], f[0]);
        f[1] = fmaf(slo, (float)acc[1], f[1]);
        f[2] = fmaf(shi, (float)acc[2], f[2]);
        f[3] = fmaf(shi, (float)acc[3], f[3]);
        bacc_lo = fmaf((j4 & 1) ? bf16hi(bw_lo) : bf16lo(bw_lo), xs, bacc_lo);
        bacc_hi = fmaf((j4 & 1) ? bf16hi(bw_hi) : bf16lo(bw_hi), xs, bacc_hi);
    }
    const float pw_lo = (t & 1) ? 65536.f : 1.f, pw_hi = pw_lo * 256.f;
    float v_lo = f[0] * pw_lo + f[1] * pw_hi, v_hi = f[2] * pw_lo + f[3] * pw_hi;
    v_lo += __shfl_xor_sync(0xffffffffu, v_lo, 1);
    v_hi += __shfl_xor_sync(0xffffffffu, v_hi, 1);
    out_lo = v_lo + bacc_lo;   // valid in lanes with t == 0 (columns 0..3 = the four digits of batch row 0)
    out_hi = v_hi + bacc_hi;
}

struct ConsumerState { uint32_t tile_i; unsigned int bar_target; int nstamp; };

// GEMV phase (consumers): tiles of this CTA out of the ring -> per-CTA partial rows in global `part`.
__device__ __forceinline__ void gemv_phase(const MegaSmem& s, const Geo& geo, float* part, ConsumerState& st, int cta) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tb = geo.t0(cta), te = geo.t0(cta + 1), nt = te - tb;
    for (int j = warp; j < nt; j += MG_CWARPS) {
        const uint32_t i = st.tile_i + j, slot = i % MG_NSLOT, par = (i / MG_NSLOT) & 1;
        mbar_wait(smem_u32(&s.full[slot]), par);
        float lo, hi;
        tile_dot(s, s.ring + (size_t)slot * Q3T_TILE_BYTES, (tb + j) % geo.nkc, lane, lo, hi);
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&s.empty[slot]));
        if ((lane & 3) == 0) { s.tile_out[j * 16 + (lane >> 2)] = lo; s.tile_out[j * 16 + (lane >> 2) + 8] = hi; }
    }
    st.tile_i += nt;
    cbar();
    // per local row tile: sum its tiles in order, publish in this CTA's slot
    if (nt > 0) {
        const int rt_first = tb / geo.nkc, rt_last = (te - 1) / geo.nkc;
        const int nloc = (rt_last - rt_first + 1) * 16;
        for (int i = tid; i < nloc; i += MG_CTHREADS) {
            const int rt = rt_first + (i >> 4), row = i & 15;
            const int j0 = max(rt * geo.nkc, tb) - tb, j1 = min((rt + 1) * geo.nkc, te) - tb;
            float v = 0.f;
            for (int j = j0; j < j1; ++j) v += s.tile_out[j * 16 + row];
            const int slot = cta - geo.owner(rt * geo.nkc);
            part[(size_t)slot * geo.N + rt * 16 + row] = v;
        }
    }
}

__device__ __forceinline__ void grid_sync(const MegaParams& p, ConsumerState& st) {
    cbar();
    if (threadIdx.x == 0) {
        st.bar_target += gridDim.x;
        __threadfence();
        atomicAdd(p.bar, 1u);
        while (ld_acquire_u32(p.bar) < st.bar_target) {}
        __threadfence();
    }
    cbar();
}

// ---- attention phase (consumers of CTA c < Hkv*nsplit) -------------------------------------------------------------------
template <int D, int REP>
__device__ __forceinline__ void attn_phase(const MegaParams& p, const MegaSmem& s, const Geo& gq, int layer, int cta) {
    constexpr int E = D / 32, EPL = D / 16, PSTR = D + 2, NHW = MG_CWARPS * 2;
    const int units = p.n_kv * p.attn_nsplit;
    if (cta >= units) return;
    const int kvh = cta / p.attn_nsplit, split = cta % p.attn_nsplit;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* q_s = reinterpret_cast<float*>(s.xfrag);              // [REP][D]
    float* part_s = q_s + REP * D;                               // [NHW][REP][PSTR]
    int* flag_s = reinterpret_cast<int*>(part_s + NHW * REP * PSTR);
    const int pos = __ldcg(p.pos), ctx = pos + 1;
    int chunk = (ctx + p.attn_nsplit - 1) / p.attn_nsplit;
    chunk = (chunk + Q3T_KV_PAGE - 1) / Q3T_KV_PAGE * Q3T_KV_PAGE;
    const int s0 = split * chunk, s1 = min(ctx, s0 + chunk);
    const bool owner = (pos >= s0 && pos < s1);
    __nv_bfloat16* pool = p.kv_pool + (size_t)layer * p.kv_layer_stride;
    const size_t page_elems = (size_t)2 * p.n_kv * Q3T_KV_PAGE * D;
    const size_t head_off = (size_t)kvh * Q3T_KV_PAGE * D;
    const size_t v_off = (size_t)p.n_kv * Q3T_KV_PAGE * D;
    const int H = p.n_heads;

    if (s0 < s1) {
        for (int vi = warp; vi < REP + 2; vi += MG_CWARPS) {
            const bool is_q = vi < REP, is_k = vi == REP;
            if (!is_q && !owner) continue;
            const int n0 = (is_q ? (kvh * REP + vi) : (is_k ? (H + kvh) : (H + p.n_kv + kvh))) * D + lane * E;
            const int ns = gq.nslots(n0 >> 4);
            float x[E];
#pragma unroll
            for (int e = 0; e < E; ++e) {
                float a = __ldcg(p.part_qkv + n0 + e);
                for (int sl = 1; sl < ns; ++sl) a += __ldcg(p.part_qkv + (size_t)sl * gq.N + n0 + e);
                x[e] = a;
            }
            if (is_q || is_k) {
                float ss = 0.f;
#pragma unroll
                for (int e = 0; e < E; ++e) ss += x[e] * x[e];
                ss = warp_sum(ss);
                const float rstd = rsqrtf(ss / (float)D + p.eps);
                const float* nw = is_q ? p.layers[layer].q_norm : p.layers[layer].k_norm;
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
                for (int e = 0; e < E; ++e) q_s[vi * D + lane * E + e] = x[e] * sc;
            } else {
                __nv_bfloat16* dst = pool + (size_t)p.block_tbl[pos / Q3T_KV_PAGE] * page_elems + head_off +
                                     (is_k ? 0 : v_off) + (size_t)(pos % Q3T_KV_PAGE) * D + lane * E;
#pragma unroll
                for (int e = 0; e < E; ++e) dst[e] = __float2bfloat16_rn(x[e]);
            }
        }
    }
    cbar();
    const int hw = lane >> 4, sl = lane & 15, hwid = warp * 2 + hw;
    float m_run[REP], l_run[REP], acc[REP][EPL], qr[REP][EPL];
#pragma unroll
    for (int r = 0; r < REP; ++r) {
        m_run[r] = -INFINITY; l_run[r] = 0.f;
#pragma unroll
        for (int e = 0; e < EPL; ++e) { acc[r][e] = 0.f; qr[r][e] = (s0 < s1) ? q_s[r * D + sl * EPL + e] : 0.f; }
    }
    for (int base = s0; base < s1; base += NHW) {
        const int tok = base + hwid;
        const bool has = tok < s1;
        const int t1 = has ? tok : s0;
        const __nv_bfloat16* kp = pool + (size_t)p.block_tbl[t1 / Q3T_KV_PAGE] * page_elems + head_off +
                                  (size_t)(t1 % Q3T_KV_PAGE) * D + sl * EPL;
        float k0[EPL], v0[EPL];
        {
            const uint4 kr = *reinterpret_cast<const uint4*>(kp);
            const uint4 vr = *reinterpret_cast<const uint4*>(kp + v_off);
            const uint32_t* ku = reinterpret_cast<const uint32_t*>(&kr);
            const uint32_t* vu = reinterpret_cast<const uint32_t*>(&vr);
#pragma unroll
            for (int i = 0; i < EPL / 2; ++i) {
                k0[2 * i] = bf16lo(ku[i]); k0[2 * i + 1] = bf16hi(ku[i]);
                v0[2 * i] = bf16lo(vu[i]); v0[2 * i + 1] = bf16hi(vu[i]);
            }
        }
#pragma unroll
        for (int r = 0; r < REP; ++r) {
            float sa = 0.f;
#pragma unroll
            for (int e = 0; e < EPL; ++e) sa = fmaf(qr[r][e], k0[e], sa);
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) sa += __shfl_xor_sync(0xffffffffu, sa, o);
            if (has) {
                const float mn = fmaxf(m_run[r], sa);
                const float corr = __expf(m_run[r] - mn), pa = __expf(sa - mn);
                l_run[r] = l_run[r] * corr + pa;
#pragma unroll
                for (int e = 0; e < EPL; ++e) acc[r][e] = fmaf(pa, v0[e], acc[r][e] * corr);
                m_run[r] = mn;
            }
        }
    }
#pragma unroll
    for (int r = 0; r < REP; ++r) {
#pragma unroll
        for (int e = 0; e < EPL; ++e) part_s[(hwid * REP + r) * PSTR + sl * EPL + e] = acc[r][e];
        if (sl == 0) { part_s[(hwid * REP + r) * PSTR + D] = m_run[r]; part_s[(hwid * REP + r) * PSTR + D + 1] = l_run[r]; }
    }
    cbar();
    float* wbase = p.attn_work + ((size_t)kvh * p.attn_nsplit) * REP * PSTR;
    for (int i = tid; i < REP * D; i += MG_CTHREADS) {
        const int r = i / D, d = i % D;
        float M = -INFINITY;
        for (int h = 0; h < NHW; ++h) M = fmaxf(M, part_s[(h * REP + r) * PSTR + D]);
        float L = 0.f, A = 0.f;
        for (int h = 0; h < NHW; ++h) {
            const float mh = part_s[(h * REP + r) * PSTR + D];
            const float w = (mh == -INFINITY) ? 0.f : __expf(mh - M);
            L = fmaf(part_s[(h * REP + r) * PSTR + D + 1], w, L);
            A = fmaf(part_s[(h * REP + r) * PSTR + d], w, A);
        }
        float* rec = wbase + ((size_t)split * REP + r) * PSTR;
        rec[d] = A;
        if (d == 0) { rec[D] = M; rec[D + 1] = L; }
    }
    __threadfence();
    cbar();
    if (tid == 0) {
        const int prev = atomicAdd(p.attn_counters + kvh, 1);
        *flag_s = (prev == p.attn_nsplit - 1);
    }
    cbar();
    if (*flag_s) {
        __threadfence();
        for (int i = tid; i < REP * D; i += MG_CTHREADS) {
            const int r = i / D, d = i % D;
            float M = -INFINITY;
            for (int sp = 0; sp < p.attn_nsplit; ++sp) M = fmaxf(M, __ldcg(wbase + ((size_t)sp * REP + r) * PSTR + D));
            float L = 0.f, A = 0.f;
            for (int sp = 0; sp < p.attn_nsplit; ++sp) {
                const float* rec = wbase + ((size_t)sp * REP + r) * PSTR;
                const float ms = __ldcg(rec + D);
                const float w = (ms == -INFINITY) ? 0.f : __expf(ms - M);
                L = fmaf(__ldcg(rec + D + 1), w, L);
                A = fmaf(__ldcg(rec + d), w, A);
            }
            p.attn_out[(size_t)(kvh * REP + r) * D + d] = A / L;
        }
        if (tid == 0) p.attn_counters[kvh] = 0;
    }
}

// ---- the kernel -------------------------------------------------------------------------------------------------------------
template <int D, int REP>
__global__ void __launch_bounds__(MG_THREADS, 1) stack_pass_kernel(const MegaParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    MegaSmem s;
    {
        unsigned char* q = smem_raw;
        s.ring = q; q += (size_t)MG_NSLOT * Q3T_TILE_BYTES;
        s.xfrag = reinterpret_cast<uint4*>(q); q += (size_t)(MG_MAXK / 64) * 512;
        s.xsum = reinterpret_cast<float*>(q); q += (MG_MAXK / 64) * 4;
        s.xscl = reinterpret_cast<float*>(q); q += (MG_MAXK / 64) * 4;
        s.resid = reinterpret_cast<float*>(q); q += (size_t)p.hidden * 4;
        s.tile_out = reinterpret_cast<float*>(q); q += MG_MAXT * 16 * 4;
        s.red = reinterpret_cast<float*>(q); q += 64 * 4;
        s.full = reinterpret_cast<uint64_t*>(q); q += MG_NSLOT * 8;
        s.empty = reinterpret_cast<uint64_t*>(q);
    }
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, cta = blockIdx.x, grid = gridDim.x;
    const int qkv_n = (p.n_heads + 2 * p.n_kv) * p.head_dim, q_dim = p.n_heads * p.head_dim;
    Geo g_qkv, g_o, g_gu, g_down, g_head;
    g_qkv.init(qkv_n, p.hidden, grid); g_o.init(p.hidden, q_dim, grid); g_gu.init(2 * p.inter, p.hidden, grid);
    g_down.init(p.hidden, p.inter, grid); g_head.init(p.has_head ? p.head.N : 16, p.hidden, grid);

    if (tid == 0) {
        for (int i = 0; i < MG_NSLOT; ++i) { mbar_init(smem_u32(&s.full[i]), 1); mbar_init(smem_u32(&s.empty[i]), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == MG_CWARPS) {
        // =========================== producer: stream this CTA's tiles of every matrix, in phase order ===============
        if (lane == 0) {
            uint32_t i = 0;
            auto stream = [&](const void* w, const Geo& geo) {
                const int tb = geo.t0(cta), te = geo.t0(cta + 1);
                const uint8_t* src = reinterpret_cast<const uint8_t*>(w) + (size_t)tb * Q3T_TILE_BYTES;
                for (int t = tb; t < te; ++t, ++i, src += Q3T_TILE_BYTES) {
                    const uint32_t slot = i % MG_NSLOT, par = (i / MG_NSLOT) & 1;
                    mbar_wait(smem_u32(&s.empty[slot]), par ^ 1);
                    const uint32_t fb = smem_u32(&s.full[slot]);
                    mbar_expect_tx(fb, Q3T_TILE_BYTES);
                    tma_load_1d(smem_u32(s.ring + (size_t)slot * Q3T_TILE_BYTES), src, Q3T_TILE_BYTES, fb);
                }
            };
            for (int l = 0; l < p.n_layers; ++l) {
                const q3t_layer& L = p.layers[l];
                stream(L.qkv.w, g_qkv); stream(L.o.w, g_o); stream(L.gate_up.w, g_gu); stream(L.down.w, g_down);
            }
            if (p.has_head) stream(p.head.w, g_head);
        }
        return;
    }

    // =============================== consumers ================================================================
    ConsumerState st;
    st.tile_i = 0; st.bar_target = 0; st.nstamp = 0;
    MG_STAMP();
    const int H = p.hidden, H4 = H >> 2;
    // residual stream: private copy per CTA
    for (int k4 = tid; k4 < H4; k4 += MG_CTHREADS)
        reinterpret_cast<float4*>(s.resid)[k4] = __ldcg(reinterpret_cast<const float4*>(p.x_in) + k4);
    cbar();

    // x = rmsnorm(resid [+ sum of partial rows of the previous down projection]) -> digit planes
    auto norm_prologue = [&](const float* add_part, const Geo* add_geo, const float* norm_w, float* hidden_out) {
        float4 xv[MG_MAXK / 4 / MG_CTHREADS];
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < MG_MAXK / 4 / MG_CTHREADS; ++i) {
            const int k4 = tid + i * MG_CTHREADS;
            if (k4 < H4) {
                float4 v = reinterpret_cast<float4*>(s.resid)[k4];
                if (add_part) {
                    const float4 a = sum_slots4(add_part, add_geo->N, k4 << 2, add_geo->nslots(k4 >> 2));
                    v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
                    reinterpret_cast<float4*>(s.resid)[k4] = v;
                }
                xv[i] = v;
                ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
            }
        }
        const float rstd = rsqrtf(cblock_sum(ss, s.red) / (float)H + p.eps);
#pragma unroll
        for (int i = 0; i < MG_MAXK / 4 / MG_CTHREADS; ++i) {
            const int k4 = tid + i * MG_CTHREADS;
            if (k4 < H4) {
                const float4 nw = reinterpret_cast<const float4*>(norm_w)[k4];
                float4 v = xv[i];
                v.x = nw.x * (v.x * rstd); v.y = nw.y * (v.y * rstd); v.z = nw.z * (v.z * rstd); v.w = nw.w * (v.w * rstd);
                if (hidden_out && cta == 0) reinterpret_cast<float4*>(hidden_out)[k4] = v;
                emit_digits(s, v, k4, lane);
            }
        }
        cbar();
    };

    for (int l = 0; l < p.n_layers; ++l) {
        const q3t_layer& L = p.layers[l];
        // ---- QKV
        norm_prologue(l > 0 ? p.part_down : nullptr, &g_down, L.input_norm, nullptr);
        MG_STAMP();
        gemv_phase(s, g_qkv, p.part_qkv, st, cta);
        MG_STAMP();
        grid_sync(p, st);
        MG_STAMP();
        // ---- attention
        attn_phase<D, REP>(p, s, g_qkv, l, cta);
        MG_STAMP();
        grid_sync(p, st);
        MG_STAMP();
        // ---- O projection: x = attention output
        {
            const int K4 = q_dim >> 2;
            for (int k4 = tid; k4 < ((K4 + 31) & ~31); k4 += MG_CTHREADS)
                if (k4 < K4) emit_digits(s, __ldcg(reinterpret_cast<const float4*>(p.attn_out) + k4), k4, lane);
            cbar();
        }
        MG_STAMP();
        gemv_phase(s, g_o, p.part_o, st, cta);
        MG_STAMP();
        grid_sync(p, st);
        MG_STAMP();
        // ---- gate/up: resid += O ; x = rmsnorm(resid)
        norm_prologue(p.part_o, &g_o, L.post_norm, nullptr);
        MG_STAMP();
        gemv_phase(s, g_gu, p.part_gu, st, cta);
        MG_STAMP();
        grid_sync(p, st);
        MG_STAMP();
        // ---- down: x = silu(gate) * up
        {
            const int I = p.inter, K4 = I >> 2;
            for (int k4 = tid; k4 < ((K4 + 31) & ~31); k4 += MG_CTHREADS) {
                if (k4 < K4) {
                    const int n = k4 << 2;
                    const float4 gt = sum_slots4(p.part_gu, g_gu.N, n, g_gu.nslots(n >> 4));
                    const float4 up = sum_slots4(p.part_gu, g_gu.N, I + n, g_gu.nslots((I + n) >> 4));
                    float4 v;
                    v.x = silu_f(gt.x) * up.x; v.y = silu_f(gt.y) * up.y; v.z = silu_f(gt.z) * up.z; v.w = silu_f(gt.w) * up.w;
                    emit_digits(s, v, k4, lane);
                }
            }
            cbar();
        }
        MG_STAMP();
        gemv_phase(s, g_down, p.part_down, st, cta);
        MG_STAMP();
        grid_sync(p, st);
        MG_STAMP();
    }
    // ---- final norm (+ head)
    norm_prologue(p.part_down, &g_down, p.final_norm, p.hidden_out);
    if (p.has_head) {
        gemv_phase(s, g_head, p.part_head, st, cta);
        grid_sync(p, st);
        const int V4 = g_head.N >> 2;
        for (int k4 = cta * MG_CTHREADS + tid; k4 < V4; k4 += grid * MG_CTHREADS)
            reinterpret_cast<float4*>(p.logits_out)[k4] = sum_slots4(p.part_head, g_head.N, k4 << 2, g_head.nslots(k4 >> 2));
    }
}

static size_t mega_smem_bytes(int hidden) {
    return (size_t)MG_NSLOT * Q3T_TILE_BYTES + (size_t)(MG_MAXK / 64) * 512 + 2 * (MG_MAXK / 64) * 4 + (size_t)hidden * 4 +
           MG_MAXT * 16 * 4 + 64 * 4 + 2 * MG_NSLOT * 8 + 128;
}

template <int D, int REP>
static int launch_mega_t(const MegaParams& p, int grid, cudaStream_t stream) {
    const size_t smem = mega_smem_bytes(p.hidden);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(stack_pass_kernel<D, REP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        attr_set = true;
    }
    Q3T_REQUIRE(smem <= 227 * 1024, "stack_pass: shared memory budget exceeded");
    cudaMemsetAsync(p.bar, 0, sizeof(unsigned int), stream);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(MG_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, stack_pass_kernel<D, REP>, p);
    Q3T_CHECK_LAUNCH("stack_pass");
    return 0;
}

int num_sms();

int launch_stack_pass(const q3t_stack_pass_args* a, cudaStream_t stream) {
    const q3t_stack& st = a->stack;
    Q3T_REQUIRE(st.layers_dev != nullptr, "stack_pass: layers_dev missing");
    Q3T_REQUIRE(st.hidden % 256 == 0 && st.inter % 256 == 0 && (st.n_heads * st.head_dim) % 256 == 0, "stack_pass: dims % 256");
    Q3T_REQUIRE(st.inter <= MG_MAXK && st.hidden <= MG_MAXK, "stack_pass: K too large");
    const int grid = num_sms();
    const int qkv_n = (st.n_heads + 2 * st.n_kv_heads) * st.head_dim;
    const int maxN = 2 * st.inter > qkv_n ? 2 * st.inter : qkv_n;
    const long long worst = ((long long)(maxN / 16) * (st.hidden / 256) + grid - 1) / grid + 1;
    const long long worst_down = ((long long)(st.hidden / 16) * (st.inter / 256) + grid - 1) / grid + 1;
    Q3T_REQUIRE(worst <= MG_MAXT && worst_down <= MG_MAXT, "stack_pass: too many tiles per CTA");
    Q3T_REQUIRE(st.n_kv_heads * st.attn_nsplit <= grid, "stack_pass: kv_heads * nsplit exceeds the grid");
    {   // a row tile may straddle at most MG_MAXSLOT CTAs
        const int geo[5][2] = {{qkv_n, st.hidden}, {st.hidden, st.n_heads * st.head_dim}, {2 * st.inter, st.hidden},
                               {st.hidden, st.inter}, {a->head.w ? a->head.N : 16, st.hidden}};
        for (int i = 0; i < (a->head.w ? 5 : 4); ++i) {
            const int nkc = geo[i][1] / 256, T = (geo[i][0] / 16) * nkc, G = T < grid ? T : grid;
            const int per = T / G;
            Q3T_REQUIRE((nkc - 1 + per - 1) / per + 1 <= MG_MAXSLOT, "stack_pass: row tile straddles too many CTAs");
        }
    }
    Q3T_REQUIRE(st.head_dim == 128, "stack_pass: head_dim must be 128");
    MegaParams p;
    memset(&p, 0, sizeof(p));
    p.layers = st.layers_dev; p.n_layers = st.n_layers; p.hidden = st.hidden; p.n_heads = st.n_heads; p.n_kv = st.n_kv_heads;
    p.head_dim = st.head_dim; p.inter = st.inter; p.eps = st.eps; p.final_norm = st.final_norm; p.inv_freq = st.inv_freq;
    p.kv_pool = (__nv_bfloat16*)st.kv_pool; p.kv_layer_stride = st.kv_layer_stride_bytes / 2; p.block_tbl = st.block_tbl;
    p.max_pages = st.max_pages; p.pos = a->pos;
    p.has_head = a->head.w != nullptr; p.head = a->head;
    p.x_in = a->x_in; p.hidden_out = a->hidden_out; p.logits_out = a->logits_out;
    float* ws = a->work;
    p.part_qkv = ws; ws += (size_t)MG_MAXSLOT * qkv_n;
    p.part_o = ws; ws += (size_t)MG_MAXSLOT * st.hidden;
    p.part_gu = ws; ws += (size_t)MG_MAXSLOT * 2 * st.inter;
    p.part_down = ws; ws += (size_t)MG_MAXSLOT * st.hidden;
    p.part_head = ws; ws += (size_t)MG_MAXSLOT * (p.has_head ? a->head.N : 0);
    p.attn_out = ws; ws += (size_t)st.n_heads * st.head_dim;
    p.attn_work = ws;
    p.attn_counters = a->counters; p.attn_nsplit = st.attn_nsplit; p.bar = a->barrier; p.timing = a->timing;
    const int rep = st.n_heads / st.n_kv_heads;
    if (rep == 2) return launch_mega_t<128, 2>(p, grid, stream);
    if (rep == 1) return launch_mega_t<128, 1>(p, grid, stream);
    if (rep == 4) return launch_mega_t<128, 4>(p, grid, stream);
    Q3T_REQUIRE(false, "stack_pass: H/Hkv must be 1, 2 or 4");
    return 2;
}

}  // namespace q3t

extern "C" int q3t_stack_pass(const q3t_stack_pass_args* a, void* stream) {
    return q3t::launch_stack_pass(a, (cudaStream_t)stream);
}

extern "C" long long q3t_stack_pass_work_floats(const q3t_stack* st, int head_n) {
    const long long qkv_n = (long long)(st->n_heads + 2 * st->n_kv_heads) * st->head_dim;
    const long long rep = st->n_heads / st->n_kv_heads;
    return q3t::MG_MAXSLOT * (qkv_n + 2LL * st->hidden + 2LL * st->inter + head_n) + (long long)st->n_heads * st->head_dim +
           (long long)st->n_kv_heads * st->attn_nsplit * rep * (st->head_dim + 2) + 64;
}
