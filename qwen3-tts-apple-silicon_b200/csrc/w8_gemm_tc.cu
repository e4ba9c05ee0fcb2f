// W8 (affine, group 64) dequant-fused GEMM on the 5th-generation tensor cores (tcgen05 + TMEM) for sm_100a:
//     Y[m, n] = epilogue( sum_k  X[m, k] * dequant(W8)[n, k] )          M = tokens (batch-64 decode, prefill), N = features
// Replaces mx.quantized_matmul (qmm path: batched decode and prefill) of the reference stack.
//
// Shape of the problem: W is streamed from HBM exactly once per launch.  At M = 64 the kernel is paced by the dequantise -> MMA
// loop (0.7 us per 128 x 64 weight block, tools/gemm_stamps.py) and by ~4.5 us of fixed cost per launch, at prefill sizes by the
// tensor pipe.  W plays the MMA "A" operand (128 output features per CTA = eight 16-row W8 tiles), the activations are the "B"
// operand (up to 128 tokens per CTA), the fp32 accumulator D[128 features x 2 x tokens] lives in TMEM.
//   warp 0      : lane 0 - cp.async.bulk (TMA, no tensor map: W8 tiles are contiguous 4352-byte records) into a 2-stage ring of
//                 raw tiles, one stage = the eight tiles of one 256-wide K chunk, starting BEFORE the previous kernel has
//                 finished (weights do not depend on it); lane 1 - the activation blocks (bf16 split rows) by 2-D TMA tensor map
//                 with hardware SWIZZLE_128B into a ring of their own (5 stages at 64 tokens)
//   warp 1      : allocates TMEM; one elected lane issues tcgen05.mma (kind::f16, bf16 x bf16 -> fp32, M=128, K=16 per
//                 instruction): per K step W_hi x [x_hi | x_lo] (N = 2 x tokens) and W_lo x x_hi (N = tokens), commits to mbarriers
//   warps 2..17 : (a) dequantise one 64-wide K block of the raw stage into the bf16 K-major SWIZZLE_128B operand layout, hi and lo
//                 plane (two warps per 16-row tile; codes are fragment-ordered in HBM: byte -> fp32 by the 2^23 trick, one FFMA
//                 with the group scale/bias, cvt.rn.bf16x2, 8-byte stores), fence.proxy.async + mbarrier arrive - also ahead of
//                 the grid dependency; (b) epilogue: warps 2..5 tcgen05.ld both accumulator halves -> staging tile; split-K =
//                 the CTAs of one output tile are a thread-block CLUSTER, every rank sums its token slice over the ranks'
//                 staging tiles through distributed shared memory (rank order: deterministic); then bias / SiLU / SwiGLU pair /
//                 residual / deferred-RMSNorm scale and statistics -> coalesced fp32 or split-bf16 stores.
// Numerics: SPLIT-bf16 operands, fp32 accumulation.  A 28-layer random-init stack amplifies a single bf16 rounding of the
// operands (2^-9) to 1.4e-2 of max|logit| (measured against the fp32 CPU oracle at the 1.7B shapes; weights alone 1.7e-2,
// TF32 operands 6e-3) - outside the 1e-2 BASELINE.json allows.  So every operand is carried as hi + lo, both bf16
// (hi = rn(v), lo = rn(v - hi): 16 mantissa bits), and a K block costs three products: W_hi.x_hi + W_lo.x_hi + W_hi.x_lo (the
// lo.lo term is 2^-18 relative) issued as two instructions per K step.  Prompt GEMMs pay 3x the MACs.  Activation rows therefore travel between kernels as "split rows" [hi(K) | lo(K)] of 2K bf16 (the same
// bytes as the fp32 row).  The batch-1/2 GEMV (w8_gemv.cu, frame_ll.cu) stays exact-integer.
#include <cuda.h>
#include "common.cuh"
#include "../../include/q3tts_b200.h"

namespace q3t {

constexpr int TC_DQ_WARPS = 16;                 // two warps per 16-row weight tile (one 32-wide K half each)
constexpr int TC_THREADS = 64 + TC_DQ_WARPS * 32;   // 18 warps
constexpr int TC_BM = 128;                      // features per CTA (MMA M)
constexpr int TC_BN_MAX = 128;                  // tokens per CTA (MMA N)
constexpr int TC_BK = 64;                       // K per operand stage (= one quantisation group = one 128-byte swizzle row)
// Two rings.  The weight operand (dequantised in the kernel) needs two stages: block kb+1 is converted while block kb is in the
// tensor pipe.  The activation operand comes by TMA with ~1.5 us of latency (measured, TC_TIMING): it gets every stage the
// budget leaves - five at 64 tokens - so its loads run far ahead of the MMAs instead of pacing them (one shared 3-stage
// ring: 0.7 us per K block, of which the tensor pipe needs 0.45).
constexpr int TC_A_STAGES = 2;
constexpr int TC_B_STAGES_MAX = 8;
constexpr int TC_RAW_STAGES = 2;
constexpr int TC_RAW_BYTES = 8 * Q3T_TILE_BYTES;            // 34 816
constexpr int TC_A_HALF = TC_BM * 128;                       // 16 384: one bf16 plane (hi or lo) of the weight block
constexpr int TC_A_BYTES = 2 * TC_A_HALF;                    // hi plane, then lo plane
constexpr int TC_OFF_A = 0;
// operand rings: TC_A_STAGES stages of {A hi | A lo} (32 KB each), then nsb stages of {B hi | B lo} (2 * bn * 128 bytes each):
// 128 tokens -> 2 stages (64 + 64 KB), 64 tokens (decode) -> 5 stages (64 + 80 KB)
constexpr int TC_AB_BUDGET = 147456;
constexpr int TC_OFF_RAW = TC_AB_BUDGET;
constexpr int TC_OFF_BAR = TC_OFF_RAW + TC_RAW_STAGES * TC_RAW_BYTES;   // 217 088
constexpr int TC_SMEM_BYTES = TC_OFF_BAR + 256 + 1024;      // + barriers + slack for the 1024-byte alignment of the base

struct GemmParams {
    const uint8_t* w; int N, K;                 // W8 tiles [N/16][K/256]
    const __nv_bfloat16* xb; long long xb_stride;   // activations, split rows [M, 2K] bf16: hi(K) | lo(K)
    int M;
    const float* lin_bias; int act; int swiglu;     // epilogue
    const float* resid; long long resid_stride;
    float* y; long long y_stride;
    __nv_bfloat16* yb;                          // optional: split bf16 rows [M, 2N] ([M, 2(N/2)] with swiglu) instead of y
    int bn;                                     // tokens per CTA (multiple of 16, <= 128)
    int nsb;                                    // activation stages (2 .. TC_B_STAGES_MAX, see TC_AB_BUDGET)
    int splits;                                 // split-K factor (gridDim.z = cluster size along z)
    float* ws; int* counters;                   // unused by the kernel since the cluster reduction (ws + ws_floats: TC_TIMING stamps)
    // deferred RMSNorm (q3t_gemm_args.y_norm_w / x_rowss): producer side writes y AND split rows of y * ynw + row statistics,
    // consumer side scales the contraction of row m by rsqrt(sum(xss[m][:]) / K + eps)
    const float* ynw; float* yss;
    const float* xss; int xss_parts; float eps;
    unsigned long long* stamps;                 // TC_TIMING builds only
};

// ---- PTX wrappers --------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void tc_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "TC_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra TC_DONE;\n"
        "bra TC_WAIT;\n"
        "TC_DONE:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_tma_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
        ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major, SWIZZLE_128B shared-memory operand descriptor: 8-row groups are 1024 bytes apart (SBO), version 1 (sm_100)
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// four fp32 values -> hi plane (rn to bf16) and lo plane (rn of what the hi plane dropped)
__device__ __forceinline__ void split_bf16x4(const float4 v, uint2& hi, uint2& lo) {
    hi.x = pack_bf16x2(v.x, v.y); hi.y = pack_bf16x2(v.z, v.w);
    lo.x = pack_bf16x2(v.x - __uint_as_float(hi.x << 16), v.y - __uint_as_float(hi.x & 0xffff0000u));
    lo.y = pack_bf16x2(v.z - __uint_as_float(hi.y << 16), v.w - __uint_as_float(hi.y & 0xffff0000u));
}

#ifdef TC_TIMING
// stamps of the CTAs (0, 0, z) go BEHIND the split-K workspace: [z][16] u64 at ws + ws_floats (tools/gemm_stamps.py)
#define TC_STAMP(i) do { if (blockIdx.x == 0 && blockIdx.y == 0 && p.stamps) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); p.stamps[blockIdx.z * 16 + (i)] = t_; } } while (0)
// per-K-block detail of CTA (0, 0, 0), first 16 blocks: [128 + kb * 8 + i] (i: 0 dequant warp sees the stage free, 1 raw chunk landed,
// 2 operand stored + fenced, 3 MMA thread sees the stage full, 4 MMAs issued + committed, 5 activation TMA issued)
#define TC_STAMP_KB(kb, i) do { if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && p.stamps && (kb) < 16) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); p.stamps[128 + (kb) * 8 + (i)] = t_; } } while (0)
#else
#define TC_STAMP(i) do { } while (0)
#define TC_STAMP_KB(kb, i) do { } while (0)
#endif

// ---- the kernel ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1) w8_gemm_tc_kernel(const GemmParams p, const __grid_constant__ CUtensorMap tmap_x) {
    extern __shared__ unsigned char tc_smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)tc_smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TC_OFF_BAR);
    uint64_t* full_a = bars;                         // [TC_A_STAGES]  weight block dequantised (16 warp arrivals)
    uint64_t* empty_a = full_a + TC_A_STAGES;        // [TC_A_STAGES]  MMA done with the stage (tcgen05.commit)
    uint64_t* full_b = empty_a + TC_A_STAGES;        // [TC_B_STAGES_MAX] activation block landed (TMA)
    uint64_t* empty_b = full_b + TC_B_STAGES_MAX;    // [TC_B_STAGES_MAX] MMA done with the stage
    uint64_t* full_raw = empty_b + TC_B_STAGES_MAX;  // [TC_RAW_STAGES] TMA landed
    uint64_t* empty_raw = full_raw + TC_RAW_STAGES;  // [TC_RAW_STAGES] 16 warp arrivals
    uint64_t* tmem_full = empty_raw + TC_RAW_STAGES; // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) TC_STAMP(0);
    const int nb = blockIdx.x, m0 = blockIdx.y * p.bn;
    const int bn = p.bn, nsb = p.nsb;
    const int b_half = bn * 128, b_bytes = 2 * b_half, off_b = TC_A_STAGES * TC_A_BYTES;
    // split-K: this CTA contracts the 256-wide chunks [kc_lo, kc_hi)
    const int nkc_all = p.K >> 8;
    const int kc_lo = (nkc_all * (int)blockIdx.z) / p.splits, kc_hi = (nkc_all * ((int)blockIdx.z + 1)) / p.splits;
    const int nkc = kc_hi - kc_lo, nkb = nkc * 4;

    if (tid == 0) {
        for (int i = 0; i < TC_A_STAGES; ++i) { tc_mbar_init(tc_smem_u32(&full_a[i]), TC_DQ_WARPS); tc_mbar_init(tc_smem_u32(&empty_a[i]), 1); }
        for (int i = 0; i < TC_B_STAGES_MAX; ++i) { tc_mbar_init(tc_smem_u32(&full_b[i]), 1); tc_mbar_init(tc_smem_u32(&empty_b[i]), 1); }
        for (int i = 0; i < TC_RAW_STAGES; ++i) { tc_mbar_init(tc_smem_u32(&full_raw[i]), 1); tc_mbar_init(tc_smem_u32(&empty_raw[i]), TC_DQ_WARPS); }
        tc_mbar_init(tc_smem_u32(tmem_full), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(tmem_slot)), "n"(2 * TC_BN_MAX) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;
    pdl_launch_dependents();
    if (tid == 0) TC_STAMP(1);

    if (warp == 0) {
        // ===================== producers: lane 0 raw W8 tiles (weights do not depend on the previous kernel),
        //                       lane 1 activation blocks (2-D tensor map, hardware 128-byte swizzle, OOB rows = 0) =============
        if (lane == 0) {
            for (int kc = 0; kc < nkc; ++kc) {
                const int rs = kc % TC_RAW_STAGES, par = (kc / TC_RAW_STAGES) & 1;
                tc_mbar_wait(tc_smem_u32(&empty_raw[rs]), par ^ 1);
                const uint32_t fb = tc_smem_u32(&full_raw[rs]);
                tc_mbar_expect_tx(fb, TC_RAW_BYTES);
                unsigned char* dst = smem + TC_OFF_RAW + rs * TC_RAW_BYTES;
                for (int rt = 0; rt < 8; ++rt)
                    tc_tma_load_1d(tc_smem_u32(dst + rt * Q3T_TILE_BYTES),
                                   p.w + ((size_t)(nb * 8 + rt) * nkc_all + kc_lo + kc) * Q3T_TILE_BYTES, Q3T_TILE_BYTES, fb);
            }
        } else if (lane == 1) {
            pdl_wait();                           // the bf16 activations come from the previous kernel
            TC_STAMP(2);
            for (int kb = 0; kb < nkb; ++kb) {
                const int st = kb % nsb, par = (kb / nsb) & 1;
                tc_mbar_wait(tc_smem_u32(&empty_b[st]), par ^ 1);
                TC_STAMP_KB(kb, 5);
                const uint32_t fb = tc_smem_u32(&full_b[st]);
                tc_mbar_expect_tx(fb, (uint32_t)bn * 256u);
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                             ::"r"(tc_smem_u32(smem + off_b + st * b_bytes)), "l"(reinterpret_cast<uint64_t>(&tmap_x)),
                               "r"((kc_lo * 4 + kb) * TC_BK), "r"(m0), "r"(fb) : "memory");
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                             ::"r"(tc_smem_u32(smem + off_b + st * b_bytes + b_half)), "l"(reinterpret_cast<uint64_t>(&tmap_x)),
                               "r"(p.K + (kc_lo * 4 + kb) * TC_BK), "r"(m0), "r"(fb) : "memory");   // lo plane: columns K.. of the split row
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =================================================================================
        if (lane == 0) {
            // Two instructions per 16-wide K step instead of one per product term: the hi and lo planes of the activations are
            // adjacent in the stage, so W_hi x [x_hi | x_lo] is ONE MMA of N = 2 bn tokens into the columns [0, 2 bn) of the
            // accumulator, and W_lo x x_hi adds into columns [0, bn); the epilogue sums the two halves.  At decode sizes the
            // tensor pipe is paced by instruction count and by the shared-memory read of the 128-row weight operand (about
            // 100 cycles per instruction at N = 64, measured with TC_TIMING), not by MACs.
            const uint32_t idesc_n = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
            const uint32_t idesc_w = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 2) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
            for (int kb = 0; kb < nkb; ++kb) {
                const int st = kb % TC_A_STAGES, par = (kb / TC_A_STAGES) & 1, sb = kb % nsb, parb = (kb / nsb) & 1;
                tc_mbar_wait(tc_smem_u32(&full_b[sb]), parb);
                tc_mbar_wait(tc_smem_u32(&full_a[st]), par);
                if (kb == 0) TC_STAMP(3);
                TC_STAMP_KB(kb, 3);
                tc_fence_after();
                const uint64_t a_hi = tc_smem_desc(tc_smem_u32(smem + TC_OFF_A + st * TC_A_BYTES));
                const uint64_t a_lo = tc_smem_desc(tc_smem_u32(smem + TC_OFF_A + st * TC_A_BYTES + TC_A_HALF));
                const uint64_t b_hl = tc_smem_desc(tc_smem_u32(smem + off_b + sb * b_bytes));     // bn rows of hi, then bn rows of lo
#pragma unroll
                for (int k = 0; k < TC_BK / 16; ++k) {    // 16 bf16 = 32 bytes = 2 descriptor units along K
                    tc_mma_bf16(tmem_d, a_hi + 2 * k, b_hl + 2 * k, idesc_w, (kb | k) != 0);
                    tc_mma_bf16(tmem_d, a_lo + 2 * k, b_hl + 2 * k, idesc_n, 1u);
                }
                tc_commit(tc_smem_u32(&empty_a[st]));      // frees the stages once these MMAs have read them
                tc_commit(tc_smem_u32(&empty_b[sb]));
                TC_STAMP_KB(kb, 4);
            }
            tc_commit(tc_smem_u32(tmem_full));              // accumulator complete
            TC_STAMP(4);
        }
    } else {
        // ===================== dequant + activation copy (8 warps), then epilogue (first 4 of them) =======================
        const int dw = (warp - 2) >> 1;           // row tile of this warp pair inside the 128-row block
        const int jh = (warp - 2) & 1;            // which 32-wide half of the 64-wide K block this warp converts
        const int g = lane >> 2, t = lane & 3;
        // (no griddepcontrol.wait here: the weight operand does not depend on the previous kernel, so the first stages are
        // dequantised under its tail; these warps wait right before the epilogue, which reads residual rows and row statistics)
        for (int kb = 0; kb < nkb; ++kb) {
            const int st = kb % TC_A_STAGES, par = (kb / TC_A_STAGES) & 1;
            const int kc = kb >> 2, j4 = kb & 3, rs = kc % TC_RAW_STAGES, rpar = (kc / TC_RAW_STAGES) & 1;
            tc_mbar_wait(tc_smem_u32(&empty_a[st]), par ^ 1);
            if (tid == 64) TC_STAMP_KB(kb, 0);
            // ---- A operand: dequantise group j4 of this warp's raw tile
            if (j4 == 0) tc_mbar_wait(tc_smem_u32(&full_raw[rs]), rpar);
            if (tid == 64) TC_STAMP_KB(kb, 1);
            {
                const unsigned char* tile = smem + TC_OFF_RAW + rs * TC_RAW_BYTES + dw * Q3T_TILE_BYTES;
                // scales / biases of rows g and g+8 for group j4: meta row = {4 bf16 scales, 4 bf16 biases}
                const uint16_t* mlo = reinterpret_cast<const uint16_t*>(tile + 4096 + g * 16);
                const uint16_t* mhi = reinterpret_cast<const uint16_t*>(tile + 4096 + (g + 8) * 16);
                const float s_lo = __uint_as_float((uint32_t)mlo[j4] << 16), b_lo = __uint_as_float((uint32_t)mlo[4 + j4] << 16);
                const float s_hi = __uint_as_float((uint32_t)mhi[j4] << 16), b_hi = __uint_as_float((uint32_t)mhi[4 + j4] << 16);
                unsigned char* adst = smem + TC_OFF_A + st * TC_A_BYTES + dw * 2048;      // 16 rows = two 8-row swizzle atoms
                {
                    const int j = jh;
                    const uint4 q = *reinterpret_cast<const uint4*>(tile + (j4 * 2 + j) * 512 + lane * 16);
                    const uint32_t qq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        // register i: row g + 8*(i&1), k = 32 j + 16 (i>>1) + 4 t + {0..3} inside the group
                        const float sc = (i & 1) ? s_hi : s_lo, bi = (i & 1) ? b_hi : b_lo;
                        float f[4];
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
                            const float qf = __uint_as_float(__byte_perm(qq[i], 0x4B000000u, 0x7650 + b)) - 8388608.f;
                            f[b] = fmaf(sc, qf, bi);
                        }
                        const int row8 = g;                       // row inside its 8-row atom
                        const int chunk = 4 * j + 2 * (i >> 1) + (t >> 1);
                        uint2 o, ol;
                        o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
                        // lo plane: what the bf16 rounding of the hi plane dropped (exact in fp32, then rounded once)
                        ol.x = pack_bf16x2(f[0] - __uint_as_float(o.x << 16), f[1] - __uint_as_float(o.x & 0xffff0000u));
                        ol.y = pack_bf16x2(f[2] - __uint_as_float(o.y << 16), f[3] - __uint_as_float(o.y & 0xffff0000u));
                        unsigned char* dsta = adst + (i & 1) * 1024 + row8 * 128 + ((chunk ^ row8) << 4) + (t & 1) * 8;
                        *reinterpret_cast<uint2*>(dsta) = o;
                        *reinterpret_cast<uint2*>(dsta + TC_A_HALF) = ol;
                    }
                }
            }
            tc_fence_proxy_async();               // generic-proxy smem writes -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) {
                tc_mbar_arrive(tc_smem_u32(&full_a[st]));
                if (j4 == 3) tc_mbar_arrive(tc_smem_u32(&empty_raw[rs]));
            }
            if (tid == 64) TC_STAMP_KB(kb, 2);
        }
    }

    // ---- epilogue.  Warps 2..5 own TMEM lanes 32*(warp%4)..+31 (= features of this CTA): they move the accumulator (both
    // halves summed) to shared memory (the raw-tile ring is free by now) as [token][feature]; then ALL 16 warps of this group
    // write token rows with float4 stores.  One store per (feature, token) from four warps - the obvious epilogue - spends
    // ~25 dependent 64-bit address instructions per element on a single warp per scheduler (0.1 us per token, measured).
    // Split-K: the CTAs (nb, mb, 0..splits-1) are ONE thread-block cluster.  Each stages its partial tile, the cluster
    // synchronises, and CTA z finishes the tokens [nt z / splits, nt (z+1) / splits): it sums the partial rows of all ranks
    // through distributed shared memory in rank order (deterministic, no float atomics, no trip through L2) and applies the
    // epilogue.  (The workspace + arrival counter + last-CTA reduction this replaces cost 9 us per launch: 256 KB of partials
    // pulled through L2 by a single SM.)
    constexpr int STG_LD = TC_BM + 4;                          // padded row: 528 bytes, 16-byte aligned
    constexpr int STG_TOK = TC_BN_MAX;                         // 128 x 132 x 4 = 67 584 B
    static_assert(STG_TOK * STG_LD * 4 + STG_TOK * 4 <= TC_RAW_STAGES * TC_RAW_BYTES, "staging tile + row scales must fit the raw ring");
    float* stg = reinterpret_cast<float*>(smem + TC_OFF_RAW);
    float* rstd_s = stg + STG_TOK * STG_LD;                     // [STG_TOK] row scales of the deferred RMSNorm
    const int dt = tid - 64;                                   // 0..511 for the epilogue warps
    const int n0 = nb * TC_BM, nt = bn;
    if (warp >= 2) {
        pdl_wait();                               // residual rows / row statistics / the rows this epilogue overwrites
        if (warp < 6) {
            tc_mbar_wait(tc_smem_u32(tmem_full), 0);
            if (tid == 64) TC_STAMP(5);
            tc_fence_after();
            const int lg = warp & 3, f_local = lg * 32 + lane;
            const uint32_t tbase = tmem_d + ((uint32_t)(lg * 32) << 16);
#pragma unroll 1
            for (int c0 = 0; c0 < nt; c0 += 8) {
                uint32_t v[8], w[8];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                             : "r"(tbase + (uint32_t)c0));
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                             : "r"(tbase + (uint32_t)(bn + c0)));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                float* d = stg + c0 * STG_LD + f_local;
#pragma unroll
                for (int c = 0; c < 8; ++c) d[c * STG_LD] = __uint_as_float(v[c]) + __uint_as_float(w[c]);     // (hi.hi + lo.hi) + hi.lo
            }
            tc_fence_before();
        }
        if (p.xss && dt < nt) {
            // consumer half of the deferred RMSNorm: the statistics of token m, summed in block order
            const int m = m0 + dt;
            float ss = 0.f;
            if (m < p.M) for (int j = 0; j < p.xss_parts; ++j) ss += __ldcg(p.xss + (size_t)m * p.xss_parts + j);
            rstd_s[dt] = rsqrtf(ss / (float)p.K + p.eps);
        }
        asm volatile("bar.sync 3, 512;" ::: "memory");
        if (tid == 64) TC_STAMP(8);
    }
    const int S = p.splits;
    uint32_t rank = 0;
    if (S > 1) {
        asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
        asm volatile("barrier.cluster.arrive.release;\nbarrier.cluster.wait.acquire;" ::: "memory");     // every partial tile is staged
        if (tid == 64) TC_STAMP(9);
    }
    if (warp >= 2) {
        const int tok_lo = (nt * (int)rank) / S, tok_hi = (nt * ((int)rank + 1)) / S;
        const uint32_t stg_u32 = tc_smem_u32(stg);
        // partial rows of rank z: the same shared-memory offset in CTA z of the cluster
        auto ld_rank = [&](int z, int float_off) -> float4 {
            uint32_t ra;
            float4 r;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(stg_u32 + (uint32_t)float_off * 4u), "r"(z));
            asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(ra) : "memory");
            return r;
        };
        auto tile4 = [&](int float_off) -> float4 {
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            float4 a[8];
#pragma unroll
            for (int z = 0; z < 8; ++z) if (z < S) a[z] = ld_rank(z, float_off);        // all loads in flight, then the sum in rank order
#pragma unroll
            for (int z = 0; z < 8; ++z) if (z < S) { x.x += a[z].x; x.y += a[z].y; x.z += a[z].z; x.w += a[z].w; }
            return x;
        };
        if (p.swiglu) {
            // fused gate/up rows: gate rows at 16j..16j+7, matching up rows at 16j+8..16j+15 -> N/2 activations per token
            const int n_items = (tok_hi - tok_lo) * 16;
            // split-K: a rank owns at most 64 tokens = 2 items per thread; their partial rows are summed into registers first so
            // that the cluster can be told early that this CTA is done reading its peers (see the wait at the end)
            float4 sg[2], su[2];
            if (S > 1) {
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int i = dt + q * 512;
                    if (i < n_items) {
                        const int off = (tok_lo + (i >> 4)) * STG_LD + 16 * ((i & 15) >> 1) + 4 * (i & 1);
                        sg[q] = tile4(off); su[q] = tile4(off + 8);
                    }
                }
                asm volatile("barrier.cluster.arrive.release;" ::: "memory");
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int i = dt + q * 512;
                if (i >= n_items) break;
                const int tok = tok_lo + (i >> 4), j = (i & 15) >> 1, h = i & 1, m = m0 + tok;
                if (m >= p.M) continue;
                const int fg = 16 * j + 4 * h;                   // gate features fg..fg+3, up features fg+8..fg+11
                float4 g, u;
                if (S > 1) { g = sg[q < 2 ? q : 0]; u = su[q < 2 ? q : 0]; }
                else { g = *reinterpret_cast<const float4*>(stg + tok * STG_LD + fg); u = *reinterpret_cast<const float4*>(stg + tok * STG_LD + fg + 8); }
                if (p.xss) {
                    const float rs = rstd_s[tok];
                    g.x *= rs; g.y *= rs; g.z *= rs; g.w *= rs; u.x *= rs; u.y *= rs; u.z *= rs; u.w *= rs;
                }
                if (p.lin_bias) {
                    const float4 bg = *reinterpret_cast<const float4*>(p.lin_bias + n0 + fg), bu = *reinterpret_cast<const float4*>(p.lin_bias + n0 + fg + 8);
                    g.x += bg.x; g.y += bg.y; g.z += bg.z; g.w += bg.w; u.x += bu.x; u.y += bu.y; u.z += bu.z; u.w += bu.w;
                }
                const float4 o = make_float4(silu_f(g.x) * u.x, silu_f(g.y) * u.y, silu_f(g.z) * u.z, silu_f(g.w) * u.w);
                if (p.yb) {
                    uint2 ob, ol; split_bf16x4(o, ob, ol);
                    __nv_bfloat16* yr = p.yb + (size_t)m * p.N + (n0 >> 1) + 8 * j + 4 * h;      // split row: 2 * (N/2) wide
                    *reinterpret_cast<uint2*>(yr) = ob;
                    *reinterpret_cast<uint2*>(yr + (p.N >> 1)) = ol;
                } else {
                    *reinterpret_cast<float4*>(p.y + (size_t)m * p.y_stride + (n0 >> 1) + 8 * j + 4 * h) = o;
                }
            }
        } else {
            // residual rows of all of this thread's items first (at most 8 float4): issued together, they cost one
            // memory round trip instead of one per item between the stores
            const int n_items = (tok_hi - tok_lo) * 32;
            float4 rres[8];
            if (p.resid) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int i = dt + q * 512, tok = tok_lo + (i >> 5), q4 = i & 31, m = m0 + tok;
                    if (i < n_items && m < p.M) rres[q] = *reinterpret_cast<const float4*>(p.resid + (size_t)m * p.resid_stride + n0 + 4 * q4);
                }
            }
            float4 sx[4];                                   // split-K: at most 64 tokens = 4 items per thread
            if (S > 1) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int i = dt + q * 512;
                    if (i < n_items) sx[q] = tile4((tok_lo + (i >> 5)) * STG_LD + 4 * (i & 31));
                }
                asm volatile("barrier.cluster.arrive.release;" ::: "memory");
            }
#pragma unroll
            for (int qq = 0; qq < 8; ++qq) {
                const int i = dt + qq * 512;
                if (i >= n_items) break;
                const int tok = tok_lo + (i >> 5), q4 = i & 31, m = m0 + tok;
                if (m >= p.M) continue;
                float4 x = (S > 1) ? sx[qq < 4 ? qq : 0] : *reinterpret_cast<const float4*>(stg + tok * STG_LD + 4 * q4);
                if (p.xss) { const float rs = rstd_s[tok]; x.x *= rs; x.y *= rs; x.z *= rs; x.w *= rs; }
                if (p.lin_bias) {
                    const float4 b = *reinterpret_cast<const float4*>(p.lin_bias + n0 + 4 * q4);
                    x.x += b.x; x.y += b.y; x.z += b.z; x.w += b.w;
                }
                if (p.act == Q3T_ACT_SILU) { x.x = silu_f(x.x); x.y = silu_f(x.y); x.z = silu_f(x.z); x.w = silu_f(x.w); }
                if (p.resid) { const float4 r = rres[qq]; x.x += r.x; x.y += r.y; x.z += r.z; x.w += r.w; }
                if (p.ynw) {
                    // producer half of the deferred RMSNorm: fp32 row (the residual stream), split row of y * w for the
                    // next GEMM, and this block's share of sum(y^2) (the 32 lanes of a warp hold one token's 128 features)
                    *reinterpret_cast<float4*>(p.y + (size_t)m * p.y_stride + n0 + 4 * q4) = x;
                    const float ss = warp_sum(x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w);
                    if (q4 == 0) p.yss[(size_t)m * gridDim.x + nb] = ss;
                    const float4 w4 = *reinterpret_cast<const float4*>(p.ynw + n0 + 4 * q4);
                    x.x *= w4.x; x.y *= w4.y; x.z *= w4.z; x.w *= w4.w;
                }
                if (p.yb) {
                    uint2 ob, ol; split_bf16x4(x, ob, ol);
                    __nv_bfloat16* yr = p.yb + (size_t)m * 2 * p.N + n0 + 4 * q4;
                    *reinterpret_cast<uint2*>(yr) = ob;
                    *reinterpret_cast<uint2*>(yr + p.N) = ol;
                } else {
                    *reinterpret_cast<float4*>(p.y + (size_t)m * p.y_stride + n0 + 4 * q4) = x;
                }
            }
        }
        if (tid == 64) TC_STAMP(6);
    }
    if (S > 1) {
        // nobody leaves while a peer may still read its tile: the epilogue warps arrived right after their last remote load
        if (warp < 2) asm volatile("barrier.cluster.arrive.release;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) TC_STAMP(7);
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(2 * TC_BN_MAX) : "memory");
    }
}

// ---- activation preparation: fp32 rows -> (RMSNorm | SwiGLU of interleaved gate/up) -> bf16 rows ------------------------
struct PrepParams {
    const float* x; long long x_stride; int M, K, prologue;
    const float* norm_w; float eps;
    const int* gather_idx; int gather_idx_stride; long long gather_row_stride;
    __nv_bfloat16* out; long long out_stride;
};

__global__ void __launch_bounds__(256) act_prep_kernel(const PrepParams p) {
    __shared__ float red[32];
    pdl_wait();
    pdl_launch_dependents();
    const int m = blockIdx.x, tid = threadIdx.x;
    const float* xr = p.x + (size_t)m * p.x_stride;
    if (p.gather_idx) xr = p.x + (long long)p.gather_idx[(size_t)m * p.gather_idx_stride] * p.gather_row_stride;
    __nv_bfloat16* o = p.out + (size_t)m * p.out_stride;
    float rstd = 1.f;
    if (p.prologue == Q3T_PRO_RMSNORM) {
        float ss = 0.f;
        for (int k = tid; k < p.K; k += 256) { const float v = xr[k]; ss += v * v; }
        rstd = rsqrtf(block_sum(ss, red) / (float)p.K + p.eps);
    }
    for (int k = tid; k < p.K; k += 256) {
        float v;
        if (p.prologue == Q3T_PRO_SWIGLU) {
            const int gi = ((k >> 3) << 4) + (k & 7);
            v = silu_f(xr[gi]) * xr[gi + 8];
        } else if (p.prologue == Q3T_PRO_RMSNORM) {
            v = p.norm_w[k] * (xr[k] * rstd);
        } else {
            v = xr[k];
        }
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        o[k] = hi;
        o[p.K + k] = __float2bfloat16_rn(v - __bfloat162float(hi));      // lo plane of the split row
    }
}

int launch_w8_gemm(const q3t_gemm_args* a, cudaStream_t stream) {
    Q3T_REQUIRE(a->M >= 1, "w8_gemm: M must be >= 1");
    Q3T_REQUIRE(a->w.N % 128 == 0 && a->w.K % 256 == 0, "w8_gemm: N%128 / K%256");
    Q3T_REQUIRE(a->xb != nullptr || a->x_bf16 != nullptr, "w8_gemm: bf16 activation scratch missing");
    Q3T_REQUIRE(!a->x_bf16 || (a->prologue == Q3T_PRO_RAW && !a->gather_idx), "w8_gemm: x_bf16 excludes prologue / gather");
    const int K = a->w.K;
    // 1. activations -> bf16 (with the fused prologue), unless the producer kernel already wrote bf16 rows
    PrepParams pp;
    memset(&pp, 0, sizeof(pp));
    pp.x = a->x; pp.x_stride = a->x_stride; pp.M = a->M; pp.K = K; pp.prologue = a->prologue; pp.norm_w = a->norm_w; pp.eps = a->eps;
    pp.gather_idx = a->gather_idx; pp.gather_idx_stride = a->gather_idx_stride; pp.gather_row_stride = a->gather_row_stride;
    pp.out = (__nv_bfloat16*)a->xb; pp.out_stride = 2 * K;
    if (!a->x_bf16) {
        launch_pdl(act_prep_kernel, dim3(a->M), dim3(256), 0, stream, pp);
        Q3T_CHECK_LAUNCH("act_prep");
    }
    // 2. the GEMM
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.w = (const uint8_t*)a->w.w; p.N = a->w.N; p.K = K; p.xb = (const __nv_bfloat16*)(a->x_bf16 ? a->x_bf16 : a->xb); p.xb_stride = 2 * K; p.M = a->M;
    p.lin_bias = a->w.lin_bias; p.act = a->act; p.swiglu = a->swiglu_out; p.resid = a->resid; p.resid_stride = a->resid_stride;
    p.y = a->y; p.y_stride = a->y_stride; p.yb = (__nv_bfloat16*)a->y_bf16;
    Q3T_REQUIRE(!a->y_norm_w || (a->y && a->y_bf16 && a->y_rowss && !a->swiglu_out), "w8_gemm: y_norm_w needs y, y_bf16, y_rowss and no swiglu_out");
    Q3T_REQUIRE(!a->x_rowss || (a->x_bf16 && a->x_rowss_parts >= 1), "w8_gemm: x_rowss needs x_bf16 and x_rowss_parts");
    p.ynw = a->y_norm_w; p.yss = a->y_rowss; p.xss = a->x_rowss; p.xss_parts = a->x_rowss_parts; p.eps = a->eps;
    int bn = (a->M + 15) / 16 * 16;
    if (bn > TC_BN_MAX) bn = TC_BN_MAX;
    if (bn < 16) bn = 16;
    p.bn = bn;
    p.nsb = (TC_AB_BUDGET - TC_A_STAGES * TC_A_BYTES) / (2 * bn * 128);
    if (p.nsb > TC_B_STAGES_MAX) p.nsb = TC_B_STAGES_MAX;
    // split-K for decode-sized problems: fill the machine when N/128 CTAs would leave most SMs idle
    p.splits = 1; p.ws = (float*)a->splitk_ws; p.counters = a->splitk_counters;
#ifdef TC_TIMING
    p.stamps = a->splitk_ws ? reinterpret_cast<unsigned long long*>(a->splitk_ws + a->splitk_ws_floats) : nullptr;
#endif
    {
        const int nkc = K / 256, nblk = p.N / TC_BM;
        // (splitk_ws != NULL is the caller's opt-in; the partial tiles themselves stay in shared memory - see the epilogue)
        if (a->splitk_ws && a->M <= bn) {
            int best = 1;
            for (int d = 2; d <= 8 && d <= nkc; ++d)          // one wave of clusters of d CTAs (portable cluster size <= 8)
                if (nkc % d == 0 && nblk * d <= 148) best = d;
            p.splits = best;
        }
    }
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(w8_gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
        attr_set = true;
    }
    static_assert(TC_SMEM_BYTES <= 227 * 1024, "w8_gemm: shared memory budget exceeded");
    // 2-D tensor map of the split bf16 activations [M, 2K]: box = 64 K-elements (128 bytes) x bn tokens, SWIZZLE_128B, zero fill
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
            snprintf(g_err, sizeof(g_err), "w8_gemm: cuTensorMapEncodeTiled is not available");
            return 3;
        }
        encode = (EncodeFn)fn;
    }
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {(cuuint64_t)2 * K, (cuuint64_t)a->M};          // split rows: hi(K) | lo(K)
    const cuuint64_t gstride[1] = {(cuuint64_t)K * 4};
    const cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)bn};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)p.xb, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { snprintf(g_err, sizeof(g_err), "w8_gemm: cuTensorMapEncodeTiled failed (%d)", (int)cr); return 3; }
    {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(p.N / TC_BM, (a->M + bn - 1) / bn, p.splits); cfg.blockDim = dim3(TC_THREADS);
        cfg.dynamicSmemBytes = (size_t)TC_SMEM_BYTES; cfg.stream = stream;
        cudaLaunchAttribute attr[2];
        int na = 0;
        if (g_use_pdl) { attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[na].val.programmaticStreamSerializationAllowed = 1; ++na; }
        if (p.splits > 1) {          // the splits of one output tile form a cluster (distributed-shared-memory reduction)
            attr[na].id = cudaLaunchAttributeClusterDimension;
            attr[na].val.clusterDim.x = 1; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = (unsigned)p.splits; ++na;
        }
        cfg.attrs = attr; cfg.numAttrs = na;
        cudaLaunchKernelEx(&cfg, w8_gemm_tc_kernel, p, tmap);
    }
    Q3T_CHECK_LAUNCH("w8_gemm_tc");
    return 0;
}

}  // namespace q3t

extern "C" int q3t_w8_gemm(const q3t_gemm_args* a, void* stream) { return q3t::launch_w8_gemm(a, (cudaStream_t)stream); }
