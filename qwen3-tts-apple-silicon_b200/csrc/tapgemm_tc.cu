// Causal tap-GEMM of the speech-tokenizer decoder on the 5th-generation tensor cores (tcgen05, kind::tf32, TMEM):
// conv1d / conv_transpose1d / linear of the codec as an IMPLICIT GEMM.
//     out[b, t, n] = epi( sum_{tap} sum_{ci} A[b, t + shift_tap, ci] * W[tap][n][ci] ),   rows outside [0, T_in) read as zero
// Replaces mx.conv1d / mx.conv_transpose1d / matmul of the reference stack for every layer with Cin % 32 == 0 (everything
// but the 96->1 output conv), with the same fused epilogues as the FP32-pipe kernel in codec.cu.
//
//   * no im2col and no padded copies: the activation operand of tap j is the SAME [B, T, Cin] tensor fetched through a 3-D
//     TMA tensor map at time coordinate t0 + shift_j; rows before 0 / past T_in are the tensor map's zero fill, which IS the
//     causal left padding (and the right trim of the transposed convs);
//   * operands stay fp32 in HBM/L2 (the codec is fp32 end to end); the tensor core reads them as TF32 (kind::tf32, 10-bit
//     mantissa, fp32 accumulate in TMEM).  Weights are rounded to TF32 once at load and the epilogue rounds what it stores
//     (cvt.rna.tf32), so the hardware's truncation never sees low mantissa bits -> unbiased 2^-11 rounding per operand;
//   * MMA tile 128 rows (time) x 128 columns (channels) x 32 (Cin slice = one 128-byte swizzle row), 2-stage TMA ring,
//     68 KB of shared memory and 320 threads -> THREE CTAs per SM: a tile is a chain launch -> set-up -> first TMA -> MMAs ->
//     accumulator -> staging -> stores -> exit, and what hides it is the number of tiles in flight per SM (3 x 8 epilogue warps
//     measured 3 % faster than 2 x 16 with a 3-stage ring; 3 x 12 spills at 40 registers and is no faster);
//   * epilogue: 4 warps move the accumulator TMEM -> shared memory, then all 8 epilogue warps apply bias / LayerScale /
//     residual / SnakeBeta / GELU / SiLU / SwiGLU-pair and store coalesced rows.
#include <cuda.h>
#include <stdlib.h>
#include <cuda_fp16.h>
#include "common.cuh"
#include "../../include/q3tts_b200.h"

namespace q3t {

constexpr int TT_EPI_WARPS = 8;
constexpr int TT_THREADS = 64 + TT_EPI_WARPS * 32;      // producer warp, MMA warp, 8 epilogue warps
constexpr int TT_BM = 128, TT_BN = 128, TT_BK = 32;
constexpr int TT_STAGES = 2;
constexpr int TT_A_BYTES = TT_BM * 128, TT_B_BYTES = TT_BN * 128;     // 16 KB each
constexpr int TT_STAGE_BYTES = TT_A_BYTES + TT_B_BYTES;
constexpr int TT_OFF_BAR = TT_STAGES * TT_STAGE_BYTES + 2048;          // ring (64 KB) + room for the 67.6 KB staging tile
constexpr int TT_SMEM_BYTES = TT_OFF_BAR + 128 + 1024;                 // + barriers + alignment slack
constexpr int TT_STG_LD = TT_BN + 4;                                   // staging tile [128 rows][132]: 16-byte rows, conflict-free for the 128-bit row writes (lane = row) and reads
static_assert(TT_BM * TT_STG_LD * 4 <= TT_OFF_BAR, "staging tile must fit the operand ring");

struct TapTcParams {
    int B, T_in, Cin, taps, shift[8], N, Cout, rows, tiles_per_item;
    const float* bias; const float* scale; const float* resid;
    float* out_raw; float* out_act; int act; const float* act_a; const float* act_b;
    int f16;          // operands are IEEE fp16 (kind::f16, 64 channels per 128-byte stage row) instead of fp32 read as TF32
    int act_f16;      // out_act is written as fp16 (the operand of an fp16 consumer)
};

__device__ __forceinline__ uint32_t tt_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tt_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void tt_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tt_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "TT_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra TT_DONE;\n"
        "bra TT_WAIT;\n"
        "TT_DONE:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
// long waits (the epilogue warps wait microseconds for an accumulator): try_wait with a suspend-time hint parks the warp in
// hardware until the phase flips instead of re-issuing the probe - 16 to 32 spinning warps per SM otherwise compete for issue
// slots with the warps of the other resident CTA that are doing the epilogue arithmetic
__device__ __forceinline__ void tt_mbar_wait_parked(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "TT_PWAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra TT_PDONE;\n"
        "bra TT_PWAIT;\n"
        "TT_PDONE:\n}\n" ::"r"(bar), "r"(parity), "r"(1000000u) : "memory");
}
__device__ __forceinline__ uint64_t tt_smem_desc(uint32_t saddr) {      // K-major, SWIZZLE_128B, 8-row groups 1024 B apart
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ float tt_gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
// sin(x) for the SnakeBeta epilogue: two-constant Cody-Waite reduction to [-pi, pi], then the SFU's sine (abs error 2^-21.4 on
// that interval).  sinf()'s full-accuracy path costs ~25 instructions per element and the epilogue of every vocoder layer is
// bound by exactly those issue slots (ncu launch lists: halving the bytes a tile pulls did not move the kernel, see
// profiles/r02_ll_experiments.txt); the error left (~5e-7 absolute for |x| < 1e4) is three orders below the fp16 / TF32
// rounding of the stored activation.
__device__ __forceinline__ float tt_sin(float x) {
    const float k = rintf(x * 0.15915494309189535f);                    // x / 2 pi
    float r = fmaf(-k, 6.2831854820251465f, x);                         // 2 pi, high part (fp32(2 pi))
    r = fmaf(-k, -1.7484555e-7f, r);                                    // 2 pi - fp32(2 pi)
    return __sinf(r);
}
// two fp32 -> packed fp16x2, round to nearest, saturating at +-65504 (one F2FP.SATFINITE instead of four min / max + convert)
__device__ __forceinline__ uint32_t tt_pack_h2_sat(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ float tt_round_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

__global__ void __launch_bounds__(TT_THREADS, 3) tapgemm_tc_kernel(const TapTcParams p, const __grid_constant__ CUtensorMap tm_a,
                                                                    const __grid_constant__ CUtensorMap tm_w) {
    extern __shared__ unsigned char tt_smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)tt_smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + TT_OFF_BAR);
    uint64_t* empty = full + TT_STAGES;
    uint64_t* tmem_full = empty + TT_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x / p.tiles_per_item, t0 = (blockIdx.x % p.tiles_per_item) * TT_BM, n0 = blockIdx.y * TT_BN;
    // fp16 operands: 64 channels per stage row; a partial last block (Cin = 96) is the tensor map's zero fill
    const int bk = p.f16 ? 2 * TT_BK : TT_BK;
    const int nkb = (p.Cin + bk - 1) / bk, nst = p.taps * nkb;

    if (tid == 0) {
        for (int i = 0; i < TT_STAGES; ++i) { tt_mbar_init(tt_smem_u32(&full[i]), 1); tt_mbar_init(tt_smem_u32(&empty[i]), 1); }
        tt_mbar_init(tt_smem_u32(tmem_full), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tt_smem_u32(tmem_slot)), "n"(TT_BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ===================== producer: A through the 3-D map at (ci, t0 + shift, b), W through the 2-D map ==========
            // K blocks outermost, taps inside
            int s = 0;
            for (int kb = 0; kb < nkb; ++kb) {
                for (int tap = 0; tap < p.taps; ++tap, ++s) {
                    const int ta = t0 + p.shift[tap];
                    const int st = s % TT_STAGES, par = (s / TT_STAGES) & 1;
                    tt_mbar_wait(tt_smem_u32(&empty[st]), par ^ 1);
                    const uint32_t fb = tt_smem_u32(&full[st]);
                    tt_mbar_expect_tx(fb, TT_STAGE_BYTES);
                    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                                 ::"r"(tt_smem_u32(smem + st * TT_STAGE_BYTES)), "l"(reinterpret_cast<uint64_t>(&tm_a)),
                                   "r"(kb * bk), "r"(ta), "r"(b), "r"(fb) : "memory");
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                                 ::"r"(tt_smem_u32(smem + st * TT_STAGE_BYTES + TT_A_BYTES)), "l"(reinterpret_cast<uint64_t>(&tm_w)),
                                   "r"(kb * bk), "r"(tap * p.N + n0), "r"(fb) : "memory");
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===================== MMA issuer: M = 128, N = 128, 32 bytes of K per instruction (8 tf32 / 16 fp16) =========
            const uint32_t fmt = p.f16 ? 0u : 2u;         // operand format field: F16 (kind::f16) / TF32 (kind::tf32)
            const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(TT_BN >> 3) << 17) | ((uint32_t)(TT_BM >> 4) << 24);
            for (int s = 0; s < nst; ++s) {
                const int st = s % TT_STAGES, par = (s / TT_STAGES) & 1;
                tt_mbar_wait(tt_smem_u32(&full[st]), par);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t a_desc = tt_smem_desc(tt_smem_u32(smem + st * TT_STAGE_BYTES));
                const uint64_t b_desc = tt_smem_desc(tt_smem_u32(smem + st * TT_STAGE_BYTES + TT_A_BYTES));
#pragma unroll
                for (int k = 0; k < TT_BK / 8; ++k) {     // 32 bytes = 2 descriptor units along K
                    const uint32_t acc = (s | k) != 0;
                    if (p.f16)
                        asm volatile(
                            "{\n.reg .pred p;\n"
                            "setp.ne.b32 p, %4, 0;\n"
                            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                            ::"r"(tmem_d), "l"(a_desc + 2 * k), "l"(b_desc + 2 * k), "r"(idesc), "r"(acc) : "memory");
                    else
                        asm volatile(
                            "{\n.reg .pred p;\n"
                            "setp.ne.b32 p, %4, 0;\n"
                            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                            ::"r"(tmem_d), "l"(a_desc + 2 * k), "l"(b_desc + 2 * k), "r"(idesc), "r"(acc) : "memory");
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tt_smem_u32(&empty[st])) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tt_smem_u32(tmem_full)) : "memory");
        }
    } else {
        // ===================== epilogue =====================================================================================
        float* stg = reinterpret_cast<float*>(smem);               // the operand ring is free once the accumulator is complete
        const int dt = tid - 64;
        // only the four warps that read TMEM poll the accumulator barrier; the other twelve sleep in the named barrier below (a
        // polling warp re-issues its probe every ~34 cycles: sixteen of them per CTA were a quarter of all issued instructions,
        // ncu source page)
        if (warp < 6) {
            tt_mbar_wait_parked(tt_smem_u32(tmem_full), 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int lg = warp & 3, r_local = lg * 32 + lane;     // TMEM lane = output row (time step) of this tile
            const uint32_t tbase = tmem_d + ((uint32_t)(lg * 32) << 16);
            const int ncols_ld = min(TT_BN, (p.N - n0 + 15) & ~15);      // only the columns this tile owns (N = 96: 96 of 128)
#pragma unroll 1
            for (int c0 = 0; c0 < ncols_ld; c0 += 16) {
                uint32_t v[16];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                               "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                             : "r"(tbase + (uint32_t)c0));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                float4* d = reinterpret_cast<float4*>(stg + r_local * TT_STG_LD + c0);
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    d[c] = make_float4(__uint_as_float(v[4 * c]), __uint_as_float(v[4 * c + 1]), __uint_as_float(v[4 * c + 2]), __uint_as_float(v[4 * c + 3]));
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        }
        asm volatile("bar.sync 1, %0;" ::"n"(TT_EPI_WARPS * 32) : "memory");
        const int ncols = p.N - n0 < TT_BN ? p.N - n0 : TT_BN;
        const int nrows = p.rows - t0 < TT_BM ? p.rows - t0 : TT_BM;
        const long long m_base = (long long)b * p.rows + t0;       // global output row of local row 0
        if (p.out_act && p.act == Q3T_ACT_SWIGLU_PAIR) {
            const int hp = ncols >> 1;
            for (int i = dt; i < nrows * hp; i += TT_EPI_WARPS * 32) {
                const int r = i / hp, j = i - r * hp;
                float v[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int n = n0 + 2 * j + q, c = n % p.Cout;
                    float x = stg[r * TT_STG_LD + 2 * j + q];
                    if (p.bias) x += p.bias[c];
                    if (p.scale) x *= p.scale[c];
                    if (p.resid) x += p.resid[(m_base + r) * p.N + n];
                    if (p.out_raw) p.out_raw[(m_base + r) * p.N + n] = tt_round_tf32(x);
                    v[q] = x;
                }
                p.out_act[(m_base + r) * (p.N / 2) + ((n0 + 2 * j) >> 1)] = tt_round_tf32(silu_f(v[0]) * v[1]);
            }
        } else {
            // Four consecutive channels per thread and item (N, Cout and n0 are multiples of 4), at most 8 items per thread.  All
            // residual loads of a thread are issued BEFORE its first store: the scalar one-element loop this replaces interleaved a
            // dependent DRAM load with two stores per element - 24 serialized round trips per warp, 34 us per tile for the 1x1
            // convolutions of the vocoder (ncu: 2 TB/s of DRAM traffic on a layer that only moves bytes).
            constexpr int ITEMS = 4;      // per batch; (128 x 128 / 4) / 256 threads = 16 items per thread = four batches
            const int nc4 = ncols >> 2, n_items = nrows * nc4;
            const unsigned inv_nc4 = (1u << 20) / (unsigned)nc4 + 1u;       // i / nc4 == (i * inv_nc4) >> 20 for i < 4096, nc4 <= 32
            const bool plain_c = p.N == p.Cout;                           // every layer but the transposed convs: channel = column
            for (int i0 = dt; i0 < n_items; i0 += ITEMS * TT_EPI_WARPS * 32) {
            float4 res[ITEMS];
            if (p.resid) {
#pragma unroll
                for (int q = 0; q < ITEMS; ++q) {
                    const int i = i0 + q * (TT_EPI_WARPS * 32);
                    if (i < n_items) {
                        const int r = (int)(((unsigned)i * inv_nc4) >> 20), j = (i - r * nc4) << 2;
                        res[q] = __ldcs(reinterpret_cast<const float4*>(p.resid + (m_base + r) * p.N + n0 + j));
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < ITEMS; ++q) {
                const int i = i0 + q * (TT_EPI_WARPS * 32);
                if (i >= n_items) break;
                const int r = (int)(((unsigned)i * inv_nc4) >> 20), j = (i - r * nc4) << 2;
                const int n = n0 + j, c = plain_c ? n : n % p.Cout;
                const float4 s4 = *reinterpret_cast<const float4*>(stg + r * TT_STG_LD + j);
                float x[4] = {s4.x, s4.y, s4.z, s4.w};
                if (p.bias) { const float4 t4 = __ldg(reinterpret_cast<const float4*>(p.bias + c)); x[0] += t4.x; x[1] += t4.y; x[2] += t4.z; x[3] += t4.w; }
                if (p.scale) { const float4 t4 = __ldg(reinterpret_cast<const float4*>(p.scale + c)); x[0] *= t4.x; x[1] *= t4.y; x[2] *= t4.z; x[3] *= t4.w; }
                if (p.resid) { x[0] += res[q].x; x[1] += res[q].y; x[2] += res[q].z; x[3] += res[q].w; }
                const long long o = (m_base + r) * p.N + n;
                if (p.out_raw)
                    *reinterpret_cast<float4*>(p.out_raw + o) = make_float4(tt_round_tf32(x[0]), tt_round_tf32(x[1]), tt_round_tf32(x[2]), tt_round_tf32(x[3]));
                if (p.out_act) {
                    if (p.act == Q3T_ACT_SNAKE) {
                        const float4 a4 = __ldg(reinterpret_cast<const float4*>(p.act_a + c)), b4 = __ldg(reinterpret_cast<const float4*>(p.act_b + c));
                        const float aa[4] = {a4.x, a4.y, a4.z, a4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) { const float sn = tt_sin(x[e] * aa[e]); x[e] = x[e] + bb[e] * (sn * sn); }
                    } else if (p.act == Q3T_ACT_GELU) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) x[e] = tt_gelu_erf(x[e]);
                    } else if (p.act == Q3T_ACT_SILU) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) x[e] = silu_f(x[e]);
                    } else if (p.act >= Q3T_ACT_ELU) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) x[e] = act_simple(x[e], p.act);
                    }
                    if (p.act_f16) {
                        // fp16 keeps 11 significant bits (TF32: 10); saturate instead of overflowing to inf
                        uint2 hv;
                        hv.x = tt_pack_h2_sat(x[0], x[1]); hv.y = tt_pack_h2_sat(x[2], x[3]);
                        *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(p.out_act) + o) = hv;
                    } else {
                        *reinterpret_cast<float4*>(p.out_act + o) = make_float4(tt_round_tf32(x[0]), tt_round_tf32(x[1]), tt_round_tf32(x[2]), tt_round_tf32(x[3]));
                    }
                }
            }
            }
        }
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(TT_BN) : "memory");
    }
}

typedef CUresult (*TtEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

bool tapgemm_tc_eligible(const q3t_tapgemm_args* a) {
    const int N = a->up * a->Cout;
    if (a->force_fp32 || a->Cin % TT_BK != 0 || N % 16 != 0 || N < 32 || a->Cout % 4 != 0) return false;
    const long long Mtot = (long long)a->B * a->T_out_rows;
    return Mtot >= 64 && a->T_in >= 1;
}

// returns 0 on success, > 0 on error, -1 when the shape is not eligible (caller falls back to the FP32-pipe kernel)
int launch_tapgemm_tc(const q3t_tapgemm_args* a, cudaStream_t stream) {
    const int N = a->up * a->Cout;
    if (!tapgemm_tc_eligible(a)) return -1;
    if (a->act == Q3T_ACT_SWIGLU_PAIR && (a->a_f16 || a->act_f16)) return -1;
    const CUtensorMapDataType dt = a->a_f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    const cuuint64_t esz = a->a_f16 ? 2 : 4;
    const cuuint32_t bk = a->a_f16 ? 2 * TT_BK : TT_BK;
    static TtEncodeFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) return -1;
        encode = (TtEncodeFn)fn;
    }
    CUtensorMap tm_a, tm_w;
    {
        const cuuint64_t gdim[3] = {(cuuint64_t)a->Cin, (cuuint64_t)a->T_in, (cuuint64_t)a->B};
        const cuuint64_t gstr[2] = {(cuuint64_t)a->Cin * esz, (cuuint64_t)a->T_in * a->Cin * esz};
        const cuuint32_t box[3] = {bk, TT_BM, 1}, es[3] = {1, 1, 1};
        if (encode(&tm_a, dt, 3, (void*)a->A, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return -1;
    }
    {
        const cuuint64_t gdim[2] = {(cuuint64_t)a->Cin, (cuuint64_t)a->taps * N};
        const cuuint64_t gstr[1] = {(cuuint64_t)a->Cin * esz};
        const cuuint32_t box[2] = {bk, TT_BN}, es[2] = {1, 1};
        if (encode(&tm_w, dt, 2, (void*)a->W, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return -1;
    }
    TapTcParams p;
    memset(&p, 0, sizeof(p));
    p.B = a->B; p.T_in = a->T_in; p.Cin = a->Cin; p.taps = a->taps;
    for (int i = 0; i < 8; ++i) p.shift[i] = a->shift[i];
    p.N = N; p.Cout = a->Cout; p.rows = a->T_out_rows; p.tiles_per_item = (a->T_out_rows + TT_BM - 1) / TT_BM;
    p.bias = a->bias; p.scale = a->scale; p.resid = a->resid; p.out_raw = a->out_raw; p.out_act = a->out_act; p.act = a->act;
    p.act_a = a->act_a; p.act_b = a->act_b; p.f16 = a->a_f16; p.act_f16 = a->act_f16;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(tapgemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TT_SMEM_BYTES);
        attr_set = true;
    }
    dim3 grid((unsigned)(a->B * p.tiles_per_item), (unsigned)((N + TT_BN - 1) / TT_BN));
    tapgemm_tc_kernel<<<grid, TT_THREADS, TT_SMEM_BYTES, stream>>>(p, tm_a, tm_w);
    Q3T_CHECK_LAUNCH("tapgemm_tc");
    return 0;
}

}  // namespace q3t
