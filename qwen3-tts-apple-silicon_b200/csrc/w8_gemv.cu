// W8 (affine, group 64) dequant-fused GEMV for batch-1/2 decode on sm_100a.
//
// Replaces mx.quantized_matmul (qmv path) of the reference stack (call sites: SURVEY.md 8a a5/a7).
//
// Design (B200-first, HBM-bound):
//   * weights are pre-tiled on the host in mma-fragment order (weights.py:pack_w8): one warp-wide
//     128-bit load IS the A operand of `mma.sync.m16n8k32.s32.u8.s8` -- coalesced 512 B requests,
//     no shared-memory staging and no shuffle for data that has zero reuse;
//   * the activation vector is converted ONCE per CTA to 32-bit block fixed point and split into four
//     signed base-256 digits that occupy four of the eight MMA columns (two batch rows fit), staged
//     in shared memory in B-fragment order (one LDS.128 per group per lane);
//   * the uint8 x int8 products accumulate exactly in int32 per 64-wide quantisation group; one
//     I2F + FFMA per (row, digit, group) applies the bf16 group scale; the group bias multiplies the
//     per-group sum of x.  y = xscale * sum_g s_g * (sum_d 256^d I_{g,d}) + sum_g b_g * X_g.
//   Per byte of weight this costs ~1/64 of an IMMA + a few scalar ops instead of the 3 ALU ops/byte of
//   a convert-and-FMA GEMV, which at 6.5 TB/s would saturate the FP32 pipe of a B200 SM.
//   * the result is deterministic: no atomics, fixed reduction order.
#include "common.cuh"
#include "../../include/q3tts_b200.h"

namespace q3t {

constexpr int GEMV_THREADS = 512;   // 16 warps x 2 units in flight: measured faster than 8 warps + PDL co-residency
constexpr int GEMV_WARPS = GEMV_THREADS / 32;

struct GemvParams {
    const uint8_t* w;
    int N, K, nrt, nkc;
    const float* lin_bias;
    int prologue;
    const float* x; long long x_stride;
    const float* norm_w; float eps;
    const int* gather_idx; int gather_idx_stride; long long gather_row_stride;
    int act;
    const float* resid; long long resid_stride;
    float* y; long long y_stride;
    int ks;  // warps cooperating on one row tile (split-K inside the CTA)
};

__device__ __forceinline__ void imma_16832(int (&c)[4], const uint32_t a0, const uint32_t a1, const uint32_t a2,
                                            const uint32_t a3, const uint32_t b0, const uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

struct UnitRegs {
    uint4 q[4];      // 2 groups x 2 mma A operands
    uint4 mlo, mhi;  // metadata rows g and g+8: {s0s1, s2s3, b0b1, b2b3}
};

__device__ __forceinline__ void load_unit(UnitRegs& r, const uint8_t* tile, int half, int lane) {
    const uint8_t* qb = tile + half * 2048 + lane * 16;
#pragma unroll
    for (int i = 0; i < 4; ++i) r.q[i] = ldg_stream(qb + i * 512);
    const int g = lane >> 2;
    r.mlo = ldg_stream(tile + 4096 + g * 16);
    r.mhi = ldg_stream(tile + 4096 + (g + 8) * 16);
}

// dynamic shared memory carve-up
struct GemvSmem {
    uint4* xfrag;   // [K/64][32] B-fragment-ordered digits
    float* xsum;    // [2][K/64] per-group sum of x (real units)
    float* xscl;    // [2][K/64] per-group fixed-point scale (block floating point, one exponent per 64 inputs)
    float* red;     // [16 warps][16 rows][2]
    float* scratch; // [32] reductions
};

constexpr int GEMV_MAXV = 4;   // float4 per thread per row: K <= 4*4*512 = 8192

__device__ __forceinline__ void l2_prefetch_bulk(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

template <int M>
__global__ void __launch_bounds__(GEMV_THREADS, 1) w8_gemv_kernel(const GemvParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int K = p.K, ngroups = K >> 6;
    GemvSmem s;
    s.xfrag = reinterpret_cast<uint4*>(smem_raw);
    s.xsum = reinterpret_cast<float*>(smem_raw + (size_t)ngroups * 512);
    s.xscl = s.xsum + 2 * ngroups;
    s.red = s.xscl + 2 * ngroups;
    s.scratch = s.red + GEMV_WARPS * 16 * 2;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K4 = K >> 2;
    const int ks = p.ks, nteams = GEMV_WARPS / ks, team = warp / ks, jw = warp % ks;
    const int rt_begin = (int)(((long long)p.nrt * blockIdx.x) / gridDim.x);
    const int rt_end = (int)(((long long)p.nrt * (blockIdx.x + 1)) / gridDim.x);
    const int U = 2 * p.nkc;

    pdl_launch_dependents();   // the next kernel may start its own weight prefetch now
    // ---- start the HBM stream before anything else: this CTA's row tiles are one contiguous slab; ask the
    // L2 for all of it now so the weight traffic overlaps the activation prologue and the compute below
    if (tid < 32) {
        const uint8_t* slab = p.w + (size_t)rt_begin * p.nkc * Q3T_TILE_BYTES;
        const size_t slab_bytes = (size_t)(rt_end - rt_begin) * p.nkc * Q3T_TILE_BYTES;   // multiple of 16
        constexpr uint32_t PIECE = 16384;
        for (size_t off = (size_t)tid * PIECE; off < slab_bytes; off += 32 * (size_t)PIECE) {
            const size_t n = slab_bytes - off < PIECE ? slab_bytes - off : PIECE;
            l2_prefetch_bulk(slab + off, (uint32_t)n);
        }
    }
    // first unit of this warp's first row tile: loads are independent of x, issue them before the prologue
    UnitRegs cur;
    {
        const int rt = rt_begin + team;
        if (rt < rt_end && jw < U)
            load_unit(cur, p.w + ((size_t)rt * p.nkc + (jw >> 1)) * Q3T_TILE_BYTES, jw & 1, lane);
    }

    // ------------------------------------------------------------------ prologue: x -> digits
    if (M == 1) {  // columns 4..7 (second batch row) must read as zero
        for (int i = tid; i < ngroups * 32; i += GEMV_THREADS)
            if (((i & 31) >> 2) >= 4) s.xfrag[i] = make_uint4(0, 0, 0, 0);
    }
    pdl_wait();   // activations written by the predecessor are visible from here on
#pragma unroll
    for (int m = 0; m < M; ++m) {
        const float* xr = p.x + m * p.x_stride;
        if (p.gather_idx) xr += (long long)p.gather_idx[m * p.gather_idx_stride] * p.gather_row_stride;
        float4 xv[GEMV_MAXV];
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < GEMV_MAXV; ++i) {
            const int k4 = tid + i * GEMV_THREADS;
            if (k4 < K4) {
                float4 v;
                if (p.prologue == Q3T_PRO_SWIGLU) {
                    // gate/up interleaved in blocks of 8: act[8j + r] = silu(x[16j + r]) * x[16j + 8 + r]
                    const int gi = ((k4 >> 1) << 4) + ((k4 & 1) << 2);
                    v = *reinterpret_cast<const float4*>(xr + gi);
                    const float4 u = *reinterpret_cast<const float4*>(xr + gi + 8);
                    v.x = silu_f(v.x) * u.x; v.y = silu_f(v.y) * u.y; v.z = silu_f(v.z) * u.z; v.w = silu_f(v.w) * u.w;
                } else {
                    v = reinterpret_cast<const float4*>(xr)[k4];
                }
                xv[i] = v;
                ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
            }
        }
        if (p.prologue == Q3T_PRO_RMSNORM) {
            const float tot = block_sum(ss, s.scratch);
            const float rstd = rsqrtf(tot / (float)K + p.eps);
#pragma unroll
            for (int i = 0; i < GEMV_MAXV; ++i) {
                const int k4 = tid + i * GEMV_THREADS;
                if (k4 < K4) {
                    const float4 nw = reinterpret_cast<const float4*>(p.norm_w)[k4];
                    float4& v = xv[i];
                    v.x = nw.x * (v.x * rstd); v.y = nw.y * (v.y * rstd); v.z = nw.z * (v.z * rstd); v.w = nw.w * (v.w * rstd);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < GEMV_MAXV; ++i) {
            const int k4 = tid + i * GEMV_THREADS;   // K % 256 == 0 -> whole warps stay converged
            if (k4 < K4) {
                const float4 v = xv[i];
                // block floating point: one scale per 64-wide quantisation group (16 consecutive lanes)
                float amax = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
                const float inv = amax > 0.f ? 1073741824.f / amax : 0.f;
                const float xscale = amax * (1.f / 1073741824.f);
                int e[4] = {__float2int_rn(v.x * inv), __float2int_rn(v.y * inv), __float2int_rn(v.z * inv),
                            __float2int_rn(v.w * inv)};
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
                for (int d = 0; d < 4; ++d) base[((4 * m + d) * 4 + t) * 4 + r] = wd[d];
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) gs += __shfl_xor_sync(0xffffffffu, gs, o);
                if ((lane & 15) == 0) {
                    s.xsum[m * ngroups + G] = (float)gs * xscale;
                    s.xscl[m * ngroups + G] = xscale;
                }
            }
        }
    }
    __syncthreads();

    // ------------------------------------------------------------------ main loop
    const int g = lane >> 2, t = lane & 3, mt = (M == 2) ? (t >> 1) : 0;
    const float pw_lo = (t & 1) ? 65536.f : 1.f, pw_hi = pw_lo * 256.f;
    const float* xscl = s.xscl + mt * ngroups;
    const float* xsum = s.xsum + mt * ngroups;

    for (int rt0 = rt_begin; rt0 < rt_end; rt0 += nteams) {
        const int rt = rt0 + team;
        float res_lo = 0.f, res_hi = 0.f;
        if (rt < rt_end) {
            const uint8_t* row_base = p.w + (size_t)rt * p.nkc * Q3T_TILE_BYTES;
            float f[4] = {0.f, 0.f, 0.f, 0.f};
            float bacc_lo = 0.f, bacc_hi = 0.f;
            UnitRegs nxt;
            int u = jw;
            if (rt0 != rt_begin && u < U) load_unit(cur, row_base + (size_t)(u >> 1) * Q3T_TILE_BYTES, u & 1, lane);
            for (; u < U; u += ks) {
                const int un = u + ks;
                if (un < U) load_unit(nxt, row_base + (size_t)(un >> 1) * Q3T_TILE_BYTES, un & 1, lane);
                const int half = u & 1;
                const uint32_t sw_lo = half ? cur.mlo.y : cur.mlo.x, sw_hi = half ? cur.mhi.y : cur.mhi.x;
                const uint32_t bw_lo = half ? cur.mlo.w : cur.mlo.z, bw_hi = half ? cur.mhi.w : cur.mhi.z;
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                    const int G = (u >> 1) * 4 + half * 2 + jj;
                    const uint4 b = s.xfrag[G * 32 + lane];
                    int acc[4] = {0, 0, 0, 0};
                    imma_16832(acc, cur.q[2 * jj].x, cur.q[2 * jj].y, cur.q[2 * jj].z, cur.q[2 * jj].w, b.x, b.y);
                    imma_16832(acc, cur.q[2 * jj + 1].x, cur.q[2 * jj + 1].y, cur.q[2 * jj + 1].z,
                               cur.q[2 * jj + 1].w, b.z, b.w);
                    const float xg = xscl[G];
                    const float slo = (jj ? bf16hi(sw_lo) : bf16lo(sw_lo)) * xg;
                    const float shi = (jj ? bf16hi(sw_hi) : bf16lo(sw_hi)) * xg;
                    f[0] = fmaf(slo, (float)acc[0], f[0]);
                    f[1] = fmaf(slo, (float)acc[1], f[1]);
                    f[2] = fmaf(shi, (float)acc[2], f[2]);
                    f[3] = fmaf(shi, (float)acc[3], f[3]);
                    const float xs = xsum[G];
                    bacc_lo = fmaf(jj ? bf16hi(bw_lo) : bf16lo(bw_lo), xs, bacc_lo);
                    bacc_hi = fmaf(jj ? bf16hi(bw_hi) : bf16lo(bw_hi), xs, bacc_hi);
                }
                if (un < U) cur = nxt;
            }
            float v_lo = f[0] * pw_lo + f[1] * pw_hi;
            float v_hi = f[2] * pw_lo + f[3] * pw_hi;
            v_lo += __shfl_xor_sync(0xffffffffu, v_lo, 1);
            v_hi += __shfl_xor_sync(0xffffffffu, v_hi, 1);
            res_lo = v_lo + bacc_lo;
            res_hi = v_hi + bacc_hi;
        }
        if ((t & 1) == 0 && (M == 2 || t == 0)) {
            const int mo = t >> 1;
            s.red[(warp * 16 + g) * 2 + mo] = res_lo;
            s.red[(warp * 16 + g + 8) * 2 + mo] = res_hi;
        }
        __syncthreads();
        if (jw == 0 && rt < rt_end) {
            const int row = lane & 15, m = lane >> 4;
            if (m < M) {
                float v = 0.f;
                for (int j = 0; j < ks; ++j) v += s.red[((warp + j) * 16 + row) * 2 + m];
                const int n = rt * 16 + row;
                if (p.lin_bias) v += p.lin_bias[n];
                if (p.act == Q3T_ACT_SILU) v = silu_f(v);
                if (p.resid) v += p.resid[m * p.resid_stride + n];
                p.y[m * p.y_stride + n] = v;
            }
        }
        __syncthreads();
    }
}

static size_t gemv_smem_bytes(int M, int K) {
    const int ng = K / 64;
    (void)M;
    return (size_t)ng * 512 + sizeof(float) * (4 * ng + GEMV_WARPS * 16 * 2 + 40);
}

static int g_num_sms = 0;
int num_sms() {
    if (!g_num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

int launch_w8_gemv(const q3t_gemv_args* a, cudaStream_t stream) {
    Q3T_REQUIRE(a->M == 1 || a->M == 2, "w8_gemv: M must be 1 or 2");
    Q3T_REQUIRE(a->w.N % 16 == 0 && a->w.K % 256 == 0, "w8_gemv: N%16 / K%256");
    Q3T_REQUIRE(a->w.K <= 8192, "w8_gemv: K too large for the shared-memory staging");
    GemvParams p;
    p.w = (const uint8_t*)a->w.w; p.N = a->w.N; p.K = a->w.K; p.nrt = p.N / 16; p.nkc = p.K / 256;
    p.lin_bias = a->w.lin_bias; p.prologue = a->prologue; p.x = a->x; p.x_stride = a->x_stride;
    p.norm_w = a->norm_w; p.eps = a->eps; p.gather_idx = a->gather_idx; p.gather_idx_stride = a->gather_idx_stride;
    p.gather_row_stride = a->gather_row_stride; p.act = a->act; p.resid = a->resid; p.resid_stride = a->resid_stride;
    p.y = a->y; p.y_stride = a->y_stride;
    const int grid = p.nrt < num_sms() ? p.nrt : num_sms();
    const int per_cta = (p.nrt + grid - 1) / grid;
    int ks = GEMV_WARPS;
    while (ks > 1 && (GEMV_WARPS / ks) < per_cta) ks >>= 1;
    while (ks > 2 * p.nkc) ks >>= 1;
    p.ks = ks;
    const size_t smem = gemv_smem_bytes(a->M, p.K);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(w8_gemv_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(w8_gemv_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_set = true;
    }
    if (a->M == 1) launch_pdl(w8_gemv_kernel<1>, dim3(grid), dim3(GEMV_THREADS), smem, stream, p);
    else launch_pdl(w8_gemv_kernel<2>, dim3(grid), dim3(GEMV_THREADS), smem, stream, p);
    Q3T_CHECK_LAUNCH("w8_gemv");
    return 0;
}

}  // namespace q3t

extern "C" int q3t_w8_gemv(const q3t_gemv_args* a, void* stream) {
    return q3t::launch_w8_gemv(a, (cudaStream_t)stream);
}

extern "C" int q3t_w8_gemv_rows(const q3t_gemv_args* a0, int n_rows, void* stream) {
    for (int r = 0; r < n_rows; r += 2) {
        q3t_gemv_args a = *a0;
        a.M = n_rows - r >= 2 ? 2 : 1;
        if (a0->gather_idx) a.gather_idx = a0->gather_idx + (long long)r * a0->gather_idx_stride;
        else a.x = a0->x + (long long)r * a0->x_stride;
        a.y = a0->y + (long long)r * a0->y_stride;
        if (a0->resid) a.resid = a0->resid + (long long)r * a0->resid_stride;
        if (const int rc = q3t_w8_gemv(&a, stream)) return rc;
    }
    return 0;
}
