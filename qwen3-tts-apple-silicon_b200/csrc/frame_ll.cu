// Persistent data-flow kernel for batch-1 decode: ONE cooperative launch runs either
//   * a token through a whole dense Qwen3 stack (talker prefill token / talker decode step), or
//   * a whole 80 ms FRAME of the generation hot loop: sample code 0 -> 16 code-predictor passes (15 sampled codes) ->
//     next-input embedding sum -> talker decode step -> logits of the next frame                         (K7)
// Replaces the per-frame body of mlx_audio's Model.generate loop (reference call site sessions/custom.py:163-170;
// structural cousin transformers qwen3_omni_moe/modeling_qwen3_omni_moe.py:3243-3279).
//
// Why it looks like this (measurements: tools/ubench.cu -> profiles/r01_ubench.txt):
//   * a grid-wide barrier on 148 CTAs costs 1.8 us (atomic + fence + poll, cooperative-groups grid.sync() included) and
//     the data still has to be read afterwards; a layer has five dependent contractions, so barriers alone would cost
//     more than the 8.2 us a layer's 53.7 MB of W8 weights need at the HBM roofline.  Here there are NO barriers and NO
//     fences between phases: every activation crosses CTAs as a 64-bit {fp32 value, 32-bit phase tag} word written with
//     one st.relaxed.gpu and polled with ld.relaxed.gpu until the tag matches (the "LL" protocol of collective
//     libraries, applied inside one GPU).  One store->poll hop is ~0.3 us, a full all-to-all exchange ~1.0 us.
//   * weights never wait for activations: a producer warp per CTA streams this CTA's tiles of EVERY matrix of the
//     launch, in program order, through a 32-slot shared-memory ring with cp.async.bulk (TMA, `UBLKCP`) + mbarrier
//     full/empty pairs (four issuing lanes: one lane tops out at ~186 cycles per 4352-byte tile);
//   * 16 consumer warps run the IMMA dequant-dot (uint8 codes x signed base-256 digits of the block-fixed-point
//     activation, exact int32 sums per quantisation group - see w8_gemv.cu) out of the ring;
//   * per-WARP data flow inside a CTA: a warp polls only the 256 inputs of its own tile's k-chunk, converts them into a
//     warp-private digit area and starts its tile - no block-wide prologue, no barrier in front of the tiles; 1/rms of the
//     RMSNorm is applied in the epilogue, the residual stream is a per-CTA double buffer (see gemv_phase);
//   * row tiles (16 output rows x all of K) are dealt to CTAs whole, so every output element has exactly one producer and
//     epilogues (bias, SwiGLU) run before the value is published; the residual stream is replicated per CTA in shared
//     memory, so residual adds never touch global memory;
//   * attention: (kv head, 64-token split) units on the first n_kv*nsplit CTAs.  The chunk's K/V pages are whole 4 KB
//     blocks of the paged cache: eight cp.async.bulk copies (one warp instruction) stage them a whole layer ahead;
//     scores by 8 threads per token, then per warp (head, dim half, token group) softmax + P.V out of shared memory.
//     Split 0 of every kv head merges the other splits' (m, l, acc) records (one poll round) and publishes normalised
//     head outputs, so every consumer of the O projection polls ONE float4 instead of nsplit records of hot lines.
//     Contexts of at most 32 tokens (every code-predictor pass) take a barrier-free path, one warp per q head (attn_tiny);
//   * the input of a code-predictor pass and its first layer's q|k|v are functions of ONE sampled code: table rows computed
//     at load (q3t_frame_args.cp_proj_rows_dev / cp_qkv0_rows_dev), two contraction phases less in 15 of the 16 passes.
// Re-use of an exchange buffer is safe without extra synchronisation because every phase is an all-to-all dependency:
// a CTA can only be one phase ahead of the slowest CTA, and each buffer is rewritten five or more phases later.
#include <stdlib.h>
#include "sampler.cuh"

namespace q3t {

typedef unsigned long long u64;

constexpr int LL_CWARPS = 16;
constexpr int LL_CTHREADS = LL_CWARPS * 32;      // 512 consumers
constexpr int LL_THREADS = LL_CTHREADS + 64;     // + producer warp + L2 prefetch warp (18 warps cost no registers: the cap
                                                 // of 96 comes from five warps on one scheduler partition either way)
#ifndef LL_NSLOT_N
#define LL_NSLOT_N 32
#endif
constexpr int LL_NSLOT = LL_NSLOT_N;             // ring slots of one 4352-byte tile
#ifndef LL_PLANES_N
#define LL_PLANES_N 4
#endif
constexpr int LL_PLANES = LL_PLANES_N;           // producer lanes issuing TMA copies (one lane needs ~95 ns per tile)
constexpr int LL_MAXK = 6144;                    // largest contraction length (talker intermediate size)
constexpr int LL_MAXH = 2048;                    // largest hidden size (one float4 per consumer thread)
constexpr int LL_MAXT = 48;                      // max tiles of one matrix per CTA
constexpr int LL_MAXSPLIT = 16;                  // attention splits per kv head
constexpr int LL_REC = 130;                      // attention record: 128 acc + m + l
constexpr int LL_SPIN_LIMIT = 1 << 22;           // watchdog: a poll that spins this long (~1 s) aborts the launch
constexpr int LL_NSTAMP = 2048;

enum { LL_MODE_STACK = 0, LL_MODE_FRAME = 1 };
enum { EPI_RAW = 0, EPI_SWIGLU = 1 };

struct LLStack {
    const q3t_layer* layers;                     // DEVICE array
    int n_layers, hidden, n_heads, n_kv, head_dim, inter;
    float eps;
    const float* final_norm; const float* inv_freq;
    __nv_bfloat16* kv_pool; long long kv_layer_stride;   // elements
    const int* block_tbl;
};

struct LLParams {
    int mode;
    LLStack talker, cp;
    // exchange buffers (64-bit {value, tag} words)
    u64 *x_qkv, *x_attn, *x_attnf, *x_o, *x_act, *x_down, *x_head, *x_proj;   // x_attnf: merged, normalised head outputs
    unsigned int* state;                         // [0] tag base (persists across launches), [1] error code
    unsigned long long* timing;                  // optional [grid][LL_NSTAMP]
    int pf_dist;                                 // L2 prefetch distance in tiles ahead of the TMA cursor (0 = off)
    int fine;                                    // sub-phase profiling stamps (Q3T_LL_FINE=1)
    int att_chunk, att_maxsplit;                 // attention geometry: smallest chunk (tokens) and most splits per kv head
    // ---- stack mode
    int which;                                   // 0 = talker stack, 1 = code-predictor stack
    const int* pos; const float* x_in; float* hidden_out; float* logits_out; q3t_w8 head;
    // ---- frame mode
    q3t_w8 codec_head, cp_proj; const q3t_w8* cp_heads;          // DEVICE array [G-1]
    const float* codec_embedding; const float* const* cp_embeddings;   // DEVICE array [G-1]
    const float* const* cp_proj_rows;            // optional DEVICE array [G-1]: projected embedding tables (see q3t_frame_args)
    const float* const* cp_qkv0_rows;            // optional: first-layer q|k|v tables of the same rows
    int emb_dim, talker_vocab, cp_vocab, n_groups;
    q3t_sampling talker_sp, cp_sp;
    float* x; float* hidden; float* logits; float* cp_logits; int keep_cp_logits;
    int* pos_talker; int* step; int* cur_codes; int* codes; int* own_codes; int max_frames;
    unsigned int* seen; int* done;
    const float* trailing; int n_trailing; const int* forced;
};

// ---- PTX helpers ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t));   // per-SM cycle counter: exact for durations inside one CTA
    return t;
}
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
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cbar() { asm volatile("bar.sync 1, %0;" ::"n"(LL_CTHREADS) : "memory"); }

__device__ __forceinline__ void ll_fail(unsigned int* state, unsigned int code) {
    atomicExch(state + 1, code);
    __threadfence_system();
    __trap();
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, unsigned int* state, unsigned int code) {
    int spins = 0;
    while (!mbar_try(bar, parity))
        if (++spins > LL_SPIN_LIMIT) ll_fail(state, code);
}

// ---- LL words -----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ll_st(u64* p, float v, uint32_t tag) {
    const u64 w = ((u64)tag << 32) | (u64)__float_as_uint(v);
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ void ll_ld2(const u64* p, u64& a, u64& b) {
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ float ll_val(u64 w) { return __uint_as_float((uint32_t)w); }
__device__ __forceinline__ bool ll_ok(u64 w, uint32_t tag) { return (uint32_t)(w >> 32) == tag; }

// NV float4s (4 consecutive words each) at p[i]; all loads are issued before the first tag is looked at
template <int NV>
__device__ __forceinline__ void ll_ld4n(const u64* const (&p)[NV], const bool (&on)[NV], uint32_t tag, float4 (&out)[NV],
                                        unsigned int* state) {
    u64 w[NV][4];
    bool ok[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        ok[i] = !on[i];
        if (on[i]) { ll_ld2(p[i], w[i][0], w[i][1]); ll_ld2(p[i] + 2, w[i][2], w[i][3]); }
    }
    int spins = 0;
    for (;;) {
        bool all = true;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            if (!ok[i]) ok[i] = ll_ok(w[i][0], tag) && ll_ok(w[i][1], tag) && ll_ok(w[i][2], tag) && ll_ok(w[i][3], tag);
            all = all && ok[i];
        }
        if (all) break;
        if (++spins > LL_SPIN_LIMIT) ll_fail(state, 0x100u | (tag & 0xffu));
#if LL_POLL_SLEEP > 0
        __nanosleep(LL_POLL_SLEEP);
#endif
#pragma unroll
        for (int i = 0; i < NV; ++i)
            if (!ok[i]) { ll_ld2(p[i], w[i][0], w[i][1]); ll_ld2(p[i] + 2, w[i][2], w[i][3]); }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i)
        out[i] = on[i] ? make_float4(ll_val(w[i][0]), ll_val(w[i][1]), ll_val(w[i][2]), ll_val(w[i][3])) : make_float4(0.f, 0.f, 0.f, 0.f);
}
__device__ __forceinline__ float4 ll_ld4(const u64* p, uint32_t tag, unsigned int* state) {
    const u64* const pp[1] = {p};
    const bool on[1] = {true};
    float4 o[1];
    ll_ld4n<1>(pp, on, tag, o, state);
    return o[0];
}

__device__ __forceinline__ void imma_16832_ll(int (&c)[4], const uint4 a, const uint32_t b0, const uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
        : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}

// ---- shared memory --------------------------------------------------------------------------------------------------------
// Phases are device functions that reach the launch parameters and the carve-up through the dynamic shared-memory base.
// LL_FN selects inlining: fully inlined is ~0.35 us per phase faster than calls (measured), at 600 KB of SASS.
extern __shared__ __align__(128) unsigned char ll_smem_raw[];

#ifndef LL_FN
#define LL_FN __forceinline__
#endif
#ifndef LL_POLL_SLEEP
#define LL_POLL_SLEEP 0                          // ns between poll rounds of the LL words (experiments: 0 is fastest)
#endif

// per-CTA view of one W8 matrix / one layer, built once per launch in shared memory: no phase starts with a dependent
// global-memory load of a descriptor, a norm-weight pointer or an integer division for its row-tile range
struct MatD {
    const uint8_t* w; const float* bias;
    int nkc, rb, re, N;
};
struct LayerD {
    MatD qkv, o, gu, down;
    const float *input_norm, *post_norm, *q_norm, *k_norm;
};
constexpr int LL_MAXLAYERS = 40;                 // talker + code-predictor layers
constexpr int LL_MAXHEADS = 24;                  // cp_proj, codec head / stack head, code-predictor heads

struct LLSmem {
    LayerD* lay;            // [LL_MAXLAYERS]  talker layers first, then code-predictor layers
    MatD* hd;               // [LL_MAXHEADS]   0 = cp_proj, 1 = codec head (or the stack-mode head), 2.. = cp heads
    uint8_t* ring;          // [LL_NSLOT][4352]
    uint4* dig;             // [LL_CWARPS][4 groups][16] warp-private digit planes of one 256-wide k-chunk (mma B-fragment order)
    float* xsum;            // [LL_CWARPS][4]
    float* xscl;            // [LL_CWARPS][4]
    unsigned char* stage;   // [32 KB] attention K/V staging (cp.async.bulk target); parked split records; sampler scratch
    float* resid;           // [2][LL_MAXH] residual stream, double buffered (see gemv_phase)
    float* tile_out;        // [2][LL_MAXT][16], double buffered by phase parity
    float* ssq;             // [2][32] per-k-chunk sums of squares of the phase input (RMSNorm), by phase parity
    float* att;             // attention scratch: q [2][128], scores [2][64], new k/v, partials
    float* red;             // [2][32] block reductions, double buffered
    float* cs;              // [64] cos(pos * inv_freq)
    float* sn;              // [64]
    int* ibuf;              // [64] sampler scratch ints
    int* pages;             // [64] KV page ids of this CTA's attention chunk (constant during a pass)
    uint64_t* full;         // [LL_NSLOT]
    uint64_t* empty;        // [LL_NSLOT]
    uint64_t* kvbar;        // attention K/V staging (cp.async.bulk completion)
    const float** prows;    // [LL_MAXHEADS] projected embedding tables of the code-predictor passes (optional)
    const float** qrows;    // [LL_MAXHEADS] first-layer q|k|v tables (optional)
};
constexpr size_t LL_STAGE_BYTES = 32768;
constexpr size_t LL_ATT_FLOATS = 4 * 128 + 8 * 128 + 32;
constexpr size_t LL_OFF_LAY = 1024;                                      // the first KB holds a copy of LLParams
constexpr size_t LL_OFF_HD = LL_OFF_LAY + LL_MAXLAYERS * sizeof(LayerD);
constexpr size_t LL_OFF_RING = (LL_OFF_HD + LL_MAXHEADS * sizeof(MatD) + 127) / 128 * 128;
constexpr size_t LL_OFF_DIG = LL_OFF_RING + (size_t)LL_NSLOT * Q3T_TILE_BYTES;
constexpr size_t LL_OFF_XSUM = LL_OFF_DIG + (size_t)LL_CWARPS * 4 * 16 * 16;
constexpr size_t LL_OFF_XSCL = LL_OFF_XSUM + LL_CWARPS * 4 * 4;
constexpr size_t LL_OFF_STAGE = LL_OFF_XSCL + LL_CWARPS * 4 * 4;
constexpr size_t LL_OFF_RESID = LL_OFF_STAGE + LL_STAGE_BYTES;
constexpr size_t LL_OFF_TILE = LL_OFF_RESID + 2 * LL_MAXH * 4;
constexpr size_t LL_OFF_SSQ = LL_OFF_TILE + 2 * LL_MAXT * 16 * 4;
constexpr size_t LL_OFF_ATT = LL_OFF_SSQ + 2 * 32 * 4;
constexpr size_t LL_OFF_RED = LL_OFF_ATT + LL_ATT_FLOATS * 4;
constexpr size_t LL_OFF_CS = LL_OFF_RED + 64 * 4;
constexpr size_t LL_OFF_SN = LL_OFF_CS + 64 * 4;
constexpr size_t LL_OFF_IBUF = LL_OFF_SN + 64 * 4;
constexpr size_t LL_OFF_PAGES = LL_OFF_IBUF + 64 * 4;
constexpr size_t LL_OFF_FULL = LL_OFF_PAGES + 64 * 4;
constexpr size_t LL_OFF_EMPTY = LL_OFF_FULL + LL_NSLOT * 8;
constexpr size_t LL_OFF_KVBAR = LL_OFF_EMPTY + LL_NSLOT * 8;
constexpr size_t LL_OFF_PROWS = LL_OFF_KVBAR + 8;
constexpr size_t LL_OFF_QROWS = LL_OFF_PROWS + LL_MAXHEADS * 8;
constexpr size_t LL_SMEM_BYTES = LL_OFF_QROWS + LL_MAXHEADS * 8;
static_assert(LL_SMEM_BYTES <= 227 * 1024, "frame_ll: shared memory budget exceeded");
static_assert(LL_OFF_STAGE % 128 == 0 && LL_OFF_DIG % 16 == 0, "frame_ll: staging / digit alignment");

__device__ __forceinline__ const LLParams& ll_params() { return *reinterpret_cast<const LLParams*>(ll_smem_raw); }
__device__ __forceinline__ LLSmem ll_smem() {
    LLSmem s;
    unsigned char* b = ll_smem_raw;
    s.lay = reinterpret_cast<LayerD*>(b + LL_OFF_LAY);
    s.hd = reinterpret_cast<MatD*>(b + LL_OFF_HD);
    s.ring = b + LL_OFF_RING;
    s.dig = reinterpret_cast<uint4*>(b + LL_OFF_DIG);
    s.xsum = reinterpret_cast<float*>(b + LL_OFF_XSUM);
    s.xscl = reinterpret_cast<float*>(b + LL_OFF_XSCL);
    s.stage = b + LL_OFF_STAGE;
    s.resid = reinterpret_cast<float*>(b + LL_OFF_RESID);
    s.tile_out = reinterpret_cast<float*>(b + LL_OFF_TILE);
    s.ssq = reinterpret_cast<float*>(b + LL_OFF_SSQ);
    s.att = reinterpret_cast<float*>(b + LL_OFF_ATT);
    s.red = reinterpret_cast<float*>(b + LL_OFF_RED);
    s.cs = reinterpret_cast<float*>(b + LL_OFF_CS);
    s.sn = reinterpret_cast<float*>(b + LL_OFF_SN);
    s.ibuf = reinterpret_cast<int*>(b + LL_OFF_IBUF);
    s.pages = reinterpret_cast<int*>(b + LL_OFF_PAGES);
    s.full = reinterpret_cast<uint64_t*>(b + LL_OFF_FULL);
    s.empty = reinterpret_cast<uint64_t*>(b + LL_OFF_EMPTY);
    s.kvbar = reinterpret_cast<uint64_t*>(b + LL_OFF_KVBAR);
    s.prows = reinterpret_cast<const float**>(b + LL_OFF_PROWS);
    s.qrows = reinterpret_cast<const float**>(b + LL_OFF_QROWS);
    return s;
}

struct CState {
    uint32_t seq;           // tiles consumed so far (ring sequence number)
    uint32_t gen;           // phase tag counter
    int nstamp;
    int nsplit;             // attention geometry of the current pass
    int chunk;
    int red_par;            // parity of the double-buffered block-reduction scratch
    uint32_t kv_par;        // phase parity of the K/V staging barrier
    int res_par;            // which half of the residual double buffer is current
    int pp;                 // phase parity of tile_out / ssq
};

// Profiling stamps (thread 0 of every CTA): {id : 20 bits | clock64 cycles of this SM : 44 bits}.  ids < 32 mark phase ends
// and are always written when a timing buffer is given; ids >= 32 are sub-phase marks, written when Q3T_LL_FINE=1.
// The production library is built WITHOUT the stamps (even disabled at run time they cost ~5 % of a talker step in
// registers and branches); csrc/build.sh also builds libq3tts_b200_prof.so with -DLL_STAMPS for tools/ll_timing.py.
#ifndef LL_STAMPS
#define LL_STAMP(id) do { } while (0)
#else
#define LL_STAMP(id) do { if (p.timing && threadIdx.x == ((p.fine >> 8) << 5) && st.nstamp < LL_NSTAMP && ((id) < 32 || (p.fine & 255))) \
    p.timing[(size_t)blockIdx.x * LL_NSTAMP + st.nstamp++] = ((unsigned long long)(id) << 44) | (gtimer() & ((1ull << 44) - 1)); } while (0)
#endif
enum { ST_START = 0, ST_QKV_PRO = 1, ST_QKV = 2, ST_ATTN = 3, ST_O_PRO = 4, ST_O = 5, ST_GU_PRO = 6, ST_GU = 7, ST_DOWN_PRO = 8,
       ST_DOWN = 9, ST_END = 10, ST_SAMPLE = 11, ST_CP_PASS = 12,
       ST_F_POLL = 32, ST_F_TILES = 34, ST_F_GBAR = 35, ST_F_ATT_A = 36, ST_F_ATT_B = 37, ST_F_ATT_C = 38, ST_F_ATT_D = 39,
       ST_F_ATT_E = 40, ST_F_ATT_F = 41, ST_F_ATT_G = 42, ST_F_MERGE = 43, ST_F_DIG = 44 };

// block sum over the 512 consumer threads; `red` is double buffered by `parity`, so one barrier per call is enough
__device__ __forceinline__ float cblock_sum(float v, float* red, int parity) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    float* r = red + parity * 32;
    if (lane == 0) r[wid] = v;
    cbar();
    return warp_sum((lane < LL_CWARPS) ? r[lane] : 0.f);
}

// v = 4 consecutive inputs starting at k = 4*k4 -> signed base-256 digit planes + per-group sum/scale.
// Whole warps call this together (16 lanes share a 64-wide quantisation group).
// k4 = index of the float4 inside the warp's 256-wide k-chunk (0..63); dig / xsum / xscl are the warp's private areas.
__device__ __forceinline__ void emit_digits(uint4* dig, float* xsum, float* xscl, float4 v, int k4, int lane) {
    float amax = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    const float inv = amax > 0.f ? 1073741824.f / amax : 0.f;
    const float xscale = amax * (1.f / 1073741824.f);
    const int e0 = __float2int_rn(v.x * inv), e1 = __float2int_rn(v.y * inv), e2 = __float2int_rn(v.z * inv), e3 = __float2int_rn(v.w * inv);
    // e = sum_i d_i 256^i with d_i in [-128,127]:  bytes of (e + 0x80808080) are d_i + 128; xor 0x80 gives the s8 encoding
    const uint32_t u0 = ((uint32_t)e0 + 0x80808080u) ^ 0x80808080u, u1 = ((uint32_t)e1 + 0x80808080u) ^ 0x80808080u;
    const uint32_t u2 = ((uint32_t)e2 + 0x80808080u) ^ 0x80808080u, u3 = ((uint32_t)e3 + 0x80808080u) ^ 0x80808080u;
    const uint32_t t0 = __byte_perm(u0, u1, 0x5140), t1 = __byte_perm(u2, u3, 0x5140);
    const uint32_t t2 = __byte_perm(u0, u1, 0x7362), t3 = __byte_perm(u2, u3, 0x7362);
    uint32_t wd[4];
    wd[0] = __byte_perm(t0, t1, 0x5410); wd[1] = __byte_perm(t0, t1, 0x7632);
    wd[2] = __byte_perm(t2, t3, 0x5410); wd[3] = __byte_perm(t2, t3, 0x7632);
    const int k = k4 << 2, G = k >> 6, kk = k & 63;
    const int r = ((kk >> 5) << 1) | ((kk >> 4) & 1), t = (kk >> 2) & 3;
    uint32_t* base = reinterpret_cast<uint32_t*>(dig + G * 16);
#pragma unroll
    for (int d = 0; d < 4; ++d) base[(d * 4 + t) * 4 + r] = wd[d];
    float gs = ((float)e0 + (float)e1) + ((float)e2 + (float)e3);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) gs += __shfl_xor_sync(0xffffffffu, gs, o);
    if ((lane & 15) == 0) { xsum[G] = gs * xscale; xscl[G] = xscale; }
}

// One 4352-byte tile out of the ring against the digit planes: 16 output rows (partial over this tile's 256 inputs).
// Lanes 16..31 carry the four unused columns of the 8-wide B operand: zeros from registers.
__device__ __forceinline__ void tile_dot(const uint4* dig, const float* xsum, const float* xscl, const uint8_t* tile, int lane, float& out_lo, float& out_hi) {
    const int g = lane >> 2, t = lane & 3;
    const uint4 mlo = *reinterpret_cast<const uint4*>(tile + 4096 + g * 16);
    const uint4 mhi = *reinterpret_cast<const uint4*>(tile + 4096 + (g + 8) * 16);
    const uint32_t slo_w[2] = {mlo.x, mlo.y}, shi_w[2] = {mhi.x, mhi.y}, blo_w[2] = {mlo.z, mlo.w}, bhi_w[2] = {mhi.z, mhi.w};
    float f[4] = {0.f, 0.f, 0.f, 0.f}, bacc_lo = 0.f, bacc_hi = 0.f;
#pragma unroll
    for (int j4 = 0; j4 < 4; ++j4) {
        const int G = j4;
        const uint4 a0 = *reinterpret_cast<const uint4*>(tile + (j4 * 2 + 0) * 512 + lane * 16);
        const uint4 a1 = *reinterpret_cast<const uint4*>(tile + (j4 * 2 + 1) * 512 + lane * 16);
        const uint4 b = lane < 16 ? dig[G * 16 + lane] : make_uint4(0u, 0u, 0u, 0u);
        int acc[4] = {0, 0, 0, 0};
        imma_16832_ll(acc, a0, b.x, b.y);
        imma_16832_ll(acc, a1, b.z, b.w);
        const float xg = xscl[G], xs = xsum[G];
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
    out_lo = v_lo + bacc_lo;   // valid in lanes with t == 0 (columns 0..3 = the four digits of the one batch row)
    out_hi = v_hi + bacc_hi;
}

// ---- GEMV phase (consumers): per-WARP data flow ---------------------------------------------------------------------------
// A warp needs only the 256 inputs of its own k-chunk: it polls those words (two float4 per lane), applies the input
// transform, writes the digit planes of that chunk into its PRIVATE 1 KB area and starts its tile - no block-wide prologue,
// no barrier before the tiles, and a chunk whose producers are late delays one warp instead of the CTA.  RMSNorm: the
// digits are taken from norm_w * x (block fixed point is scale-invariant per group) and 1/rms multiplies the row sums in
// the epilogue; the per-chunk sums of squares meet in shared memory at the one barrier a phase keeps (tiles -> row sums).
// Residual stream: x = resid + delta is written by the warp of tile j < nkc into the OTHER half of a double buffer (several
// warps may convert the same chunk - 16 warps, 4 or 8 chunks - and all of them read the old half).  tile_out / ssq are
// double buffered by phase parity: a fast warp may be in the tiles of phase p+1 while a slow one is in the epilogue of p.
enum { IN_NORM = 0, IN_LL = 1, IN_PLAIN = 2 };
struct PhaseIn {
    int kind;
    const u64* words; uint32_t tag;      // IN_NORM: residual delta (or nullptr); IN_LL: the input vector
    const float* norm_w; float inv_h, eps;   // IN_NORM
    const float* plain;                  // IN_PLAIN: fp32 vector in global memory (table row / written by an earlier launch)
    float* hidden_out;                   // IN_NORM: CTA 0 stores the normalised input here (final norm), or nullptr
};
__device__ __forceinline__ PhaseIn in_norm(const u64* add, uint32_t tag, const float* w, int H, float eps, float* hidden_out = nullptr) {
    PhaseIn in; in.kind = IN_NORM; in.words = add; in.tag = tag; in.norm_w = w; in.inv_h = 1.f / (float)H; in.eps = eps;
    in.plain = nullptr; in.hidden_out = hidden_out; return in;
}
__device__ __forceinline__ PhaseIn in_ll(const u64* words, uint32_t tag) {
    PhaseIn in; in.kind = IN_LL; in.words = words; in.tag = tag; in.norm_w = nullptr; in.inv_h = 0.f; in.eps = 0.f;
    in.plain = nullptr; in.hidden_out = nullptr; return in;
}
__device__ __forceinline__ PhaseIn in_plain(const float* x) {
    PhaseIn in; in.kind = IN_PLAIN; in.words = nullptr; in.tag = 0u; in.norm_w = nullptr; in.inv_h = 0.f; in.eps = 0.f;
    in.plain = x; in.hidden_out = nullptr; return in;
}

__device__ LL_FN void gemv_phase(CState& st, const MatD& W, const PhaseIn& in, int epi, u64* ll_out, float* plain_out, uint32_t tag) {
    const LLParams& p = ll_params();
    const LLSmem s = ll_smem();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nkc = W.nkc, rb = W.rb, nloc = W.re - W.rb;
    const int nt = nloc * nkc;
    const bool norm = in.kind == IN_NORM, add = norm && in.words != nullptr;
    float* tout = s.tile_out + st.pp * (LL_MAXT * 16);
    float* ssq = s.ssq + st.pp * 32;
    const float* res_old = s.resid + st.res_par * LL_MAXH;
    float* res_new = s.resid + (add ? (st.res_par ^ 1) : st.res_par) * LL_MAXH;
    uint4* dig = s.dig + warp * 64;
    float* xsum = s.xsum + warp * 4;
    float* xscl = s.xscl + warp * 4;
    // 16 warps, nkc chunks: when nkc divides 16 a warp keeps ONE chunk for the whole phase and warps kc, kc + nkc, ... share
    // it - warp kc polls and converts (into its own area), the others wait on named barrier 2 + kc and read that area.
    const bool share = (LL_CWARPS % nkc) == 0 && nkc < LL_CWARPS;
    // a norm phase walks every chunk even where this CTA has no tile: the residual stream must stay complete
    const int jmax = norm ? (nt > nkc ? nt : nkc) : nt;
    int kc = warp % nkc, cur = -1;           // tile j covers k-chunk j % nkc (j = warp, warp + 16, ...)
    const int kstep = LL_CWARPS % nkc;
    for (int j = warp; j < jmax; j += LL_CWARPS) {
        const bool mine = !share || warp < nkc;          // this warp polls and converts chunk kc itself
        if (kc != cur && !mine) {
            if (j < nt) {
                const int users = ((nt < LL_CWARPS ? nt : LL_CWARPS) - kc + nkc - 1) / nkc;
                asm volatile("bar.sync %0, %1;" ::"r"(2 + kc), "r"(32 * users) : "memory");
                dig = s.dig + kc * 64; xsum = s.xsum + kc * 4; xscl = s.xscl + kc * 4;
            }
            cur = kc;
        }
        if (kc != cur) {
            const int k4a = kc * 64 + lane, k4b = k4a + 32;
            float4 va, vb;
            if (in.kind == IN_PLAIN) {
                va = __ldcg(reinterpret_cast<const float4*>(in.plain) + k4a);
                vb = __ldcg(reinterpret_cast<const float4*>(in.plain) + k4b);
            } else {
                float4 na = make_float4(1.f, 1.f, 1.f, 1.f), nb = na;
                if (norm) {                      // in flight while the words are polled
                    na = __ldg(reinterpret_cast<const float4*>(in.norm_w) + k4a);
                    nb = __ldg(reinterpret_cast<const float4*>(in.norm_w) + k4b);
                }
                va = make_float4(0.f, 0.f, 0.f, 0.f); vb = va;
                if (in.words) {
                    const u64* const pp[2] = {in.words + 4 * (size_t)k4a, in.words + 4 * (size_t)k4b};
                    const bool on[2] = {true, true};
                    float4 o[2];
                    ll_ld4n<2>(pp, on, in.tag, o, p.state);
                    va = o[0]; vb = o[1];
                }
                if (norm) {
                    const float4 ra = reinterpret_cast<const float4*>(res_old)[k4a], rb4 = reinterpret_cast<const float4*>(res_old)[k4b];
                    va.x += ra.x; va.y += ra.y; va.z += ra.z; va.w += ra.w;
                    vb.x += rb4.x; vb.y += rb4.y; vb.z += rb4.z; vb.w += rb4.w;
                    if (add && j < nkc) {        // the one writer of this chunk's new residual
                        reinterpret_cast<float4*>(res_new)[k4a] = va;
                        reinterpret_cast<float4*>(res_new)[k4b] = vb;
                    }
                    float ss = (va.x * va.x + va.y * va.y) + (va.z * va.z + va.w * va.w) + (vb.x * vb.x + vb.y * vb.y) + (vb.z * vb.z + vb.w * vb.w);
                    ss = warp_sum(ss);
                    if (j < nkc && lane == 0) ssq[kc] = ss;
                    va.x *= na.x; va.y *= na.y; va.z *= na.z; va.w *= na.w;
                    vb.x *= nb.x; vb.y *= nb.y; vb.z *= nb.z; vb.w *= nb.w;
                }
            }
            if (j < nt) {
                __syncwarp();                    // the previous tile's reads of the private area are done
                emit_digits(dig, xsum, xscl, va, lane, lane);
                emit_digits(dig, xsum, xscl, vb, lane + 32, lane);
                __syncwarp();
                if (share) {
                    const int users = ((nt < LL_CWARPS ? nt : LL_CWARPS) - kc + nkc - 1) / nkc;
                    if (users > 1) asm volatile("bar.arrive %0, %1;" ::"r"(2 + kc), "r"(32 * users) : "memory");
                }
            }
            cur = kc;
        }
        if (j < nt) {
            const uint32_t i = st.seq + j, slot = i % LL_NSLOT, par = (i / LL_NSLOT) & 1;
            mbar_wait(smem_u32(&s.full[slot]), par, p.state, 0x200u);
            float lo, hi;
            tile_dot(dig, xsum, xscl, s.ring + (size_t)slot * Q3T_TILE_BYTES, lane, lo, hi);
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&s.empty[slot]));
            if ((lane & 3) == 0) { tout[j * 16 + (lane >> 2)] = lo; tout[j * 16 + (lane >> 2) + 8] = hi; }
        }
        kc += kstep; if (kc >= nkc) kc -= nkc;
    }
    st.seq += nt;
    st.pp ^= 1;
    if (add) st.res_par ^= 1;
    LL_STAMP(ST_F_TILES);
    cbar();
    LL_STAMP(ST_F_GBAR);
    float rstd = 1.f;
    if (norm) {
        float t = 0.f;
        for (int k = 0; k < nkc; ++k) t += ssq[k];
        rstd = rsqrtf(t * in.inv_h + in.eps);
        if (in.hidden_out && blockIdx.x == 0) {
            const int H4 = nkc * 64;
            for (int k4 = tid; k4 < H4; k4 += LL_CTHREADS) {
                const float4 x = reinterpret_cast<const float4*>(res_new)[k4];
                const float4 nw = __ldg(reinterpret_cast<const float4*>(in.norm_w) + k4);
                reinterpret_cast<float4*>(in.hidden_out)[k4] = make_float4(nw.x * (x.x * rstd), nw.y * (x.y * rstd), nw.z * (x.z * rstd), nw.w * (x.w * rstd));
            }
        }
    }
    // rows: four lanes per output row add every fourth k-chunk, two shuffles finish the sum (fixed order); a warp
    // covers 8 rows per iteration.  SwiGLU tiles hold gate rows 0..7 and the matching up rows 8..15 (weights interleaved
    // at load): slots 0..3 of a warp take gate rows, slots 4..7 the matching up rows, paired with one more shuffle.
    const int q = lane & 3, sl = lane >> 2;
    const float* bias = W.bias;
    for (int i0 = warp << 3; i0 < nloc * 16; i0 += LL_CTHREADS / 4) {
        int rtl, r;
        if (epi == EPI_SWIGLU) { const int blk = i0 >> 3; rtl = blk >> 1; r = ((blk & 1) << 2) + (sl & 3) + ((sl >> 2) << 3); }
        else { const int i = i0 + sl; rtl = i >> 4; r = i & 15; }
        float v = 0.f;
        for (int k = q; k < nkc; k += 4) v += tout[(rtl * nkc + k) * 16 + r];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v *= rstd;
        if (epi == EPI_SWIGLU) {
            const float u = __shfl_xor_sync(0xffffffffu, v, 16);
            if (sl < 4 && q == 0) ll_st(ll_out + (size_t)(rb + rtl) * 8 + (r & 7), silu_f(v) * u, tag);
        } else if (q == 0) {
            const int n = (rb + rtl) * 16 + r;
            if (bias) v += bias[n];
            ll_st(ll_out + n, v, tag);
            if (plain_out) plain_out[n] = v;
        }
    }
}

// final RMSNorm without a following contraction (stack mode, hidden state only): block-wide, once per launch at most
__device__ LL_FN void final_norm_only(CState& st, const u64* ll_add, uint32_t tag_add, const float* norm_w, float* hidden_out, int H,
                                      float eps) {
    const LLParams& p = ll_params();
    const LLSmem s = ll_smem();
    const int tid = threadIdx.x, H4 = H >> 2;
    const bool on = tid < H4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f), nw = make_float4(0.f, 0.f, 0.f, 0.f);
    if (on) {
        nw = __ldg(reinterpret_cast<const float4*>(norm_w) + tid);
        v = reinterpret_cast<const float4*>(s.resid + st.res_par * LL_MAXH)[tid];
        if (ll_add) {
            const float4 a = ll_ld4(ll_add + 4 * tid, tag_add, p.state);
            v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
        }
    }
    const float ss = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    st.red_par ^= 1;
    const float rstd = rsqrtf(cblock_sum(ss, s.red, st.red_par) / (float)H + eps);
    if (on && hidden_out && blockIdx.x == 0)
        reinterpret_cast<float4*>(hidden_out)[tid] = make_float4(nw.x * (v.x * rstd), nw.y * (v.y * rstd), nw.z * (v.z * rstd), nw.w * (v.w * rstd));
}

// ---- attention geometry (uniform over the grid) -----------------------------------------------------------------------------
__device__ __forceinline__ void attn_geometry(int ctx, int n_kv, int grid, int chunk_min, int split_max, int& chunk, int& nsplit) {
    int maxsplit = grid / n_kv;
    if (maxsplit > split_max) maxsplit = split_max;
    if (maxsplit < 1) maxsplit = 1;
    chunk = chunk_min;
    if (ctx > chunk * maxsplit) {
        chunk = (ctx + maxsplit - 1) / maxsplit;
        chunk = (chunk + Q3T_KV_PAGE - 1) / Q3T_KV_PAGE * Q3T_KV_PAGE;
    }
    nsplit = (ctx + chunk - 1) / chunk;
}

// ---- attention phase: q/k RMSNorm + RoPE + KV-page write + split-KV GQA decode attention -> LL records ----------------------
// K/V staging: the rows of this CTA's (kv head, chunk) are whole 4 KB blocks of the paged cache ([page][k|v][head][16][128]
// bf16), so 64 tokens are eight cp.async.bulk copies issued as ONE warp instruction; the staging buffer is dedicated, so
// the copies of layer l+1 are issued right after the O projection of layer l and nobody computes an address per row.
constexpr int LL_KV_ROUND = 64;                  // tokens staged per round (4 pages)
static_assert(LL_KV_ROUND * 512 <= LL_STAGE_BYTES, "attention K/V staging buffer");
static_assert((LL_MAXSPLIT - 1) * 2 * LL_REC * 4 <= LL_STAGE_BYTES, "parked split records must fit the staging buffer");
constexpr int LL_SAMPLE_MAXV = 3072;             // largest vocabulary the in-kernel sampler holds in the staging buffer
static_assert(LL_SAMPLE_MAXV * (4 + 4 + 2) + 1024 <= LL_STAGE_BYTES && LL_SAMPLE_MAXV <= SAMPLE_MAXV, "sampler scratch must fit the staging buffer");

__device__ __forceinline__ unsigned char* kv_stage(const LLSmem& s) { return s.stage; }

// lanes 0..2*npg-1 of ONE warp (all 32 lanes call): pages [pg0, pg0 + npg) of this CTA's chunk -> staging area, one 4 KB
// block per lane, completion on s.kvbar.  One warp instruction issues all copies; the warp loses ~0.1 us.
__device__ __forceinline__ void kv_issue(const LLSmem& s, const LLStack& S, int layer, int kvh, int pg0, int npg, int lane) {
    if (lane >= 2 * npg) return;
    const int pg = lane >> 1, is_v = lane & 1;
    const size_t page_elems = (size_t)2 * S.n_kv * Q3T_KV_PAGE * 128;
    const __nv_bfloat16* src = S.kv_pool + (size_t)layer * S.kv_layer_stride + (size_t)s.pages[pg0 + pg] * page_elems +
                               (size_t)(is_v * S.n_kv + kvh) * Q3T_KV_PAGE * 128;
    const uint32_t bar = smem_u32(s.kvbar), dst = smem_u32(kv_stage(s)) + pg * 8192 + is_v * 4096;
    // generic-proxy accesses to the staging area (parked records, sampler scratch) were ordered before this thread by a block barrier
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (lane == 0) mbar_expect_tx(bar, (uint32_t)npg * 8192u);   // a copy that lands first only drives the tx count negative
    tma_load_1d(dst, src, 4096, bar);
}

// called by every consumer thread once all reads of the staging buffer are behind a block barrier (round 0 of the chunk)
__device__ __forceinline__ void attn_prefetch(const CState& st, const LLStack& S, int layer, int pos) {
    const int cta = blockIdx.x;
    if (cta >= S.n_kv * st.nsplit || (threadIdx.x >> 5) != LL_CWARPS - 1) return;
    const LLSmem s = ll_smem();
    const int split = cta / S.n_kv, kvh = cta - split * S.n_kv;
    const int s0 = split * st.chunk, s1 = min(pos + 1, s0 + st.chunk);
    const int np = (min(s1 - s0, LL_KV_ROUND) + Q3T_KV_PAGE - 1) / Q3T_KV_PAGE;
    kv_issue(s, S, layer, kvh, 0, np, threadIdx.x & 31);
}

// Attention of a context of at most 32 tokens held by ONE split (every code-predictor pass: <= 17 positions): no block
// barrier at all.  Warps 0..REP-1 each own a q head end to end (poll, RMSNorm, RoPE, one token per lane for the scores with
// the 16-byte chunks walked in a lane-rotated order - conflict-free -, warp softmax, P.V with four output dims per lane,
// normalised words published directly); warps REP / REP+1 write the new k / v row to the cache and into the staged page and
// arrive on a named barrier the q warps wait on; the other twelve warps walk straight on to the O projection.
// Out of line on purpose: inlined, its registers push spills into the contraction phases of the same giant function
// (measured: +9 % per talker step, which never takes this path); the call is made by the four working warps only.
template <int REP>
__device__ __noinline__ void attn_tiny(const LLStack& S, const LayerD& LD, int layer, int pos, const u64* ll_qkv,
                                       uint32_t tag_qkv, u64* ll_attnf, uint32_t tag_out, uint32_t par, const float* plain_qkv) {
    constexpr int D = 128;
    const LLParams& p = ll_params();
    const LLSmem s = ll_smem();
    const int kvh = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = pos + 1;
    float* q_s = s.att;                           // [REP][2][16][4], as in attn_phase
    unsigned char* kv_s = kv_stage(s);
    const bool is_q = warp < REP, is_k = warp == REP;
    float4 nw4 = make_float4(1.f, 1.f, 1.f, 1.f);
    if (is_q || is_k) nw4 = __ldg(reinterpret_cast<const float4*>(is_q ? LD.q_norm : LD.k_norm) + lane);
    const int n0 = (is_q ? (kvh * REP + warp) : (is_k ? (S.n_heads + kvh) : (S.n_heads + S.n_kv + kvh))) * D + lane * 4;
    const float4 xv = plain_qkv ? __ldcg(reinterpret_cast<const float4*>(plain_qkv + n0)) : ll_ld4(ll_qkv + n0, tag_qkv, p.state);
    float x[4] = {xv.x, xv.y, xv.z, xv.w};
    if (is_q || is_k) {
        const float nw[4] = {nw4.x, nw4.y, nw4.z, nw4.w};
        float ss = x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3];
        ss = warp_sum(ss);
        const float rstd = rsqrtf(ss / (float)D + S.eps);
#pragma unroll
        for (int e = 0; e < 4; ++e) x[e] = nw[e] * (x[e] * rstd);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float other = __shfl_xor_sync(0xffffffffu, x[e], 16);
            const float cs = s.cs[(lane & 15) * 4 + e], sn = s.sn[(lane & 15) * 4 + e];
            x[e] = (lane < 16) ? (x[e] * cs - other * sn) : (x[e] * cs + other * sn);
        }
    }
    const uint32_t nbar = (REP + 2) * 32;
    if (!is_q) {
        __nv_bfloat16 hb[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) hb[e] = __float2bfloat16_rn(x[e]);
        const size_t page_elems = (size_t)2 * S.n_kv * Q3T_KV_PAGE * D;
        __nv_bfloat16* dst = S.kv_pool + (size_t)layer * S.kv_layer_stride + (size_t)s.pages[pos / Q3T_KV_PAGE] * page_elems +
                             (size_t)kvh * Q3T_KV_PAGE * D + (is_k ? 0 : (size_t)S.n_kv * Q3T_KV_PAGE * D) +
                             (size_t)(pos % Q3T_KV_PAGE) * D + lane * 4;
        mbar_wait(smem_u32(s.kvbar), par, p.state, 0x600u);
        *reinterpret_cast<uint2*>(kv_s + (pos >> 4) * 8192 + (is_k ? 0 : 4096) + (pos & 15) * 256 + lane * 8) =
            *reinterpret_cast<const uint2*>(hb);
        asm volatile("bar.arrive 10, %0;" ::"r"(nbar) : "memory");
        // the cache write is off the critical path; the proxy fence orders it before the cp.async.bulk read of the next pass
        *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<const uint2*>(hb);
        asm volatile("fence.proxy.async.global;" ::: "memory");
        return;
    }
    {
        const float sc = rsqrtf((float)D);
        *reinterpret_cast<float4*>(q_s + warp * D + (((lane & 1) << 4) + (lane >> 1)) * 4) =
            make_float4(x[0] * sc, x[1] * sc, x[2] * sc, x[3] * sc);
    }
    __syncwarp();
    mbar_wait(smem_u32(s.kvbar), par, p.state, 0x600u);
    asm volatile("bar.sync 10, %0;" ::"r"(nbar) : "memory");
    // scores: lane = token
    float sa = 0.f;
    {
        const unsigned char* krow = kv_s + (lane >> 4) * 8192 + (lane & 15) * 256;
        const float* qh = q_s + warp * D;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int c = (i + lane) & 15;
            const uint4 kk = *reinterpret_cast<const uint4*>(krow + c * 16);
            const float4 qa = *reinterpret_cast<const float4*>(qh + c * 4);
            const float4 qb = *reinterpret_cast<const float4*>(qh + (16 + c) * 4);
            sa = fmaf(qa.x, bf16lo(kk.x), fmaf(qa.y, bf16hi(kk.x), fmaf(qa.z, bf16lo(kk.y), fmaf(qa.w, bf16hi(kk.y), sa))));
            sa = fmaf(qb.x, bf16lo(kk.z), fmaf(qb.y, bf16hi(kk.z), fmaf(qb.z, bf16lo(kk.w), fmaf(qb.w, bf16hi(kk.w), sa))));
        }
    }
    if (lane >= n) sa = -INFINITY;                // stale slots may hold anything: select, never arithmetic
    float m = sa;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    const float pj = __expf(sa - m);
    const float l = warp_sum(pj);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const unsigned char* vcol = kv_s + 4096 + lane * 8;
    for (int t = 0; t < n; ++t) {
        const float pv = __shfl_sync(0xffffffffu, pj, t);
        const uint2 vv = *reinterpret_cast<const uint2*>(vcol + (t >> 4) * 8192 + (t & 15) * 256);
        acc[0] = fmaf(pv, bf16lo(vv.x), acc[0]); acc[1] = fmaf(pv, bf16hi(vv.x), acc[1]);
        acc[2] = fmaf(pv, bf16lo(vv.y), acc[2]); acc[3] = fmaf(pv, bf16hi(vv.y), acc[3]);
    }
    const float il = 1.f / l;
    u64* fin = ll_attnf + (size_t)(kvh * REP + warp) * D + lane * 4;
#pragma unroll
    for (int e = 0; e < 4; ++e) ll_st(fin + e, acc[e] * il, tag_out);
}

template <int REP>
__device__ LL_FN void attn_phase(CState& st, const LLStack& S, const LayerD& LD, int layer, int pos, const u64* ll_qkv,
                                 uint32_t tag_qkv, u64* ll_attn, u64* ll_attnf, uint32_t tag_out, const float* plain_qkv) {
    constexpr int D = 128;
    constexpr int TG = LL_CWARPS / (2 * REP), TPG = LL_KV_ROUND / TG;   // token groups per round / tokens per group
    const LLParams& p = ll_params();
    const LLSmem s = ll_smem();
    const int chunk = st.chunk, nsplit = st.nsplit;
    const int cta = blockIdx.x;
    if (cta >= S.n_kv * nsplit) return;
#ifndef LL_NO_TINY
    if (nsplit == 1 && pos < 32) {
        const uint32_t par = st.kv_par;
        st.kv_par ^= 1;                           // every thread keeps the parity; only the working warps wait
        if ((threadIdx.x >> 5) < REP + 2) attn_tiny<REP>(S, LD, layer, pos, ll_qkv, tag_qkv, ll_attnf, tag_out, par, plain_qkv);
        return;
    }
#endif
    const int split = cta / S.n_kv, kvh = cta - split * S.n_kv;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ctx = pos + 1, s0 = split * chunk, s1 = min(ctx, s0 + chunk), n = s1 - s0;
    const bool owner = (pos >= s0 && pos < s1);
    float* q_s = s.att;                           // [REP][2][16][4]: q pre-scaled by 1/sqrt(D), 16-byte chunks of a K row apart
    float* sc_s = s.att + 2 * D;                  // [REP][64] scores of the round
    __nv_bfloat16* new_s = reinterpret_cast<__nv_bfloat16*>(s.att + 3 * D);   // [2][D]: k, v of the new token as stored
    float* part_s = s.att + 4 * D;                // [TG][REP][D] partial outputs, then [TG][REP] partial sums, then [REP] max
    float* lpart_s = part_s + TG * REP * D;
    float* m_s = lpart_s + TG * REP;
    unsigned char* kv_s = kv_stage(s);
    const int n_rounds = (n + LL_KV_ROUND - 1) / LL_KV_ROUND;
    const int rd_own = owner ? (pos - s0) / LL_KV_ROUND : -1;

    // 1. q heads of this kv head (+ k, v of the new token on the split that owns it)
    __nv_bfloat16* kv_dst = nullptr;
    if (warp < REP + 2) {
        const bool is_q = warp < REP, is_k = warp == REP;
        if (is_q || owner) {
            float4 nw4 = make_float4(1.f, 1.f, 1.f, 1.f);
            if (is_q || is_k) nw4 = __ldg(reinterpret_cast<const float4*>(is_q ? LD.q_norm : LD.k_norm) + lane);
            const int n0 = (is_q ? (kvh * REP + warp) : (is_k ? (S.n_heads + kvh) : (S.n_heads + S.n_kv + kvh))) * D + lane * 4;
            const float4 xv = plain_qkv ? __ldcg(reinterpret_cast<const float4*>(plain_qkv + n0)) : ll_ld4(ll_qkv + n0, tag_qkv, p.state);
            LL_STAMP(ST_F_ATT_B);   // B: q words arrived
            float x[4] = {xv.x, xv.y, xv.z, xv.w};
            if (is_q || is_k) {
                const float nw[4] = {nw4.x, nw4.y, nw4.z, nw4.w};
                float ss = x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3];
                ss = warp_sum(ss);
                const float rstd = rsqrtf(ss / (float)D + S.eps);
#pragma unroll
                for (int e = 0; e < 4; ++e) x[e] = nw[e] * (x[e] * rstd);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float other = __shfl_xor_sync(0xffffffffu, x[e], 16);
                    const float cs = s.cs[(lane & 15) * 4 + e], sn = s.sn[(lane & 15) * 4 + e];
                    x[e] = (lane < 16) ? (x[e] * cs - other * sn) : (x[e] * cs + other * sn);
                }
            }
            if (is_q) {
                const float sc = rsqrtf((float)D);
                // element lane*4+k sits in 16-byte chunk lane>>1 of a K row, half lane&1 of that chunk
                *reinterpret_cast<float4*>(q_s + warp * D + (((lane & 1) << 4) + (lane >> 1)) * 4) =
                    make_float4(x[0] * sc, x[1] * sc, x[2] * sc, x[3] * sc);
            } else {
                __nv_bfloat16 hb[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) hb[e] = __float2bfloat16_rn(x[e]);
                const size_t page_elems = (size_t)2 * S.n_kv * Q3T_KV_PAGE * D;
                __nv_bfloat16* dst = S.kv_pool + (size_t)layer * S.kv_layer_stride + (size_t)s.pages[(pos - s0) / Q3T_KV_PAGE] * page_elems +
                                     (size_t)kvh * Q3T_KV_PAGE * D + (is_k ? 0 : (size_t)S.n_kv * Q3T_KV_PAGE * D) +
                                     (size_t)(pos % Q3T_KV_PAGE) * D + lane * 4;
                *reinterpret_cast<uint2*>(new_s + (is_k ? 0 : D) + lane * 4) = *reinterpret_cast<const uint2*>(hb);
                kv_dst = dst;                     // the cache write itself waits until the scores are out (off the critical path)
            }
        }
    }
    LL_STAMP(ST_F_ATT_C);   // C: q normalised, rotated, stored
    // 2. rounds of 64 staged tokens: scores by 8 threads per token (both heads), then per warp (head, 64-dim half, token
    //    group): max of the round, probabilities of the group's tokens in lanes, P.V with one shuffle per token
    const int r_w = warp % REP, dh = (warp / REP) & 1, tg = warp / (2 * REP);
    float m_run = -INFINITY, l_run = 0.f, acc0 = 0.f, acc1 = 0.f;
    for (int rd = 0; rd < n_rounds; ++rd) {
        const int nr = min(n - rd * LL_KV_ROUND, LL_KV_ROUND);          // valid tokens of this round
        if (rd > 0 && warp == LL_CWARPS - 1) kv_issue(s, S, layer, kvh, rd * (LL_KV_ROUND / Q3T_KV_PAGE), (nr + Q3T_KV_PAGE - 1) / Q3T_KV_PAGE, lane);
        mbar_wait(smem_u32(s.kvbar), st.kv_par, p.state, 0x600u);
        st.kv_par ^= 1;
        if (rd == rd_own && (warp == REP || warp == REP + 1)) {         // the new token's rows replace the stale slot;
            const int tl = pos - s0 - rd * LL_KV_ROUND, which = warp - REP;  // every lane copies what it stored itself
            *reinterpret_cast<uint2*>(kv_s + (tl >> 4) * 8192 + which * 4096 + (tl & 15) * 256 + lane * 8) =
                *reinterpret_cast<const uint2*>(new_s + which * D + lane * 4);
        }
        LL_STAMP(ST_F_ATT_A);   // A: staged rows present
        cbar();
        LL_STAMP(ST_F_ATT_D);   // D: barrier
        {
            const int tok = tid >> 3, o = tid & 7;
            const unsigned char* krow = kv_s + (tok >> 4) * 8192 + (tok & 15) * 256;
            float sa[REP];
#pragma unroll
            for (int r = 0; r < REP; ++r) sa[r] = 0.f;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int c = i * 8 + o;
                const uint4 kk = *reinterpret_cast<const uint4*>(krow + c * 16);
                const float k0 = bf16lo(kk.x), k1 = bf16hi(kk.x), k2 = bf16lo(kk.y), k3 = bf16hi(kk.y);
                const float k4 = bf16lo(kk.z), k5 = bf16hi(kk.z), k6 = bf16lo(kk.w), k7 = bf16hi(kk.w);
#pragma unroll
                for (int r = 0; r < REP; ++r) {
                    const float4 qa = *reinterpret_cast<const float4*>(q_s + r * D + c * 4);
                    const float4 qb = *reinterpret_cast<const float4*>(q_s + r * D + (16 + c) * 4);
                    sa[r] = fmaf(qa.x, k0, fmaf(qa.y, k1, fmaf(qa.z, k2, fmaf(qa.w, k3, sa[r]))));
                    sa[r] = fmaf(qb.x, k4, fmaf(qb.y, k5, fmaf(qb.z, k6, fmaf(qb.w, k7, sa[r]))));
                }
            }
#pragma unroll
            for (int r = 0; r < REP; ++r) {
                sa[r] += __shfl_xor_sync(0xffffffffu, sa[r], 1);
                sa[r] += __shfl_xor_sync(0xffffffffu, sa[r], 2);
                sa[r] += __shfl_xor_sync(0xffffffffu, sa[r], 4);
                if (o == 0) sc_s[r * LL_KV_ROUND + tok] = tok < nr ? sa[r] : -INFINITY;   // stale slots may hold anything
            }
        }
        cbar();
        LL_STAMP(ST_F_ATT_E);   // E: scores
        {
            float mr = fmaxf(sc_s[r_w * LL_KV_ROUND + lane], sc_s[r_w * LL_KV_ROUND + lane + 32]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mr = fmaxf(mr, __shfl_xor_sync(0xffffffffu, mr, o));
            const float mn = fmaxf(m_run, mr);                           // finite: every round has a valid token
            const float corr = __expf(m_run - mn);
            const float pj = lane < TPG ? __expf(sc_s[r_w * LL_KV_ROUND + tg * TPG + lane] - mn) : 0.f;
            l_run = l_run * corr + warp_sum(pj);
            acc0 *= corr; acc1 *= corr;
            m_run = mn;
            const unsigned char* vcol = kv_s + 4096 + dh * 128 + lane * 4;
            if (tg * TPG < nr) {                                         // (a code-predictor pass never has more than 17 tokens)
#pragma unroll
            for (int j = 0; j < TPG; ++j) {
                // unconditional load (the 16 of them pipeline); a stale slot may hold anything, its weight is exactly zero
                const int t = tg * TPG + j;
                const float pv = __shfl_sync(0xffffffffu, pj, j);
                const uint32_t vv = *reinterpret_cast<const uint32_t*>(vcol + (t >> 4) * 8192 + (t & 15) * 256);
                acc0 = fmaf(pv, t < nr ? bf16lo(vv) : 0.f, acc0);
                acc1 = fmaf(pv, t < nr ? bf16hi(vv) : 0.f, acc1);
            }
            }
        }
        if (rd + 1 < n_rounds) cbar();                                   // staged rows and scores are read
    }
    *reinterpret_cast<float2*>(part_s + (tg * REP + r_w) * D + dh * 64 + lane * 2) = make_float2(acc0, acc1);
    if (dh == 0 && lane == 0) { lpart_s[tg * REP + r_w] = l_run; if (tg == 0) m_s[r_w] = m_run; }
    LL_STAMP(ST_F_ATT_F);   // F: partial outputs stored
    cbar();
    LL_STAMP(ST_F_ATT_G);   // G: barrier
    if (kv_dst) {           // new k / v row -> cache.  A later pass of this launch (code predictor) reads the page back with
                            // cp.async.bulk: the proxy fence orders the store before it (block barriers order the threads)
        *reinterpret_cast<uint2*>(kv_dst) = *reinterpret_cast<const uint2*>(new_s + (warp == REP ? 0 : D) + lane * 4);
        asm volatile("fence.proxy.async.global;" ::: "memory");
    }
    // 3. this CTA's partial for (head r, dim d): the token groups summed out of shared memory (all share one max)
    const bool mine = tid < REP * D;
    const int r = tid / D, d = tid % D;
    float M = -INFINITY, L = 0.f, A = 0.f;
    if (mine) {
        M = m_s[r];
#pragma unroll
        for (int g = 0; g < TG; ++g) { A += part_s[(g * REP + r) * D + d]; L += lpart_s[g * REP + r]; }
    }
    u64* fin = ll_attnf + (size_t)(kvh * REP + r) * D + d;
    if (nsplit == 1) {           // the only split of its kv head: publish the normalised output directly
        if (mine) ll_st(fin, A * (1.f / L), tag_out);
        return;
    }
    if (split != 0) {            // partial record for the merger (split 0 of the same kv head)
        if (mine) {
            u64* rec = ll_attn + ((size_t)(kvh * REP + r) * LL_MAXSPLIT + split) * LL_REC;
            ll_st(rec + d, A, tag_out);
            if (d == 0) { ll_st(rec + D, M, tag_out); ll_st(rec + D + 1, L, tag_out); }
        }
        return;
    }
    // 6. merger: the records of the other splits are requested by all 512 threads at once (one 16-byte load per thread
    //    covers four splits of two heads -> one L2 round trip at ctx <= 320), parked in the staging buffer (its K/V rows are
    //    dead), merged with this CTA's own partial and published as plain normalised words: every consumer of the O
    //    projection then needs ONE poll of its own words instead of nsplit dependent record reads of lines that 148 CTAs
    //    hammer at once.
    float* tmp = reinterpret_cast<float*>(s.stage);
    auto tmp_at = [&](int f) -> float& { return tmp[f]; };
    const int n_items = (nsplit - 1) * REP * 64;
    for (int it0 = 0; it0 < n_items; it0 += LL_CTHREADS) {
        const int it = it0 + tid;
        if (it < n_items) {
            const int spi = it / (REP * 64), rem = it - spi * (REP * 64), rr = rem >> 6, j = rem & 63;
            const u64* rec = ll_attn + ((size_t)(kvh * REP + rr) * LL_MAXSPLIT + spi + 1) * LL_REC;
            const bool ml = (j == 0);
            u64 a0, a1, m0 = 0, l0 = 0;
            ll_ld2(rec + 2 * j, a0, a1);
            if (ml) ll_ld2(rec + D, m0, l0);
            bool oka = false, okm = !ml;
            int spins = 0;
            for (;;) {
                if (!oka) oka = ll_ok(a0, tag_out) && ll_ok(a1, tag_out);
                if (!okm) okm = ll_ok(m0, tag_out) && ll_ok(l0, tag_out);
                if (oka && okm) break;
                if (++spins > LL_SPIN_LIMIT) ll_fail(p.state, 0x300u);
                if (!oka) ll_ld2(rec + 2 * j, a0, a1);
                if (!okm) ll_ld2(rec + D, m0, l0);
            }
            const int f0 = (spi * REP + rr) * LL_REC;
            tmp_at(f0 + 2 * j) = ll_val(a0); tmp_at(f0 + 2 * j + 1) = ll_val(a1);
            if (ml) { tmp_at(f0 + D) = ll_val(m0); tmp_at(f0 + D + 1) = ll_val(l0); }
        }
    }
    LL_STAMP(ST_F_MERGE);
    cbar();
    if (mine) {
        float Mx = M;
        for (int sp = 0; sp < nsplit - 1; ++sp) Mx = fmaxf(Mx, tmp_at((sp * REP + r) * LL_REC + D));
        const float w0 = __expf(M - Mx);
        float Ls = L * w0, As = A * w0;
        for (int sp = 0; sp < nsplit - 1; ++sp) {
            const int f0 = (sp * REP + r) * LL_REC;
            const float wt = __expf(tmp_at(f0 + D) - Mx);
            Ls = fmaf(tmp_at(f0 + D + 1), wt, Ls);
            As = fmaf(tmp_at(f0 + d), wt, As);
        }
        ll_st(fin, As * (1.f / Ls), tag_out);
    }
}

// ---- one token through a dense stack (consumers) ------------------------------------------------------------------------------
// The current half of the residual buffer must hold the stack input (or zeros when first_add carries it as LL words).
// Returns the words the caller's final norm still has to add (the last down projection).
struct StackOut { const u64* add; uint32_t tag; };

// qkv0 != nullptr: q|k|v of the first layer come from a table row (plain fp32), its QKV phase is skipped
// att_only: this CTA runs only the attention of the pass (code-predictor passes: the kv-head CTAs hold no row tiles of the
// code predictor, so they poll their q|k|v words the moment they are published instead of after a QKV epilogue of their own:
// 86-88 -> 82-83 us per pass, profiles/r02_ll_experiments.txt).  Phase tags advance identically on every CTA.
__device__ LL_FN StackOut stack_consume(CState& st, const LLStack& S, const LayerD* lay, int pos, const u64* first_add, uint32_t first_tag,
                                        const float* qkv0 = nullptr, bool att_only = false) {
    const LLParams& p = ll_params();
    const LLSmem s = ll_smem();
    const int tid = threadIdx.x;
    const int rep = S.n_heads / S.n_kv;
    // RoPE table of this position (rotate_half convention, fp32 cos/sin as the oracle)
    if (tid < S.head_dim / 2) {
        float sn, cs;
        sincosf((float)pos * S.inv_freq[tid], &sn, &cs);
        s.cs[tid] = cs; s.sn[tid] = sn;
    }
    attn_geometry(pos + 1, S.n_kv, gridDim.x, p.att_chunk, p.att_maxsplit, st.chunk, st.nsplit);
    if ((int)blockIdx.x < S.n_kv * st.nsplit) {
        // page ids of this CTA's attention chunk: constant during the pass, so no attention phase starts with a
        // dependent block-table load in front of its K/V copies
        const int split = blockIdx.x / S.n_kv, p0 = (split * st.chunk) / Q3T_KV_PAGE;
        const int np = min(st.chunk / Q3T_KV_PAGE, (pos / Q3T_KV_PAGE) - p0 + 1);
        if (tid < np && tid < 64) s.pages[tid] = S.block_tbl[p0 + tid];
    }
    cbar();                                       // page ids, RoPE table and the caller's residual are visible to every warp
    attn_prefetch(st, S, 0, pos);
    const u64* add = first_add;
    uint32_t add_tag = first_tag;
    for (int l = 0; l < S.n_layers; ++l) {
        const LayerD& L = lay[l];
        // ---- QKV
        const float* plain_qkv = l == 0 ? qkv0 : nullptr;
        uint32_t t_qkv = 0u;
        if (!plain_qkv) {
            t_qkv = ++st.gen;
            if (!att_only) gemv_phase(st, L.qkv, in_norm(add, add_tag, L.input_norm, S.hidden, S.eps), EPI_RAW, p.x_qkv, nullptr, t_qkv);
        }
        LL_STAMP(ST_QKV);
        // ---- attention (first n_kv*nsplit CTAs)
        const uint32_t t_att = ++st.gen;
        if (rep == 2) attn_phase<2>(st, S, L, l, pos, p.x_qkv, t_qkv, p.x_attn, p.x_attnf, t_att, plain_qkv);
        else attn_phase<1>(st, S, L, l, pos, p.x_qkv, t_qkv, p.x_attn, p.x_attnf, t_att, plain_qkv);
        LL_STAMP(ST_ATTN);
        if (att_only) {
            // no contraction on this CTA: a barrier of its own puts the reads of the staging buffer behind the next layer's copies
            if (l + 1 < S.n_layers) { cbar(); attn_prefetch(st, S, l + 1, pos); }
            st.gen += 3;
            add = p.x_down; add_tag = st.gen;
            continue;
        }
        // ---- O projection
        const uint32_t t_o = ++st.gen;
        gemv_phase(st, L.o, in_ll(p.x_attnf, t_att), EPI_RAW, p.x_o, nullptr, t_o);
        LL_STAMP(ST_O);
        // every read of the staging buffer (K/V rows, parked records) is behind the barrier of the O projection: the next
        // layer's rows stream in under gate/up, down and QKV
        if (l + 1 < S.n_layers) attn_prefetch(st, S, l + 1, pos);
        // ---- gate/up (+ SwiGLU in the epilogue)
        const uint32_t t_act = ++st.gen;
        gemv_phase(st, L.gu, in_norm(p.x_o, t_o, L.post_norm, S.hidden, S.eps), EPI_SWIGLU, p.x_act, nullptr, t_act);
        LL_STAMP(ST_GU);
        // ---- down
        const uint32_t t_down = ++st.gen;
        gemv_phase(st, L.down, in_ll(p.x_act, t_act), EPI_RAW, p.x_down, nullptr, t_down);
        LL_STAMP(ST_DOWN);
        add = p.x_down; add_tag = t_down;
    }
    StackOut o; o.add = add; o.tag = add_tag;
    return o;
}

// ---- producer side: the same program, streaming instead of computing ---------------------------------------------------------
struct Producer {
    unsigned int* state; int lane; uint32_t seq; int pf_dist;
    // lane ids 0..LL_PLANES-1 (producer warp): TMA copies into the ring; id LL_PLANES (lane 0 of the prefetch warp, its own
    // warp so that its spin does not time-slice with the TMA lanes): L2 prefetch of the same tiles, matrix by matrix, as long
    // as the matrix starts no more than pf_dist tiles ahead of the TMA cursor - HBM keeps streaming while the ring is full
    __device__ LL_FN void stream(const MatD& W) {
        const LLSmem s = ll_smem();
        const uint8_t* src = W.w + (size_t)W.rb * W.nkc * Q3T_TILE_BYTES;
        const int nt = (W.re - W.rb) * W.nkc;
        volatile int* issued = s.ibuf + 63;
        if (lane == LL_PLANES) {
            if (nt > 0) {
                int spins = 0;
                while ((int)seq - *issued > pf_dist)
                    if (++spins > LL_SPIN_LIMIT) ll_fail(state, 0x500u);
                for (int t = 0; t < nt; t += 4) {
                    const int n = nt - t < 4 ? nt - t : 4;
                    l2_prefetch_bulk(src + (size_t)t * Q3T_TILE_BYTES, (uint32_t)n * Q3T_TILE_BYTES);
                }
            }
            seq += nt;
            return;
        }
        for (int t = 0; t < nt; ++t, ++seq, src += Q3T_TILE_BYTES) {
            if ((int)(seq % LL_PLANES) != lane) continue;
            const uint32_t slot = seq % LL_NSLOT, par = (seq / LL_NSLOT) & 1;
            mbar_wait(smem_u32(&s.empty[slot]), par ^ 1, state, 0x400u);
            const uint32_t fb = smem_u32(&s.full[slot]);
            mbar_expect_tx(fb, Q3T_TILE_BYTES);
            tma_load_1d(smem_u32(s.ring + (size_t)slot * Q3T_TILE_BYTES), src, Q3T_TILE_BYTES, fb);
            if (lane == 0) *issued = (int)seq;
        }
    }
    __device__ __forceinline__ void stack(const LayerD* lay, int n_layers, bool skip_qkv0 = false) {
        for (int l = 0; l < n_layers; ++l) {
            if (!(skip_qkv0 && l == 0)) stream(lay[l].qkv);
            stream(lay[l].o); stream(lay[l].gu); stream(lay[l].down);
        }
    }
};

// ---- greedy fast path: argmax straight out of the polled words - no score array, ONE block barrier -------------------------
// Same choice as sample_core's greedy branch (largest score, lowest index on ties; 0 when every score is -inf or NaN).
// The two-level reduction scratch is double buffered by `par` (the caller flips it), so no barrier is needed after the read.
__device__ __noinline__ int sample_greedy(const float* plain, const u64* ll, uint32_t tag, int V, const q3t_sampling& sp,
                                          const unsigned int* seen, int step, int par) {
    const LLParams& p = ll_params();
    const LLSmem s = ll_smem();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float bv = -INFINITY; int bi = 0x7fffffff;
    for (int k4 = tid; k4 < (V >> 2); k4 += LL_CTHREADS) {
        const float4 v = plain ? __ldcg(reinterpret_cast<const float4*>(plain) + k4) : ll_ld4(ll + 4 * (size_t)k4, tag, p.state);
        const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int i = (k4 << 2) + e;
            const float sc = sample_score(x[e], i, sp, seen, step);
            if (sc > bv || (sc == bv && i < bi)) { bv = sc; bi = i; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    float* rf = s.red + par * 32;
    int* ri = s.ibuf + par * 16;
    if (lane == 0) { rf[warp] = bv; ri[warp] = bi; }
    cbar();
    bv = lane < LL_CWARPS ? rf[lane] : -INFINITY;
    bi = lane < LL_CWARPS ? ri[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    return bi == 0x7fffffff ? 0 : bi;
}

// ---- in-kernel sampler (every CTA computes the same choice; CTA 0 records it) ------------------------------------------------
// scores come from plain logits (previous launch) or from LL words of the head GEMV of this launch
__device__ __noinline__ int sample_here(const float* plain, const u64* ll, uint32_t tag, int V, const q3t_sampling& sp,
                                 const unsigned int* seen, int step, int g) {
    const LLParams& p = ll_params();
    const LLSmem s = ll_smem();
    const int tid = threadIdx.x;
    float* sc = reinterpret_cast<float*>(s.stage);            // [V]  (the attention staging buffer is idle between passes)
    float* pe = sc + LL_SAMPLE_MAXV;                           // [V]
    unsigned short* cand = reinterpret_cast<unsigned short*>(pe + LL_SAMPLE_MAXV);   // [V]
    SampleScratch scr;
    scr.sc = sc; scr.pe = pe; scr.cand = cand;
    scr.hist = reinterpret_cast<unsigned int*>(cand + LL_SAMPLE_MAXV);               // [256]
    scr.redf = s.red; scr.redi = s.ibuf; scr.sh_i = s.ibuf + 32;
    const int V4 = V >> 2;
    for (int k4 = tid; k4 < V4; k4 += LL_CTHREADS) {
        float4 v;
        if (plain) v = __ldcg(reinterpret_cast<const float4*>(plain) + k4);
        else v = ll_ld4(ll + 4 * (size_t)k4, tag, p.state);
        const int i = k4 << 2;
        sc[i] = sample_score(v.x, i, sp, seen, step);
        sc[i + 1] = sample_score(v.y, i + 1, sp, seen, step);
        sc[i + 2] = sample_score(v.z, i + 2, sp, seen, step);
        sc[i + 3] = sample_score(v.w, i + 3, sp, seen, step);
    }
    cbar();
    const float u = sp.do_sample ? hash_uniform(sp.seed, step, g, 0) : 0.f;
    const int choice = sample_core<LL_CTHREADS>(scr, V, sp, u, tid, [] { cbar(); });
    cbar();                                    // scratch reads done before the next pass stages K/V rows there
    return choice;
}

// one code-predictor pass: projected input -> 5 layers (-> head -> sampled code)
// prow != nullptr: the projected input row is a table lookup (no projection phase in this pass)
__device__ __noinline__ int cp_pass(CState& st, const float* src, const float* prow, const float* qrow, int pos, int g_head, int step) {
    const LLParams& p = ll_params();
    const LLSmem s = ll_smem();
    const int tid = threadIdx.x, Hc = p.cp.hidden, G = p.n_groups;
    const u64* first_add = nullptr;
    uint32_t t_proj = 0u;
    // the kv-head CTAs of the code predictor (its contexts never leave the one-CTA-per-kv-head path) run its attention only: they
    // hold no row tiles of its matrices (see the kernel: build_mat with cp_skip)
    const bool att_only = (int)blockIdx.x < p.cp.n_kv && (int)gridDim.x > 2 * p.cp.n_kv;
    if (prow) {
        if (!att_only)
            for (int k4 = tid; k4 < (Hc >> 2); k4 += LL_CTHREADS)
                reinterpret_cast<float4*>(s.resid + st.res_par * LL_MAXH)[k4] = __ldcg(reinterpret_cast<const float4*>(prow) + k4);
    } else {
        t_proj = ++st.gen;
        if (!att_only) {
            gemv_phase(st, s.hd[0], in_plain(src), EPI_RAW, p.x_proj, nullptr, t_proj);
            for (int k4 = tid; k4 < (Hc >> 2); k4 += LL_CTHREADS)
                reinterpret_cast<float4*>(s.resid + st.res_par * LL_MAXH)[k4] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        first_add = p.x_proj;
    }
    const StackOut so = stack_consume(st, p.cp, s.lay + p.talker.n_layers, pos, first_add, t_proj, prow ? qrow : nullptr, att_only);
    if (g_head < 0) return 0;
    const uint32_t t_head = ++st.gen;
    float* lg = p.cp_logits ? (p.keep_cp_logits ? p.cp_logits + (size_t)g_head * p.cp_vocab : p.cp_logits) : nullptr;
    if (!att_only) gemv_phase(st, s.hd[2 + g_head], in_norm(so.add, so.tag, p.cp.final_norm, Hc, p.cp.eps), EPI_RAW, p.x_head, lg, t_head);
    int c;
    if (!p.cp_sp.do_sample) { st.red_par ^= 1; c = sample_greedy(nullptr, p.x_head, t_head, p.cp_vocab, p.cp_sp, nullptr, step, st.red_par); }
    else c = sample_here(nullptr, p.x_head, t_head, p.cp_vocab, p.cp_sp, nullptr, step, g_head + 1);
    const long long fo = (long long)step * G + g_head + 1;
    const bool rec = (blockIdx.x == 0 && tid == 0);
    if (rec && p.own_codes && step < p.max_frames) p.own_codes[fo] = c;
    if (p.forced) c = p.forced[fo];
    if (rec) {
        p.cur_codes[g_head + 1] = c;
        if (step < p.max_frames) p.codes[fo] = c;
    }
    return c;
}

// cta / grid = index / count among the CTAs that share the matrix (cta < 0: this CTA holds none of its row tiles)
__device__ __forceinline__ void build_mat(MatD& d, const q3t_w8& w, int cta, int grid) {
    d.w = reinterpret_cast<const uint8_t*>(w.w); d.bias = w.lin_bias; d.nkc = w.K >> 8; d.N = w.N;
    const unsigned nrt = (unsigned)(w.N >> 4);
    if (cta < 0) { d.rb = d.re = 0; return; }
    d.rb = (int)((nrt * (unsigned)cta) / (unsigned)grid);
    d.re = (int)((nrt * (unsigned)(cta + 1)) / (unsigned)grid);
}
__device__ __forceinline__ void build_layer(LayerD& d, const q3t_layer& L, int cta, int grid) {
    build_mat(d.qkv, L.qkv, cta, grid); build_mat(d.o, L.o, cta, grid);
    build_mat(d.gu, L.gate_up, cta, grid); build_mat(d.down, L.down, cta, grid);
    d.input_norm = L.input_norm; d.post_norm = L.post_norm; d.q_norm = L.q_norm; d.k_norm = L.k_norm;
}

// ---- the kernel ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LL_THREADS, 1) frame_ll_kernel(const LLParams p_in) {
    static_assert(sizeof(LLParams) <= LL_OFF_LAY, "LLParams must fit the first KB of shared memory");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, cta = blockIdx.x;
    {   // launch parameters -> shared memory (device functions read them through ll_params())
        const uint32_t* src = reinterpret_cast<const uint32_t*>(&p_in);
        uint32_t* dst = reinterpret_cast<uint32_t*>(ll_smem_raw);
        for (int i = tid; i < (int)(sizeof(LLParams) / 4); i += LL_THREADS) dst[i] = src[i];
    }
    const LLParams& p = ll_params();
    const LLSmem s = ll_smem();
    const int G = p_in.n_groups, nA = p_in.talker.n_layers, nB = p_in.mode == LL_MODE_FRAME ? p_in.cp.n_layers : 0;
    // descriptor tables (this CTA's row-tile ranges included)
    if (tid < nA) build_layer(s.lay[tid], p_in.talker.layers[tid], cta, gridDim.x);
    // code-predictor matrices are dealt over the CTAs that are NOT its kv-head (attention) CTAs
    const int cp_skip = (int)gridDim.x > 2 * p_in.cp.n_kv ? p_in.cp.n_kv : 0;
    if (tid >= nA && tid < nA + nB) build_layer(s.lay[tid], p_in.cp.layers[tid - nA], cta - cp_skip, (int)gridDim.x - cp_skip);
    if (p_in.mode == LL_MODE_FRAME) {
        if (tid == 64) build_mat(s.hd[0], p_in.cp_proj, cta - cp_skip, (int)gridDim.x - cp_skip);
        if (tid == 65) build_mat(s.hd[1], p_in.codec_head, cta, gridDim.x);
        if (tid >= 66 && tid < 66 + G - 1) build_mat(s.hd[2 + tid - 66], p_in.cp_heads[tid - 66], cta - cp_skip, (int)gridDim.x - cp_skip);
        if (p_in.cp_proj_rows && tid >= 96 && tid < 96 + G - 1) s.prows[tid - 96] = p_in.cp_proj_rows[tid - 96];
        if (p_in.cp_qkv0_rows && tid >= 128 && tid < 128 + G - 1) s.qrows[tid - 128] = p_in.cp_qkv0_rows[tid - 128];
    } else if (tid == 65 && p_in.head.w) build_mat(s.hd[1], p_in.head, cta, gridDim.x);
    if (tid == 0) {
        for (int i = 0; i < LL_NSLOT; ++i) { mbar_init(smem_u32(&s.full[i]), 1); mbar_init(smem_u32(&s.empty[i]), 1); }
        mbar_init(smem_u32(s.kvbar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid == 0) s.ibuf[63] = 0;
    __syncthreads();

    if (warp >= LL_CWARPS) {
        // =========================== producer / L2 prefetcher ================================================
        const int role = warp == LL_CWARPS ? (lane < LL_PLANES ? lane : -1) : (lane == 0 && p.pf_dist > 0 ? LL_PLANES : -1);
        if (role >= 0) {
            Producer pr{p.state, role, 0u, p.pf_dist};
            if (p.mode == LL_MODE_STACK) {
                pr.stack(s.lay, nA);
                if (p.head.w) pr.stream(s.hd[1]);
            } else {
                pr.stream(s.hd[0]); pr.stack(s.lay + nA, nB);
                for (int g = 0; g < G - 1; ++g) {
                    if (!p.cp_proj_rows) pr.stream(s.hd[0]);
                    pr.stack(s.lay + nA, nB, p.cp_proj_rows && p.cp_qkv0_rows); pr.stream(s.hd[2 + g]);
                }
                pr.stack(s.lay, nA); pr.stream(s.hd[1]);
            }
        }
        return;
    }

    // =============================== consumers ================================================================
    CState st;
    st.seq = 0; st.nstamp = 0; st.red_par = 0; st.kv_par = 0; st.res_par = 0; st.pp = 0; st.nsplit = 1; st.chunk = 128;
    st.gen = *reinterpret_cast<volatile unsigned int*>(p.state);
    LL_STAMP(ST_START);
    if (p.mode == LL_MODE_STACK) {
        const int pos = __ldcg(p.pos);
        for (int k4 = tid; k4 < (p.talker.hidden >> 2); k4 += LL_CTHREADS)
            reinterpret_cast<float4*>(s.resid + st.res_par * LL_MAXH)[k4] = __ldcg(reinterpret_cast<const float4*>(p.x_in) + k4);
        const StackOut so = stack_consume(st, p.talker, s.lay, pos, nullptr, 0u);
        if (p.head.w) {
            const uint32_t t_head = ++st.gen;
            gemv_phase(st, s.hd[1], in_norm(so.add, so.tag, p.talker.final_norm, p.talker.hidden, p.talker.eps, p.hidden_out), EPI_RAW,
                       p.x_head, p.logits_out, t_head);
        } else if (p.hidden_out) {
            final_norm_only(st, so.add, so.tag, p.talker.final_norm, p.hidden_out, p.talker.hidden, p.talker.eps);
        }
        LL_STAMP(ST_END);
    } else {
        const int step = __ldcg(p.step), pos_t = __ldcg(p.pos_talker);
        const int H = p.talker.hidden, E = p.emb_dim;
        const long long fo = (long long)step * G;             // offset of this frame in forced / own / codes
        const bool rec = (cta == 0 && tid == 0);
        const bool xon = tid < (H >> 2);
        // ---- code 0 from the talker logits of the previous launch
        int code;
        if (!p.talker_sp.do_sample) { st.red_par ^= 1; code = sample_greedy(p.logits, nullptr, 0u, p.talker_vocab, p.talker_sp, p.seen, step, st.red_par); }
        else code = sample_here(p.logits, nullptr, 0u, p.talker_vocab, p.talker_sp, p.seen, step, 0);
        if (rec && p.own_codes && step < p.max_frames) p.own_codes[fo] = code;
        if (p.forced) code = p.forced[fo];
        if (rec) {
            p.cur_codes[0] = code;
            if (step < p.max_frames) p.codes[fo] = code;
            if (p.done && code == p.talker_sp.eos_id) p.done[0] = 1;
        }
        const int code0 = code;
        LL_STAMP(ST_SAMPLE);
        // ---- code predictor: position 0 = projected talker hidden, then one pass per residual codebook.
        // The next talker input is accumulated on the way, in registers: emb0[c0] + emb1[c1] + ... in order
        cp_pass(st, p.hidden, nullptr, nullptr, 0, -1, step);
        LL_STAMP(ST_CP_PASS);
        float4 xn = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int g = 0; g < G - 1; ++g) {
            const float* row = (g == 0 ? p.codec_embedding : p.cp_embeddings[g - 1]) + (size_t)code * E;
            if (xon) {
                const float4 r4 = __ldg(reinterpret_cast<const float4*>(row) + tid);
                if (g == 0) xn = r4; else { xn.x += r4.x; xn.y += r4.y; xn.z += r4.z; xn.w += r4.w; }
            }
            code = cp_pass(st, row, p.cp_proj_rows ? s.prows[g] + (size_t)code * p.cp.hidden : nullptr,
                           p.cp_qkv0_rows ? s.qrows[g] + (size_t)code * ((p.cp.n_heads + 2 * p.cp.n_kv) * p.cp.head_dim) : nullptr, g + 1, g, step);
            LL_STAMP(ST_CP_PASS);
        }
        // ---- next talker input: running sum + last code's row, then the trailing text row
        {
            const float* last = p.cp_embeddings[G - 2] + (size_t)code * E;
            const int trow = step < p.n_trailing - 1 ? step : p.n_trailing - 1;
            const float* tr = p.trailing + (size_t)trow * H;
            if (xon) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(last) + tid);
                const float4 c = __ldcg(reinterpret_cast<const float4*>(tr) + tid);
                xn.x = (xn.x + b.x) + c.x; xn.y = (xn.y + b.y) + c.y; xn.z = (xn.z + b.z) + c.z; xn.w = (xn.w + b.w) + c.w;
                reinterpret_cast<float4*>(s.resid + st.res_par * LL_MAXH)[tid] = xn;
                if (cta == 0) reinterpret_cast<float4*>(p.x)[tid] = xn;
            }
        }
        // ---- talker decode step
        const StackOut so = stack_consume(st, p.talker, s.lay, pos_t, nullptr, 0u);
        const uint32_t t_head = ++st.gen;
        gemv_phase(st, s.hd[1], in_norm(so.add, so.tag, p.talker.final_norm, H, p.talker.eps, p.hidden), EPI_RAW, p.x_head, p.logits, t_head);
        LL_STAMP(ST_END);
        // state other CTAs read at the start of the launch is only updated here, after the last all-to-all exchange
        if (rec) {
            *p.pos_talker = pos_t + 1; *p.step = step + 1;
            if (p.seen) atomicOr(p.seen + (code0 >> 5), 1u << (code0 & 31));
        }
    }
    if (cta == 0 && tid == 0) *reinterpret_cast<volatile unsigned int*>(p.state) = st.gen;
}

// ---- host side ------------------------------------------------------------------------------------------------------------------
int num_sms();

static void fill_stack(LLStack& d, const q3t_stack& st) {
    d.layers = st.layers_dev; d.n_layers = st.n_layers; d.hidden = st.hidden; d.n_heads = st.n_heads; d.n_kv = st.n_kv_heads;
    d.head_dim = st.head_dim; d.inter = st.inter; d.eps = st.eps; d.final_norm = st.final_norm; d.inv_freq = st.inv_freq;
    d.kv_pool = (__nv_bfloat16*)st.kv_pool; d.kv_layer_stride = st.kv_layer_stride_bytes / 2; d.block_tbl = st.block_tbl;
}

// run-time tunables (profiling aids; the defaults are what the committed measurements use)
struct LLTune { int pf, fine, chunk, msplit; };
static const LLTune& ll_tune() {
    static LLTune t = {-1, 0, 64, LL_MAXSPLIT};
    if (t.pf < 0) {
        const char* e;
        t.pf = (e = getenv("Q3T_LL_PF")) ? atoi(e) : 0;
        t.fine = (e = getenv("Q3T_LL_FINE")) ? atoi(e) : 0;
        if ((e = getenv("Q3T_LL_CHUNK"))) { t.chunk = atoi(e); t.chunk = t.chunk < 16 ? 16 : (t.chunk + 15) / 16 * 16; }
        if ((e = getenv("Q3T_LL_MAXSPLIT"))) { t.msplit = atoi(e); t.msplit = t.msplit < 1 ? 1 : (t.msplit > LL_MAXSPLIT ? LL_MAXSPLIT : t.msplit); }
    }
    return t;
}

static int check_stack(const q3t_stack& st, int grid) {
    Q3T_REQUIRE(st.layers_dev != nullptr, "frame_ll: layers_dev missing");
    Q3T_REQUIRE(st.head_dim == 128, "frame_ll: head_dim must be 128");
    Q3T_REQUIRE(st.n_heads == st.n_kv_heads || st.n_heads == 2 * st.n_kv_heads, "frame_ll: H/Hkv must be 1 or 2");
    Q3T_REQUIRE(st.hidden % 256 == 0 && st.inter % 256 == 0, "frame_ll: dims % 256");
    Q3T_REQUIRE(st.hidden <= LL_MAXH && st.inter <= LL_MAXK && st.n_heads * st.head_dim <= LL_MAXK, "frame_ll: dims too large");
    Q3T_REQUIRE(st.n_kv_heads <= grid, "frame_ll: more kv heads than CTAs");
    const int qkv_n = (st.n_heads + 2 * st.n_kv_heads) * st.head_dim;
    const int n_max = 2 * st.inter > qkv_n ? 2 * st.inter : qkv_n;
    const long long t1 = ((long long)(n_max / 16) + grid - 1) / grid * (st.hidden / 256);
    const long long t2 = ((long long)(st.hidden / 16) + grid - 1) / grid * (st.inter / 256);
    Q3T_REQUIRE(t1 <= LL_MAXT && t2 <= LL_MAXT, "frame_ll: too many tiles per CTA");
    return 0;
}

// exchange-buffer layout inside the caller's workspace (64-bit words)
struct LLSizes { long long qkv, attn, qdim, hid, act; };
static LLSizes ll_sizes(const q3t_stack* t, const q3t_stack* c) {
    LLSizes z = {0, 0, 0, 0, 0};
    const q3t_stack* ss[2] = {t, c};
    for (int i = 0; i < 2; ++i) {
        if (!ss[i]) continue;
        const long long q = (long long)(ss[i]->n_heads + 2 * ss[i]->n_kv_heads) * ss[i]->head_dim;
        const long long a = (long long)ss[i]->n_heads * LL_MAXSPLIT * LL_REC;
        z.qkv = q > z.qkv ? q : z.qkv; z.attn = a > z.attn ? a : z.attn;
        const long long qd = (long long)ss[i]->n_heads * ss[i]->head_dim;
        z.qdim = qd > z.qdim ? qd : z.qdim;
        z.hid = ss[i]->hidden > z.hid ? ss[i]->hidden : z.hid; z.act = ss[i]->inter > z.act ? ss[i]->inter : z.act;
    }
    return z;
}
static long long up16(long long v) { return (v + 15) / 16 * 16; }
static long long ll_words(const q3t_stack* t, const q3t_stack* c, int head_max) {
    const LLSizes z = ll_sizes(t, c);
    return up16(z.qkv) + up16(z.attn) + up16(z.qdim) + 3 * up16(z.hid) + up16(z.act) + up16(head_max) + 64;
}
static void carve(LLParams& p, void* work, const q3t_stack* t, const q3t_stack* c) {
    const LLSizes z = ll_sizes(t, c);
    u64* w = (u64*)work;
    p.x_qkv = w; w += up16(z.qkv);
    p.x_attn = w; w += up16(z.attn);
    p.x_attnf = w; w += up16(z.qdim);
    p.x_o = w; w += up16(z.hid);
    p.x_down = w; w += up16(z.hid);
    p.x_proj = w; w += up16(z.hid);
    p.x_act = w; w += up16(z.act);
    p.x_head = w;
}

static int launch_ll(LLParams& p, cudaStream_t stream) {
    const LLTune& t = ll_tune();
    p.pf_dist = t.pf; p.fine = t.fine; p.att_chunk = t.chunk; p.att_maxsplit = t.msplit;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(frame_ll_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LL_SMEM_BYTES);
        attr_set = true;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(num_sms()); cfg.blockDim = dim3(LL_THREADS); cfg.dynamicSmemBytes = LL_SMEM_BYTES; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, frame_ll_kernel, p);
    Q3T_CHECK_LAUNCH("frame_ll");
    return 0;
}

int launch_stack_pass(const q3t_stack_pass_args* a, cudaStream_t stream) {
    const int grid = num_sms();
    if (int rc = check_stack(a->stack, grid)) return rc;
    Q3T_REQUIRE(a->stack.n_layers <= LL_MAXLAYERS, "stack_pass: too many layers");
    Q3T_REQUIRE(a->ll_work && a->ll_state, "stack_pass: workspace missing");
    if (a->head.w) Q3T_REQUIRE(a->head.N % 16 == 0 && a->head.K == a->stack.hidden, "stack_pass: head shape");
    LLParams p;
    memset(&p, 0, sizeof(p));
    p.mode = LL_MODE_STACK;
    fill_stack(p.talker, a->stack);
    carve(p, a->ll_work, &a->stack, nullptr);
    Q3T_REQUIRE(ll_words(&a->stack, nullptr, a->head.w ? a->head.N : 0) * 8 <= a->ll_work_bytes, "stack_pass: workspace too small");
    p.state = a->ll_state; p.timing = a->timing;
    p.pos = a->pos; p.x_in = a->x_in; p.hidden_out = a->hidden_out; p.logits_out = a->logits_out; p.head = a->head;
    p.n_groups = 1;
    return launch_ll(p, stream);
}

int launch_frame_ll(const q3t_frame_args* f, cudaStream_t stream) {
    const int grid = num_sms();
    if (int rc = check_stack(f->talker, grid)) return rc;
    if (int rc = check_stack(f->cp, grid)) return rc;
    Q3T_REQUIRE(f->B == 1, "frame_ll: batch 1 only");
    Q3T_REQUIRE(f->talker.n_layers + f->cp.n_layers <= LL_MAXLAYERS && f->talker.n_layers <= 64 && f->n_groups + 1 <= LL_MAXHEADS,
                "frame_ll: too many layers / code groups");
    Q3T_REQUIRE(f->ll_work && f->ll_state && f->cp_heads_dev && f->cp_embeddings_dev, "frame_ll: workspace / device tables missing");
    Q3T_REQUIRE(f->talker_vocab <= LL_SAMPLE_MAXV && f->cp_vocab <= LL_SAMPLE_MAXV && f->talker_vocab % 16 == 0 && f->cp_vocab % 16 == 0,
                "frame_ll: vocabulary size");
    Q3T_REQUIRE(f->cp_proj.K == f->talker.hidden && f->cp_proj.N == f->cp.hidden,
                "frame_ll: cp_proj shape (embedding width must equal the talker hidden size)");
    const int head_max = f->talker_vocab > f->cp_vocab ? f->talker_vocab : f->cp_vocab;
    Q3T_REQUIRE(ll_words(&f->talker, &f->cp, head_max) * 8 <= f->ll_work_bytes, "frame_ll: workspace too small");
    LLParams p;
    memset(&p, 0, sizeof(p));
    p.mode = LL_MODE_FRAME;
    fill_stack(p.talker, f->talker); fill_stack(p.cp, f->cp);
    carve(p, f->ll_work, &f->talker, &f->cp);
    p.state = f->ll_state; p.timing = f->ll_timing;
    p.codec_head = f->codec_head; p.cp_proj = f->cp_proj; p.cp_heads = f->cp_heads_dev;
    p.codec_embedding = f->codec_embedding; p.cp_embeddings = f->cp_embeddings_dev; p.cp_proj_rows = f->cp_proj_rows_dev;
    p.cp_qkv0_rows = f->cp_proj_rows_dev ? f->cp_qkv0_rows_dev : nullptr;
    p.emb_dim = f->talker.hidden; p.talker_vocab = f->talker_vocab; p.cp_vocab = f->cp_vocab; p.n_groups = f->n_groups;
    p.talker_sp = f->talker_sp; p.cp_sp = f->cp_sp;
    p.x = f->x; p.hidden = f->hidden; p.logits = f->logits; p.cp_logits = f->cp_logits; p.keep_cp_logits = f->keep_cp_logits;
    p.pos_talker = f->pos; p.step = f->step; p.cur_codes = f->cur_codes; p.codes = f->codes; p.own_codes = f->own_codes;
    p.max_frames = f->max_frames; p.seen = f->seen; p.done = f->done; p.trailing = f->trailing; p.n_trailing = f->n_trailing;
    p.forced = f->forced_codes;
    return launch_ll(p, stream);
}

}  // namespace q3t

extern "C" int q3t_stack_pass(const q3t_stack_pass_args* a, void* stream) {
    return q3t::launch_stack_pass(a, (cudaStream_t)stream);
}

extern "C" long long q3t_ll_work_bytes(const q3t_stack* talker, const q3t_stack* cp, int head_max) {
    return q3t::ll_words(talker, cp, head_max) * 8;
}
