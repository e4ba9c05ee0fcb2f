// Frame orchestration of the generation hot loop: every kernel of a talker step / a whole frame is
// enqueued from C++ on one stream (CUDA-graph capturable; no host round trip between the 16 sampled
// codes of a frame).  Replaces the per-frame body of mlx_audio's `Model.generate` loop
// (cousin driver transformers qwen3_omni_moe/modeling_qwen3_omni_moe.py:3243-3279).
#include <stdlib.h>
#include "common.cuh"
#include "../../include/q3tts_b200.h"

namespace q3t {

int launch_w8_gemv(const q3t_gemv_args* a, cudaStream_t stream);
int launch_w8_gemm(const q3t_gemm_args* a, cudaStream_t stream);
int launch_attn_decode(const q3t_attn_args* a, cudaStream_t stream);
int launch_attn_prefill(const q3t_attn_prefill_args* a, cudaStream_t stream);
int launch_sample(const q3t_sample_args* a, cudaStream_t stream);
int launch_stack_pass(const q3t_stack_pass_args* a, cudaStream_t stream);
int launch_frame_ll(const q3t_frame_args* f, cudaStream_t stream);

// ---- small kernels --------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rmsnorm_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                      float* __restrict__ y, int H, float eps) {
    __shared__ float red[32];
    pdl_wait();
    pdl_launch_dependents();
    const float* xr = x + (size_t)blockIdx.x * H;
    float* yr = y + (size_t)blockIdx.x * H;
    float ss = 0.f;
    for (int i = threadIdx.x; i < H; i += blockDim.x) { const float v = xr[i]; ss += v * v; }
    const float tot = block_sum(ss, red);
    const float rstd = rsqrtf(tot / (float)H + eps);
    for (int i = threadIdx.x; i < H; i += blockDim.x) yr[i] = w[i] * (xr[i] * rstd);
}

// step_per_row: one frame counter per slot (continuous batching); active (or NULL = all): only active slots advance
__global__ void advance_kernel(int* pos, int B, int* step, int step_per_row, const int* active) {
    pdl_wait();
    pdl_launch_dependents();
    const int i = threadIdx.x;
    const bool on = i < B && (!active || active[i]);
    if (on) pos[i] += 1;
    if (step && step_per_row) { if (on) step[i] += 1; }
    else if (step && i == 0) *step += 1;
}

// next talker input: emb_talker[c0] + sum_{g=1..G-1} emb_cp[g-1][c_g] summed sequentially in
// fp32, then + trailing_text[min(step, n-1)]; also archives the frame's codes.
__global__ void __launch_bounds__(256) next_input_kernel(const float* __restrict__ codec_emb,
                                                         const float* const* __restrict__ cp_emb, const int* cur_codes,
                                                         int G, int H, const float* __restrict__ trailing,
                                                         int n_trailing, const int* step_p, int step_per_row, float* x, int* codes,
                                                         int max_frames) {
    pdl_wait();
    pdl_launch_dependents();
    const int b = blockIdx.x, step = step_p[step_per_row ? b : 0];
    const int* cc = cur_codes + b * G;
    const int trow = step < n_trailing - 1 ? step : n_trailing - 1;
    const float* tr = trailing + ((size_t)b * n_trailing + trow) * H;
    for (int i = threadIdx.x; i < H; i += blockDim.x) {
        float acc = codec_emb[(size_t)cc[0] * H + i];
        for (int g = 1; g < G; ++g) acc = acc + cp_emb[g - 1][(size_t)cc[g] * H + i];
        x[(size_t)b * H + i] = acc + tr[i];
    }
    if (threadIdx.x < G && step < max_frames) codes[((size_t)b * max_frames + step) * G + threadIdx.x] = cc[threadIdx.x];
}

// input of code-predictor pass g for every sequence of a lock-step batch: row cur_codes[b][g] of the projected embedding
// table (and of the first layer's q|k|v table) - q3t_frame_args.cp_proj_rows_dev / cp_qkv0_rows_dev - instead of two GEMMs
__global__ void __launch_bounds__(256) cp_rows_gather_kernel(const float* const* __restrict__ proj_tabs,
                                                             const float* const* __restrict__ qkv_tabs, int g,
                                                             const int* cur_codes, int G, int Hc, int qkvd, float* xc, float* qkv) {
    pdl_wait();
    pdl_launch_dependents();
    const int b = blockIdx.x, code = cur_codes[(size_t)b * G + g];
    const float4* pr = reinterpret_cast<const float4*>(proj_tabs[g] + (size_t)code * Hc);
    float4* xo = reinterpret_cast<float4*>(xc + (size_t)b * Hc);
    for (int i = threadIdx.x; i < Hc / 4; i += blockDim.x) xo[i] = pr[i];
    if (qkv_tabs) {
        const float4* qr = reinterpret_cast<const float4*>(qkv_tabs[g] + (size_t)code * qkvd);
        float4* qo = reinterpret_cast<float4*>(qkv + (size_t)b * qkvd);
        for (int i = threadIdx.x; i < qkvd / 4; i += blockDim.x) qo[i] = qr[i];
    }
}

int launch_rmsnorm(const float* x, const float* w, float* y, int M, int H, float eps, cudaStream_t s) {
    launch_pdl(rmsnorm_kernel, dim3(M), dim3(256), 0, s, x, w, y, H, eps);
    Q3T_CHECK_LAUNCH("rmsnorm");
    return 0;
}

#define Q3T_TRY(expr) do { int rc__ = (expr); if (rc__) return rc__; } while (0)

// rows through one W8 matrix: B <= 2 -> exact-integer GEMV (two rows per launch); more rows -> tcgen05 GEMM (bf16 operands)
// scratch of the call in flight, set by the entry points below; thread_local: two engines driven from two host threads (ctypes
// releases the GIL) must not see each other's buffers
static thread_local void* g_xb = nullptr;    // bf16 scratch
static thread_local void* g_xb2 = nullptr;   // second bf16 scratch: attention output / SwiGLU activations handed from kernel to kernel as bf16
static thread_local float* g_ws = nullptr; static thread_local long long g_ws_floats = 0; static thread_local int* g_counters = nullptr;
static thread_local float* g_rowss = nullptr;   // row statistics of the deferred RMSNorm ([rows, 32] fp32); NULL = norms are prologue launches
static int gemv_rows(const q3t_w8& w, int B, int prologue, const float* x, long long xs, const float* norm_w, float eps,
                     const int* gidx, int gidx_stride, long long grow, int act, const float* resid, long long rs,
                     float* y, long long ys, cudaStream_t s) {
    if (B > 2 && g_xb && w.N % 128 == 0) {
        q3t_gemm_args a;
        memset(&a, 0, sizeof(a));
        a.w = w; a.M = B; a.prologue = prologue; a.x = x; a.x_stride = xs; a.norm_w = norm_w; a.eps = eps;
        a.gather_idx = gidx; a.gather_idx_stride = gidx_stride; a.gather_row_stride = grow; a.act = act;
        a.resid = resid; a.resid_stride = rs; a.y = y; a.y_stride = ys; a.xb = g_xb;
        a.splitk_ws = g_ws; a.splitk_ws_floats = g_ws_floats; a.splitk_counters = g_counters;
        return launch_w8_gemm(&a, s);
    }
    for (int b0 = 0; b0 < B; b0 += 2) {
        q3t_gemv_args a;
        memset(&a, 0, sizeof(a));
        a.w = w; a.M = (B - b0) >= 2 ? 2 : 1; a.prologue = prologue;
        a.x = x + (gidx ? 0 : b0 * xs); a.x_stride = gidx ? 0 : xs; a.norm_w = norm_w; a.eps = eps;
        a.gather_idx = gidx ? gidx + (long long)b0 * gidx_stride : nullptr; a.gather_idx_stride = gidx_stride;
        a.gather_row_stride = grow; a.act = act;
        a.resid = resid ? resid + b0 * rs : nullptr; a.resid_stride = rs;
        a.y = y + b0 * ys; a.y_stride = ys;
        Q3T_TRY(launch_w8_gemv(&a, s));
    }
    return 0;
}

// tcgen05 GEMM whose input rows are already bf16 (x_bf16) and / or whose output stays bf16 (y_bf16): no prologue launch.
// x_rowss: the input rows are weighted but not yet normalised (deferred RMSNorm, consumer half); y_norm_w: write fp32 y AND
// the weighted split rows + statistics for the next GEMM (producer half) - q3t_gemm_args in include/q3tts_b200.h.
static int gemm_bf16(const q3t_w8& w, int M, const void* x_bf16, int prologue, const float* x, long long xs, const float* norm_w,
                     float eps, int swiglu_out, const float* resid, long long rs, float* y, long long ys, void* y_bf16,
                     cudaStream_t s, const float* x_rowss = nullptr, int x_parts = 0, const float* y_norm_w = nullptr,
                     float* y_rowss = nullptr) {
    q3t_gemm_args a;
    memset(&a, 0, sizeof(a));
    a.w = w; a.M = M; a.prologue = prologue; a.x = x; a.x_stride = xs; a.norm_w = norm_w; a.eps = eps; a.swiglu_out = swiglu_out;
    a.resid = resid; a.resid_stride = rs; a.y = y; a.y_stride = ys; a.xb = g_xb; a.x_bf16 = x_bf16; a.y_bf16 = y_bf16;
    a.splitk_ws = g_ws; a.splitk_ws_floats = g_ws_floats; a.splitk_counters = g_counters;
    a.x_rowss = x_rowss; a.x_rowss_parts = x_parts; a.y_norm_w = y_norm_w; a.y_rowss = y_rowss;
    return launch_w8_gemm(&a, s);
}

// layers l .. of a dense Qwen3 stack on the batched path, from the O projection on: attention output (split rows in xb2) ->
// x updated in place.  With `rowss` the two RMSNorms that follow a residual GEMM are deferred: the O projection hands the
// gate/up GEMM split rows of x * post_norm in xb, the down projection hands the next layer's q|k|v GEMM split rows of
// x * input_norm - 5 launches per layer instead of 7.  Returns through *x_ready whether xb holds the next layer's input.
static int layer_tail_bf16(const q3t_stack& st, int l, int M, float* x, void* xb, void* xb2, float* rowss, bool* x_ready,
                           cudaStream_t s) {
    const q3t_layer& L = st.layers_host[l];
    const int hid = st.hidden, parts = hid / 128;
    const bool defer = rowss != nullptr && hid % 128 == 0 && parts <= 32;
    const float* next_norm = (defer && l + 1 < st.n_layers) ? st.layers_host[l + 1].input_norm : nullptr;
    Q3T_TRY(gemm_bf16(L.o, M, xb2, Q3T_PRO_RAW, nullptr, 0, nullptr, 0.f, 0, x, hid, x, hid, defer ? xb : nullptr, s,
                      nullptr, 0, defer ? L.post_norm : nullptr, defer ? rowss : nullptr));
    if (defer)
        Q3T_TRY(gemm_bf16(L.gate_up, M, xb, Q3T_PRO_RAW, nullptr, 0, nullptr, st.eps, 1, nullptr, 0, nullptr, 0, xb2, s, rowss, parts));
    else
        Q3T_TRY(gemm_bf16(L.gate_up, M, nullptr, Q3T_PRO_RMSNORM, x, hid, L.post_norm, st.eps, 1, nullptr, 0, nullptr, 0, xb2, s));
    Q3T_TRY(gemm_bf16(L.down, M, xb2, Q3T_PRO_RAW, nullptr, 0, nullptr, 0.f, 0, x, hid, x, hid, next_norm ? xb : nullptr, s,
                      nullptr, 0, next_norm, next_norm ? rowss : nullptr));
    *x_ready = next_norm != nullptr;
    return 0;
}

// one token through a dense Qwen3 stack; x [B, hidden] is updated in place (residual stream)
// skip_qkv0: f->qkv already holds the first layer's q|k|v rows (table lookup, see cp_rows_gather_kernel)
static int stack_forward(const q3t_frame_args* f, const q3t_stack& st, float* x, const int* pos, cudaStream_t s, bool skip_qkv0 = false) {
    const int B = f->B, hid = st.hidden, qd = st.n_heads * st.head_dim, kvd = st.n_kv_heads * st.head_dim;
    const int qkvd = qd + 2 * kvd;
    bool x_ready = false;       // g_xb holds split rows of x * input_norm of this layer + g_rowss its statistics (deferred RMSNorm)
    for (int l = 0; l < st.n_layers; ++l) {
        const q3t_layer& L = st.layers_host[l];
        if (x_ready)
            Q3T_TRY(gemm_bf16(L.qkv, B, g_xb, Q3T_PRO_RAW, nullptr, 0, nullptr, st.eps, 0, nullptr, 0, f->qkv, qkvd, nullptr, s,
                              g_rowss, hid / 128));
        else if (!(skip_qkv0 && l == 0))
            Q3T_TRY(gemv_rows(L.qkv, B, Q3T_PRO_RMSNORM, x, hid, L.input_norm, st.eps, nullptr, 0, 0, 0, nullptr, 0, f->qkv,
                              qkvd, s));
        x_ready = false;
        q3t_attn_args a;
        memset(&a, 0, sizeof(a));
        a.qkv = f->qkv; a.q_norm_w = L.q_norm; a.k_norm_w = L.k_norm; a.eps = st.eps; a.inv_freq = st.inv_freq;
        a.kv_pool = (char*)st.kv_pool + (long long)l * st.kv_layer_stride_bytes;
        a.block_tbl = st.block_tbl; a.max_pages = st.max_pages; a.pos = pos; a.out = f->attn; a.work = f->attn_work;
        a.counters = f->attn_counters; a.B = B; a.H = st.n_heads; a.Hkv = st.n_kv_heads; a.D = st.head_dim;
        a.nsplit = st.attn_nsplit;
        // batched path: the attention output and the SwiGLU activations go from kernel to kernel as bf16 rows (the GEMM
        // rounds its operands to bf16 anyway - bit-identical to the fp32 round trip through act_prep_kernel, two launches less)
        const bool chain = B > 2 && g_xb && g_xb2 && L.o.N % 128 == 0 && L.gate_up.N % 128 == 0 && L.down.N % 128 == 0;
        if (chain) a.out_bf16 = g_xb2;
        Q3T_TRY(launch_attn_decode(&a, s));
        if (chain) {
            const bool next_ok = l + 1 < st.n_layers && st.layers_host[l + 1].qkv.N % 128 == 0;
            Q3T_TRY(layer_tail_bf16(st, l, B, x, g_xb, g_xb2, next_ok || l + 1 == st.n_layers ? g_rowss : nullptr, &x_ready, s));
            continue;
        }
        Q3T_TRY(gemv_rows(L.o, B, Q3T_PRO_RAW, f->attn, qd, nullptr, 0.f, nullptr, 0, 0, 0, x, hid, x, hid, s));
        Q3T_TRY(gemv_rows(L.gate_up, B, Q3T_PRO_RMSNORM, x, hid, L.post_norm, st.eps, nullptr, 0, 0, 0, nullptr, 0, f->gu,
                          2 * st.inter, s));
        Q3T_TRY(gemv_rows(L.down, B, Q3T_PRO_SWIGLU, f->gu, 2 * st.inter, nullptr, 0.f, nullptr, 0, 0, 0, x, hid, x, hid,
                          s));
    }
    return 0;
}

// batch-1 persistent path: one cooperative launch for the whole talker stack (+ final norm + head)
static int mega_pass(const q3t_frame_args* f, const q3t_stack& st, const q3t_w8* head, const int* pos, const float* x_in,
                     float* hidden_out, float* logits_out, cudaStream_t s) {
    q3t_stack_pass_args a;
    memset(&a, 0, sizeof(a));
    a.stack = st;
    if (head) a.head = *head;
    a.pos = pos; a.x_in = x_in; a.hidden_out = hidden_out; a.logits_out = logits_out;
    a.ll_work = f->ll_work; a.ll_work_bytes = f->ll_work_bytes; a.ll_state = f->ll_state; a.timing = f->ll_timing;
    return launch_stack_pass(&a, s);
}

static inline bool mega_on(const q3t_frame_args* f) { return f->use_mega && f->B == 1 && f->ll_work && f->ll_state; }

static int talker_step(const q3t_frame_args* f, int want_logits, int bump_step, cudaStream_t s) {
    const q3t_stack& t = f->talker;
    Q3T_REQUIRE(f->B >= 1 && f->B <= 1024, "talker_step: batch out of range");
    if (mega_on(f)) {
        Q3T_TRY(mega_pass(f, t, want_logits ? &f->codec_head : nullptr, f->pos, f->x, f->hidden, f->logits, s));
    } else {
        Q3T_TRY(stack_forward(f, t, f->x, f->pos, s));
        Q3T_TRY(launch_rmsnorm(f->x, t.final_norm, f->hidden, f->B, t.hidden, t.eps, s));
        if (want_logits)
            Q3T_TRY(gemv_rows(f->codec_head, f->B, Q3T_PRO_RAW, f->hidden, t.hidden, nullptr, 0.f, nullptr, 0, 0, 0,
                              nullptr, 0, f->logits, f->talker_vocab, s));
    }
    launch_pdl(advance_kernel, dim3(1), dim3(1024), 0, s, f->pos, f->B, bump_step ? f->step : (int*)nullptr, f->step_per_row, f->active);
    Q3T_CHECK_LAUNCH("advance");
    return 0;
}

static int sample_into(const q3t_frame_args* f, const float* logits, int V, const q3t_sampling& sp, unsigned int* seen,
                       int g, int* done, cudaStream_t s) {
    q3t_sample_args a;
    memset(&a, 0, sizeof(a));
    const int G = f->n_groups;
    a.logits = logits; a.B = f->B; a.V = V; a.logits_stride = V; a.sp = sp; a.seen = seen; a.step = f->step;
    a.step_stride = f->step_per_row ? 1 : 0;
    a.rng_stream = g; a.out = f->cur_codes + g; a.out_stride = G;
    a.fo_stride = (long long)f->max_frames * G; a.fo_step_stride = G;
    a.forced = f->forced_codes ? f->forced_codes + g : nullptr;
    a.own = f->own_codes ? f->own_codes + g : nullptr;
    a.done = done;
    return launch_sample(&a, s);
}

static int frame(const q3t_frame_args* f, cudaStream_t s) {
    const q3t_stack& c = f->cp;
    const int B = f->B, G = f->n_groups, H = f->talker.hidden, Hc = c.hidden;
    // batch 1: the whole frame (sampler, 16 code-predictor passes, next input, talker step) is ONE persistent launch
    if (mega_on(f)) return launch_frame_ll(f, s);
    // code 0 from the talker logits
    Q3T_TRY(sample_into(f, f->logits, f->talker_vocab, f->talker_sp, f->seen, 0, f->done, s));
    // code predictor: position 0 = projected talker hidden, position 1 = projected embedding of code 0
    Q3T_TRY(gemv_rows(f->cp_proj, B, Q3T_PRO_RAW, f->hidden, H, nullptr, 0.f, nullptr, 0, 0, 0, nullptr, 0, f->xc, Hc, s));
    Q3T_TRY(stack_forward(f, c, f->xc, f->cp_pos, s));
    for (int g = 0; g < G - 1; ++g) {
        const float* table = g == 0 ? f->codec_embedding : f->cp_embeddings_host[g - 1];
        const bool tabs = f->cp_proj_rows_dev != nullptr, qtabs = tabs && f->cp_qkv0_rows_dev != nullptr;
        if (tabs) {
            launch_pdl(cp_rows_gather_kernel, dim3(B), dim3(256), 0, s, f->cp_proj_rows_dev, qtabs ? f->cp_qkv0_rows_dev : nullptr, g,
                       (const int*)f->cur_codes, G, Hc, (c.n_heads + 2 * c.n_kv_heads) * c.head_dim, f->xc, f->qkv);
            Q3T_CHECK_LAUNCH("cp_rows_gather");
        } else {
            Q3T_TRY(gemv_rows(f->cp_proj, B, Q3T_PRO_RAW, table, 0, nullptr, 0.f, f->cur_codes + g, G, H, 0, nullptr, 0, f->xc,
                              Hc, s));
        }
        float* lg = f->keep_cp_logits ? f->cp_logits + (size_t)g * B * f->cp_vocab : f->cp_logits;
        Q3T_TRY(stack_forward(f, c, f->xc, f->cp_pos + (size_t)(g + 1) * B, s, qtabs));
        Q3T_TRY(gemv_rows(f->cp_heads_host[g], B, Q3T_PRO_RMSNORM, f->xc, Hc, c.final_norm, c.eps, nullptr, 0, 0, 0,
                          nullptr, 0, lg, f->cp_vocab, s));
        Q3T_TRY(sample_into(f, lg, f->cp_vocab, f->cp_sp, nullptr, g + 1, nullptr, s));
    }
    launch_pdl(next_input_kernel, dim3(B), dim3(256), 0, s, f->codec_embedding, f->cp_embeddings_dev,
               (const int*)f->cur_codes, G, H, f->trailing, f->n_trailing, (const int*)f->step, f->step_per_row, f->x, f->codes,
               f->max_frames);
    Q3T_CHECK_LAUNCH("next_input");
    return talker_step(f, 1, 1, s);
}

// ---- talker prefill as GEMMs over all prompt tokens of all sequences (tcgen05 path) -------------------------------------
static int talker_prefill(const q3t_prefill_args* a, cudaStream_t s) {
    const q3t_frame_args* f = a->f;
    const q3t_stack& st = f->talker;
    const int M = a->M, hid = st.hidden, qd = st.n_heads * st.head_dim, kvd = st.n_kv_heads * st.head_dim, qkvd = qd + 2 * kvd;
    Q3T_REQUIRE(M >= 1 && a->x && a->pos && a->seq_of_row && a->qkv && a->attn && a->gu && a->xb, "talker_prefill: arguments");
    g_xb = a->xb; g_xb2 = nullptr; g_ws = nullptr; g_ws_floats = 0; g_counters = nullptr; g_rowss = nullptr;
    bool x_ready = false;
    for (int l = 0; l < st.n_layers; ++l) {
        const q3t_layer& L = st.layers_host[l];
        if (x_ready)
            Q3T_TRY(gemm_bf16(L.qkv, M, a->xb, Q3T_PRO_RAW, nullptr, 0, nullptr, st.eps, 0, nullptr, 0, a->qkv, qkvd, nullptr, s,
                              a->rowss, hid / 128));
        else
            Q3T_TRY(gemv_rows(L.qkv, M, Q3T_PRO_RMSNORM, a->x, hid, L.input_norm, st.eps, nullptr, 0, 0, 0, nullptr, 0, a->qkv, qkvd, s));
        x_ready = false;
        q3t_attn_args t;
        memset(&t, 0, sizeof(t));
        t.qkv = a->qkv; t.q_norm_w = L.q_norm; t.k_norm_w = L.k_norm; t.eps = st.eps; t.inv_freq = st.inv_freq;
        t.kv_pool = (char*)st.kv_pool + (long long)l * st.kv_layer_stride_bytes;
        t.block_tbl = st.block_tbl; t.max_pages = st.max_pages; t.pos = a->pos; t.out = a->attn; t.work = a->attn_work;
        t.counters = a->attn_counters; t.B = M; t.H = st.n_heads; t.Hkv = st.n_kv_heads; t.D = st.head_dim;
        t.seq_of_row = a->seq_of_row;
        const bool tc_attn = a->blocks && a->n_blocks > 0 && st.head_dim == 128 && st.n_heads == 2 * st.n_kv_heads;
        t.nsplit = 1; t.mode = 1;                    // pass 1: K/V rows of every prompt token into the cache
        if (!tc_attn) Q3T_TRY(launch_attn_decode(&t, s));
        if (tc_attn) {
            // both passes in csrc/attn_prefill.cu: K/V write, then attention on the tensor cores (32 rows of a sequence share
            // every K/V tile)
            q3t_attn_prefill_args u;
            memset(&u, 0, sizeof(u));
            u.qkv = a->qkv; u.q_norm_w = L.q_norm; u.eps = st.eps; u.inv_freq = st.inv_freq; u.kv_pool = t.kv_pool;
            u.block_tbl = st.block_tbl; u.max_pages = st.max_pages; u.pos = a->pos; u.seq_of_row = a->seq_of_row;
            u.blocks = a->blocks; u.n_blocks = a->n_blocks; u.out = a->attn; u.H = st.n_heads; u.Hkv = st.n_kv_heads; u.D = st.head_dim;
            u.k_norm_w = L.k_norm; u.M = M;
            const bool chain = a->xb2 && M > 2 && L.o.N % 128 == 0 && L.gate_up.N % 128 == 0 && L.down.N % 128 == 0;
            if (chain) { u.out = nullptr; u.out_bf16 = a->xb2; }
            Q3T_TRY(launch_attn_prefill(&u, s));
            if (chain) {          // attention output and SwiGLU activations go on as bf16 rows (as in the decode frames)
                const bool next_ok = l + 1 < st.n_layers && st.layers_host[l + 1].qkv.N % 128 == 0;
                Q3T_TRY(layer_tail_bf16(st, l, M, a->x, a->xb, a->xb2, next_ok || l + 1 == st.n_layers ? a->rowss : nullptr, &x_ready, s));
                continue;
            }
        } else {
            t.mode = 2;                              // pass 2: causal attention of every row over its prefix
            Q3T_TRY(launch_attn_decode(&t, s));
        }
        Q3T_TRY(gemv_rows(L.o, M, Q3T_PRO_RAW, a->attn, qd, nullptr, 0.f, nullptr, 0, 0, 0, a->x, hid, a->x, hid, s));
        Q3T_TRY(gemv_rows(L.gate_up, M, Q3T_PRO_RMSNORM, a->x, hid, L.post_norm, st.eps, nullptr, 0, 0, 0, nullptr, 0, a->gu,
                          2 * st.inter, s));
        Q3T_TRY(gemv_rows(L.down, M, Q3T_PRO_SWIGLU, a->gu, 2 * st.inter, nullptr, 0.f, nullptr, 0, 0, 0, a->x, hid, a->x, hid, s));
    }
    return 0;
}

static int talker_tail(const q3t_frame_args* f, cudaStream_t s) {
    const q3t_stack& t = f->talker;
    g_xb = f->gemm_xb; g_xb2 = nullptr; g_ws = f->gemm_ws; g_ws_floats = f->gemm_ws_floats; g_counters = f->gemm_counters; g_rowss = nullptr;
    Q3T_TRY(launch_rmsnorm(f->x, t.final_norm, f->hidden, f->B, t.hidden, t.eps, s));
    return gemv_rows(f->codec_head, f->B, Q3T_PRO_RAW, f->hidden, t.hidden, nullptr, 0.f, nullptr, 0, 0, 0, nullptr, 0, f->logits,
                     f->talker_vocab, s);
}

}  // namespace q3t

extern "C" int q3t_rmsnorm(const float* x, const float* w, float* y, int M, int H, float eps, void* stream) {
    return q3t::launch_rmsnorm(x, w, y, M, H, eps, (cudaStream_t)stream);
}
extern "C" int q3t_talker_step(const q3t_frame_args* f, int want_logits, void* stream) {
    q3t::g_xb = f->gemm_xb; q3t::g_xb2 = getenv("Q3T_NO_BF16_CHAIN") ? nullptr : f->gemm_xb2; q3t::g_ws = f->gemm_ws; q3t::g_ws_floats = f->gemm_ws_floats; q3t::g_counters = f->gemm_counters; q3t::g_rowss = q3t::g_xb2 ? f->gemm_rowss : nullptr;
    return q3t::talker_step(f, want_logits, 0, (cudaStream_t)stream);
}
extern "C" int q3t_frame(const q3t_frame_args* f, void* stream) {
    q3t::g_xb = f->gemm_xb; q3t::g_xb2 = getenv("Q3T_NO_BF16_CHAIN") ? nullptr : f->gemm_xb2; q3t::g_ws = f->gemm_ws; q3t::g_ws_floats = f->gemm_ws_floats; q3t::g_counters = f->gemm_counters; q3t::g_rowss = q3t::g_xb2 ? f->gemm_rowss : nullptr;
    return q3t::frame(f, (cudaStream_t)stream);
}
extern "C" int q3t_talker_prefill(const q3t_prefill_args* a, void* stream) { return q3t::talker_prefill(a, (cudaStream_t)stream); }
extern "C" int q3t_talker_tail(const q3t_frame_args* f, void* stream) { return q3t::talker_tail(f, (cudaStream_t)stream); }
