/*
 * q3tts_b200.h -- C ABI of the B200-native Qwen3-TTS generation hot path.
 *
 * This is the drop-in boundary BELOW the two Python callables the reference imports
 * (`mlx_audio.tts.utils.load_model`, reference src/qwen3_tts/io.py:111-112, and
 * `mlx_audio.tts.generate.generate_audio`, sessions/custom.py:163-170, design.py:76-81,
 * clone.py:218-224).  In the reference stack these entry points are MLX graph ops that end in
 * Metal kernels (un-vendored: mlx==0.30.3 / mlx-audio==0.3.1, pyproject.toml:38-41); each
 * function below names the MLX op it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`;
 *   - the caller (PyTorch) allocates every buffer, including workspaces; no function allocates,
 *     frees or synchronises; everything is enqueued on `stream` (a cudaStream_t passed as void*)
 *     and is CUDA-graph capturable;
 *   - return value 0 = ok, non-zero = error (message via q3t_last_error()); no C++ exception
 *     crosses the boundary;
 *   - activations are fp32 row-major, K/V pages are bf16, weights are "W8 tiles" (below).
 *
 * W8 tile format (affine 8-bit, group 64; MLX `quantized_matmul` weights re-laid for HBM streaming)
 *   matrix [N, K], N % 16 == 0, K % 256 == 0, stored as (N/16)*(K/256) tiles of 4352 bytes ordered
 *   [row_tile][k_chunk]; a tile = 4096 B of uint8 codes in mma-fragment order + 16 rows x
 *   {4 bf16 scales, 4 bf16 biases}.  See qwen3_tts_b200/weights.py:pack_w8 for the exact permutation.
 */
#ifndef Q3TTS_B200_H
#define Q3TTS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define Q3T_ABI_VERSION 2   /* 2: round 2 - fields appended to q3t_gemm_args, q3t_tapgemm_args, q3t_frame_args, q3t_prefill_args, q3t_sample_args; new entry points */
#define Q3T_TILE_BYTES 4352
#define Q3T_KV_PAGE 16

int q3t_abi_version(void);
const char* q3t_last_error(void);
/* number of kernel launches enqueued by this library since load (bench.py "gpu_launches") */
unsigned long long q3t_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * W8 GEMV  (replaces mx.quantized_matmul, qmv path, M = 1..2 rows per launch)
 *   y[m, :] = epilogue( W8 . prologue(x[m, :]) )
 * prologue: RAW            x is [M, K]
 *           RMSNORM        x is [M, K]; x * rsqrt(mean(x^2)+eps) * norm_w      (mx.fast.rms_norm)
 *           SWIGLU         x is [M, 2K] with gate/up interleaved in blocks of 8 (x[16j+r] = gate[8j+r],
 *                          x[16j+8+r] = up[8j+r], r < 8: the row order of the fused gate_up matrix); silu(gate) * up
 * gather  : if `gather_idx` != NULL, row m of x is x + gather_idx[m*gather_idx_stride] * gather_row_stride
 * epilogue: (+ lin_bias[N]) -> act (0 none, 1 SiLU) -> (+ resid[m, :])   ; y may alias resid
 * ------------------------------------------------------------------------------------------- */
enum { Q3T_PRO_RAW = 0, Q3T_PRO_RMSNORM = 1, Q3T_PRO_SWIGLU = 2 };

typedef struct {
    const void* w;            /* W8 tiles */
    int N, K;
    const float* lin_bias;    /* [N] or NULL */
} q3t_w8;

typedef struct {
    q3t_w8 w;
    int M;                    /* 1 or 2 */
    int prologue;
    const float* x;  long long x_stride;        /* floats between rows */
    const float* norm_w;  float eps;            /* RMSNORM only */
    const int* gather_idx;  int gather_idx_stride;  long long gather_row_stride;
    int act;
    const float* resid;  long long resid_stride;
    float* y;  long long y_stride;
} q3t_gemv_args;

int q3t_w8_gemv(const q3t_gemv_args* a, void* stream);
/* The same contraction over n_rows rows (a->M is ignored): two rows per launch, enqueued from C.  Used to project whole
 * embedding tables at load (q3t_frame_args.cp_proj_rows_dev) without a host round trip per pair of rows. */
int q3t_w8_gemv_rows(const q3t_gemv_args* a, int n_rows, void* stream);

/* ---------------------------------------------------------------------------------------------
 * W8 GEMM on the tcgen05 tensor cores (replaces mx.quantized_matmul, qmm path: batched decode, prefill)
 *   y[m, :] = epilogue( W8 . prologue(x[m, :]) ),  m < M (any M >= 1), N % 128 == 0, K % 256 == 0
 * Same prologues / gather / epilogue as q3t_w8_gemv; operands are rounded to bf16, accumulation is fp32 in TMEM.
 * swiglu_out != 0: W is a fused gate/up matrix with rows interleaved in blocks of 8; the kernel writes
 *   y[m, 8j + r] = silu(gate[8j + r]) * up[8j + r]   (N/2 values per row) and ignores act / resid.
 * xb: bf16 scratch of at least M*K elements (the prologue's output, consumed by the GEMM).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    q3t_w8 w;
    int M;
    int prologue;
    const float* x;  long long x_stride;
    const float* norm_w;  float eps;
    const int* gather_idx;  int gather_idx_stride;  long long gather_row_stride;
    int act;  int swiglu_out;
    const float* resid;  long long resid_stride;
    float* y;  long long y_stride;
    void* xb;                                  /* scratch for the prepared activations: split bf16 rows [M, 2K] (4*M*K bytes) */
    /* optional split-K workspace (decode-sized M): partial sums [splits][M][N] fp32 + zero-initialised counters [1024] */
    float* splitk_ws;  long long splitk_ws_floats;  int* splitk_counters;
    /* optional bf16 chaining between kernels (saves the fp32 round trip and the prologue launch).  Every operand of the
     * tensor-core GEMM is carried as hi + lo, both bf16 (hi = rn(v), lo = rn(v - hi)); a "split row" of width K is
     * [hi(K) | lo(K)] = 2K bf16, contiguous:
     *   x_bf16: the input rows are ALREADY split rows [M, 2K] (prologue must be Q3T_PRO_RAW, x is ignored);
     *   y_bf16: the epilogue writes split rows [M, 2N] (or [M, 2(N/2)] with swiglu_out) INSTEAD of fp32 y. */
    const void* x_bf16;  void* y_bf16;
    /* optional DEFERRED RMSNorm between two GEMMs of a residual stream (no prologue launch, no fp32 re-read of the row):
     *   producer (y_norm_w != NULL; needs y, y_bf16, y_rowss; not with swiglu_out): besides the fp32 rows y the epilogue writes
     *     split rows of y[m, n] * y_norm_w[n] - weighted, NOT yet normalised - to y_bf16, and the sum of squares of y[m, :] over
     *     each block of 128 features to y_rowss[m * (N/128) + n/128];
     *   consumer (x_rowss != NULL; needs x_bf16): the contraction of row m is multiplied by
     *     rsqrt(sum_j x_rowss[m * x_rowss_parts + j] / K + eps) before bias / activation / SwiGLU - the normalisation commutes
     *     with the contraction, so the pair computes W . RMSNorm(y) with the row statistics summed in a fixed order. */
    const float* y_norm_w;  float* y_rowss;
    const float* x_rowss;  int x_rowss_parts;
} q3t_gemm_args;

int q3t_w8_gemm(const q3t_gemm_args* a, void* stream);

/* RMSNorm rows: y[m,:] = x[m,:] * rsqrt(mean(x^2)+eps) * w            (mx.fast.rms_norm) */
int q3t_rmsnorm(const float* x, const float* w, float* y, int M, int H, float eps, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Paged-KV GQA decode attention with fused per-head q/k RMSNorm + RoPE + KV-page write
 *   (replaces mx.fast.rms_norm x2, mx.fast.rope x2, cache update, mx.fast.scaled_dot_product_attention)
 * qkv       [B, (H + 2*Hkv) * D] fp32: q heads, then k heads, then v heads (fused QKV GEMV output)
 * kv_pool   bf16 pages of ONE layer: [n_pages][2 (k,v)][Hkv][Q3T_KV_PAGE][D]
 * block_tbl [B, max_pages] int32 page ids;  pos [B] int32 = index of the NEW token (ctx = pos+1)
 * inv_freq  [D/2] fp32 RoPE inverse frequencies (host-computed exactly as the oracle does)
 * out       [B, H*D] fp32
 * work      fp32 workspace >= B*Hkv*nsplit*(H/Hkv)*(D+2); counters: int32 [B*Hkv], zero-initialised once
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    const float* qkv;
    const float* q_norm_w; const float* k_norm_w; float eps;
    const float* inv_freq;
    void* kv_pool;
    const int* block_tbl; int max_pages;
    const int* pos;
    float* out;
    float* work; int* counters;
    int B, H, Hkv, D, nsplit;
    int mode;                 /* 0 = fused decode step; 1 = only write K/V of every row (prefill pass 1, use nsplit = 1);
                                 2 = attention only, K/V of the row itself already in the cache (prefill pass 2) */
    const int* seq_of_row;    /* optional [B]: block-table row of launch row b (prefill rows of one sequence share pages) */
    void* out_bf16;           /* optional split bf16 rows [B, 2*H*D] (hi | lo): written INSTEAD of `out` (feeds q3t_w8_gemm.x_bf16 directly) */
} q3t_attn_args;

int q3t_attn_decode(const q3t_attn_args* a, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Causal GQA attention of the PROMPT rows on the tensor cores (prefill pass 2; replaces q RMSNorm + RoPE +
 * mx.fast.scaled_dot_product_attention over the prompt).  K/V of every row are already in the paged cache
 * (q3t_attn_decode mode 1).  Rows are ragged: row m is position pos[m] of sequence seq_of_row[m]; `blocks` [n_blocks, 2]
 * = (first row, row count <= 32) lists runs of consecutive rows of ONE sequence (one CTA per block and kv head).
 * Built for D = 128 and H = 2 * Hkv.  Writes out [M, H*D] fp32, or out_bf16 as split rows
 * [M, 2*H*D] = hi(H*D) | lo(H*D) (see q3t_gemm_args.x_bf16) when given.
 * k_norm_w != NULL (with M = number of rows): pass 1 runs first in the same call - K (RMSNorm + RoPE) and V of every row
 * are written to the cache, bit-identical to q3t_attn_decode mode 1.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    const float* qkv;  const float* q_norm_w;  float eps;  const float* inv_freq;
    void* kv_pool;  const int* block_tbl;  int max_pages;
    const int* pos;  const int* seq_of_row;  const int* blocks;  int n_blocks;
    float* out;  void* out_bf16;
    int H, Hkv, D;
    const float* k_norm_w;  int M;
} q3t_attn_prefill_args;

int q3t_attn_prefill(const q3t_attn_prefill_args* a, void* stream);

/* ---------------------------------------------------------------------------------------------
 * On-device sampler (replaces mx.argmax / mx.random.categorical + host-side logits processors)
 * Order of operations = HF generation: repetition penalty over the set of
 * previously generated ids, min_new_tokens EOS mask, suppress range (except eos), temperature,
 * top-k (ties kept), top-p, then argmax (lowest index wins) or inverse-CDF categorical draw.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int do_sample; float temperature; int top_k; float top_p; float repetition_penalty;
    int min_new_tokens; int suppress_lo, suppress_hi, eos_id;
    unsigned long long seed;
} q3t_sampling;

typedef struct {
    const float* logits; int B, V; long long logits_stride;
    q3t_sampling sp;
    unsigned int* seen;       /* [B, ceil(V/32)] bitmask of generated ids (NULL: no penalty, no update) */
    const int* step;          /* frames generated so far (min_new_tokens, RNG counter): device scalar, or [B] with step_stride = 1 */
    int rng_stream;           /* distinguishes the 16 draws of a frame */
    const float* uniforms;    /* optional [B] externally supplied uniforms (tests) */
    int* out; long long out_stride;             /* out[b*out_stride] = chosen id (fixed address) */
    long long fo_stride, fo_step_stride;        /* addressing of forced/own: [b*fo_stride + step*fo_step_stride] */
    const int* forced;        /* optional teacher forcing: overrides the choice */
    int* own;                 /* optional: what the sampler itself picked (before forcing) */
    int* done;                /* optional [B]: set to 1 when the chosen id == eos_id */
    int step_stride;          /* 0: `step` is one scalar shared by the lock-step batch; 1: one frame counter per row */
} q3t_sample_args;

int q3t_sample(const q3t_sample_args* a, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Whole-frame enqueue: talker decode step -> sample code0 -> 15 code-predictor passes -> next input
 *   (replaces the per-frame body of mlx_audio's Model.generate loop; SURVEY 3.1 "HOT LOOP")
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    const float* input_norm; q3t_w8 qkv;
    const float* q_norm; const float* k_norm;
    q3t_w8 o;
    const float* post_norm; q3t_w8 gate_up; q3t_w8 down;
} q3t_layer;

typedef struct {
    int hidden, n_layers, n_heads, n_kv_heads, head_dim, inter;
    float eps;
    const q3t_layer* layers_host;   /* HOST array [n_layers] */
    const q3t_layer* layers_dev;    /* the same array in DEVICE memory (persistent stack-pass kernel) */
    const float* final_norm;
    const float* inv_freq;          /* [head_dim/2] */
    void* kv_pool; long long kv_layer_stride_bytes;   /* bf16 pages, per-layer stride */
    const int* block_tbl; int max_pages;
    int attn_nsplit;
} q3t_stack;

/* ---------------------------------------------------------------------------------------------
 * Persistent stack pass (batch 1): ONE cooperative launch = every layer of a dense Qwen3 stack for one token
 * + final RMSNorm + optional head GEMV (csrc/frame_ll.cu).  Weights stream through a TMA/mbarrier shared-memory
 * ring; activations cross CTAs as 64-bit {value, phase tag} words (no grid barrier).  Replaces 5*n_layers+2
 * launches of the kernels above.  In matrices handed to this kernel the fused gate/up rows are interleaved in
 * blocks of 8 (rows 16j..16j+7 = gate rows 8j.., rows 16j+8..16j+15 = the matching up rows).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    q3t_stack stack;
    q3t_w8 head;            /* head.w == NULL: no head GEMV */
    const int* pos;         /* [1] position of the token being fed */
    const float* x_in;      /* [hidden] */
    float* hidden_out;      /* [hidden] post-final-norm, or NULL */
    float* logits_out;      /* [head.N] */
    void* ll_work;          /* exchange buffers, q3t_ll_work_bytes() bytes, zero-initialised once */
    long long ll_work_bytes;
    unsigned int* ll_state; /* [2] zero-initialised once: phase-tag counter (persists across launches), error code */
    unsigned long long* timing;  /* optional [grid][2048] clock64 stamps; written only by the -DLL_STAMPS profiling build (libq3tts_b200_prof.so), or NULL */
} q3t_stack_pass_args;

int q3t_stack_pass(const q3t_stack_pass_args* a, void* stream);
/* workspace size for the persistent kernels; `cp` may be NULL (stack pass only) */
long long q3t_ll_work_bytes(const q3t_stack* talker, const q3t_stack* cp, int head_max);

typedef struct {
    int B;
    /* talker */
    q3t_stack talker;
    q3t_w8 codec_head;              /* [V, H] */
    int talker_vocab;
    const float* codec_embedding;   /* [V, H] fp32 */
    q3t_sampling talker_sp;
    /* code predictor */
    q3t_stack cp;
    q3t_w8 cp_proj;                 /* [Hc, H] + bias */
    const float* const* cp_embeddings_host;   /* HOST array [G-1] of device ptrs [Vc, H] */
    const float* const* cp_embeddings_dev;    /* the same table in device memory */
    const q3t_w8* cp_heads_host;    /* HOST array [G-1] */
    int cp_vocab, n_groups;
    q3t_sampling cp_sp;
    /* per-utterance state (device) */
    float* x;              /* [B, H]  talker residual stream: holds the next input on entry */
    float* hidden;         /* [B, H]  post-final-norm talker hidden */
    float* logits;         /* [B, V]  */
    float* cp_logits;      /* [B, Vc] (per pass; also [G-1, B, Vc] when keep_cp_logits) */
    int keep_cp_logits;
    float* xc;             /* [B, Hc] code-predictor residual stream */
    float* qkv;            /* [B, max(q+2kv dims)] */
    float* attn;           /* [B, max(H*D)] */
    float* gu;             /* [B, 2*max(inter)] */
    float* attn_work; int* attn_counters;
    int* pos;              /* [B] talker position of the token being fed */
    int* cp_pos;           /* [G, B] constant table: cp_pos[g][b] = g */
    int* step;             /* device scalar: frame index */
    int* cur_codes;        /* [B, G] codes of the frame being generated */
    int* codes;            /* [B, max_frames, G] int32 */
    int* own_codes;        /* optional [B, max_frames, G]: un-forced choices (parity tests) */
    int max_frames;
    unsigned int* seen;    /* [B, ceil(V/32)] */
    int* done;             /* [B] */
    const float* trailing; /* [B, n_trailing, H]; row min(step, n_trailing-1) is added (last row = tts_pad) */
    int n_trailing;
    const int* forced_codes;  /* optional [B, max_frames, G]: teacher forcing (parity tests) */
    void* gemm_xb;         /* bf16 scratch of split rows [B, 2 * max K] for the tcgen05 GEMM (used when B > 2), or NULL */
    float* gemm_ws; long long gemm_ws_floats; int* gemm_counters;   /* split-K workspace of the GEMM (optional) */
    /* persistent-kernel path (used when use_mega != 0 and B == 1): the whole frame is ONE launch (csrc/frame_ll.cu) */
    int use_mega;
    const q3t_w8* cp_heads_dev;     /* cp_heads_host in DEVICE memory */
    void* ll_work; long long ll_work_bytes; unsigned int* ll_state;
    unsigned long long* ll_timing;  /* optional profiling stamps (profiling build only), or NULL */
    void* gemm_xb2;        /* optional second bf16 scratch [B, 2 * max K]: attention output and SwiGLU activations stay (split) bf16 */
    /* optional (persistent-kernel path): DEVICE array [G-1] of device tables; table g = cp_proj applied to every row of the
     * embedding table that feeds code-predictor pass g (g = 0: codec_embedding [V, Hc]; g >= 1: cp_embeddings[g-1] [Vc, Hc]),
     * fp32.  The input of a pass is a table row - a function of one sampled code - so its projection is a lookup; with the
     * tables the kernel skips one contraction phase per pass.  NULL = project in the kernel. */
    const float* const* cp_proj_rows_dev;
    /* optional, needs cp_proj_rows_dev: DEVICE array [G-1] of tables [V_g, (H+2Hkv)*D] fp32 = the first code-predictor layer's
     * fused q|k|v projection of RMSNorm(projected row) - equally a function of one code.  With it the first layer of those
     * passes starts at the attention (one more contraction phase less per pass).  NULL = compute in the kernel. */
    const float* const* cp_qkv0_rows_dev;
    /* continuous batching (batched path, B > 2; qwen3_tts_b200/serving.py): rows of the batch are SLOTS that requests enter
     * and leave at different frames.  step_per_row != 0: `step` is [B], one frame counter per slot (sampler, trailing-text row
     * and the codes archive index by it).  active [B] (or NULL = all): only active slots advance their position and frame
     * counter; an idle slot keeps pos = 0 over a scratch page and its outputs are ignored. */
    int step_per_row;
    const int* active;
    /* optional [B, 32] fp32: row statistics of the deferred RMSNorm between the GEMMs of the batched path (q3t_gemm_args.y_rowss);
     * NULL = every norm is a prologue launch (act_prep_kernel) */
    float* gemm_rowss;
} q3t_frame_args;

/* Talker prefill as GEMMs: M rows = the prompt tokens of all sequences, concatenated (no padding).
 * x [M, H] is the residual stream (in: prompt embeddings; out: last-layer output, pre final norm); pos[m] = position of
 * row m in its sequence, seq_of_row[m] = its sequence (block-table row).  Workspaces: qkv [M, (H+2Hkv)*D], attn [M, H*D],
 * gu [M, 2*inter] fp32; xb bf16 [M, max(hidden, inter)].  Fills the KV cache of every layer. */
typedef struct {
    const q3t_frame_args* f;
    int M;
    float* x; const int* pos; const int* seq_of_row;
    float* qkv; float* attn; float* gu; void* xb;
    float* attn_work; int* attn_counters;    /* >= M*Hkv*(H/Hkv)*(D+2) floats, M*Hkv ints (zeroed) */
    const int* blocks; int n_blocks;         /* optional row blocks for q3t_attn_prefill (see there); 0 = per-row decode kernel */
    void* xb2;                               /* optional second bf16 scratch of split rows [M, 2*max(H*D, inter)]: attention output and SwiGLU
                                                activations stay bf16 between kernels (needs `blocks`) */
    float* rowss;                            /* optional [M, 32] fp32: deferred RMSNorm statistics (needs xb2; see q3t_gemm_args.y_rowss) */
} q3t_prefill_args;
int q3t_talker_prefill(const q3t_prefill_args* a, void* stream);
/* final RMSNorm of `x` [B, H] -> `hidden`, codec head -> `logits` (the tail of a talker step) */
int q3t_talker_tail(const q3t_frame_args* f, void* stream);

/* one talker forward for the token currently in `x` (prefill token or decode step), logits optional */
int q3t_talker_step(const q3t_frame_args* f, int want_logits, void* stream);
/* everything after the talker logits of a frame + the next talker step */
int q3t_frame(const q3t_frame_args* f, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Speech-tokenizer decoder operators (replace mx.take/add, mx.conv1d, mx.conv_transpose1d, ...)
 * Activations are time-major fp32 [B, T, C].
 * ------------------------------------------------------------------------------------------- */
/* bit-exact RVQ gather + left-to-right fp32 sum: codes [B, G, T] int32, tables HOST array of G device
 * ptrs [size, dim]; groups [g_lo, g_hi) summed into out [B, T, dim].  A code outside [0, codebook_size) (EOS / control
 * ids in the frames a finished sequence of a lock-step batch keeps producing) contributes a zero vector. */
int q3t_rvq_gather_sum(const int* codes, const float* const* tables_host, int B, int G, int T, int g_lo, int g_hi,
                       int dim, int codebook_size, float* out, void* stream);

enum { Q3T_ACT_NONE = 0, Q3T_ACT_SILU = 1, Q3T_ACT_GELU = 2, Q3T_ACT_SNAKE = 3, Q3T_ACT_SWIGLU_PAIR = 4,
       /* reference-clip encoders (speech-tokenizer encoder: ELU; ECAPA speaker encoder: ReLU, sigmoid, tanh) */
       Q3T_ACT_ELU = 5, Q3T_ACT_RELU = 6, Q3T_ACT_SIGMOID = 7, Q3T_ACT_TANH = 8 };

/* Generic causal tap-GEMM:  out[b, t*up + p, co] = epi( bias[co] + sum_{j<taps} sum_ci A[b, t + shift_j, ci] * W[j][p*Cout + co][ci] )
 *   conv1d (k taps, dilation d): up = 1, shift_j = -(k-1-j)*d          (rows before 0 read as zero)
 *   conv_transpose1d (k = 2*up): taps = 2 with shifts given by `shift`; N = up*Cout
 *   linear: taps = 1, shift 0
 * epilogue: v = acc + bias; v = v * scale[co] (optional); v += resid (optional) ; out_raw = v (optional);
 *           out_act = act(v) (optional; SNAKE uses act_a = exp(alpha), act_b = 1/(exp(beta)+1e-9) per channel) */
typedef struct {
    const float* A; int B, T_in, Cin;          /* [B, T_in, Cin] */
    const float* W;                            /* [taps][N][Cin], N = up*Cout */
    const float* bias;                         /* [Cout] or NULL */
    int taps; int shift[8]; int up; int Cout;
    int T_out_rows;                            /* GEMM rows per batch item (t range) */
    const float* scale;                        /* [Cout] or NULL */
    const float* resid;                        /* [B, T_out_rows*up, Cout] or NULL */
    float* out_raw;                            /* or NULL */
    float* out_act; int act; const float* act_a; const float* act_b;
    int force_fp32;                            /* != 0: FP32-pipe kernel even where the tcgen05 TF32 path is eligible (the
                                                  reference-clip encoders: a nearest-codebook search follows) */
    /* 16-bit operands between the layers of the vocoder (tcgen05 path only; ask q3t_tapgemm_tc_eligible first):
     *   a_f16   != 0: A and W point to IEEE fp16 data of the same shapes (kind::f16: 11 significant bits - one more than the
     *                 TF32 read of fp32 data - at half the bytes through HBM, L2 and shared memory; fp32 accumulation);
     *   act_f16 != 0: out_act is written as fp16 (saturating at +-65504) for such a consumer; out_raw / resid stay fp32. */
    int a_f16;  int act_f16;
} q3t_tapgemm_args;

int q3t_tapgemm(const q3t_tapgemm_args* a, void* stream);
/* 1 when q3t_tapgemm would run this call on the tcgen05 kernel (shape rule of csrc/tapgemm_tc.cu), else 0 */
int q3t_tapgemm_tc_eligible(const q3t_tapgemm_args* a);
/* which kernel served the q3t_tapgemm calls so far: out3[0] tcgen05 TF32 tap-GEMM, out3[1] FP32-pipe kernel although
 * Cin % 32 == 0 (fewer than 64 GEMM rows, N not a multiple of 16), out3[2] FP32-pipe kernel because Cin % 32 != 0.
 * reset != 0 zeroes the counters after reading.  (Test / evidence aid: parity at the BASELINE shapes asserts out3[1] == 0.) */
void q3t_tapgemm_stats(unsigned long long* out3, int reset);

/* depthwise causal conv k (weights [C, k]) + LayerNorm over C (ConvNeXt front half) */
int q3t_dwconv_ln(const float* x, const float* dw_w, const float* dw_b, const float* ln_w, const float* ln_b,
                  float ln_eps, int B, int T, int C, int ksize, float* out, void* stream);

/* sliding-window causal MHA for the codec transformer: qkv [B, T, 3*H*D] (RoPE is applied to q and k IN PLACE
 * first) -> out [B, T, H*D] */
int q3t_window_attn(float* qkv, const float* inv_freq, int B, int T, int H, int D, int window, float* out,
                    void* stream);

/* elementwise SnakeBeta on [rows, C]: a = exp(alpha), b = 1/(exp(beta)+1e-9) precomputed per channel */
int q3t_snake(const float* x, const float* a, const float* b, long long rows, int C, float* y, void* stream);

/* output convolution of the vocoder (mx.conv1d C -> 1, causal, dilation 1) fused with clamp(-1, 1):
 * act [B, T, C] fp32 time-major, W [taps, C] (tap j reads row t - (taps-1-j)), bias [1] or NULL -> wav [B, T] */
int q3t_conv_out_clamp(const float* act, int B, int T, int C, const float* W, const float* bias, int taps, float* wav,
                       void* stream);
/* the same with act as fp16 (written by a q3t_tapgemm with act_f16) */
int q3t_conv_out_clamp_h(const void* act_f16, int B, int T, int C, const float* W, const float* bias, int taps, float* wav,
                         void* stream);

/* final: clamp(x, -1, 1) and optional PCM16 conversion */
int q3t_clamp_pcm16(const float* x, long long n, float* y, int16_t* pcm, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Reference-clip side of voice cloning (replaces what generate_audio(ref_audio=, ref_text=) makes mlx_audio compute from the
 * clip: reference sessions/clone.py:218-224): operators of the speech-tokenizer ENCODER (Mimi) and of the ECAPA-TDNN speaker
 * encoder that are not convolutions / linears (those are q3t_tapgemm calls with force_fp32).  fp32, time-major [B, T, C].
 * ------------------------------------------------------------------------------------------- */
/* residual VQ encode (transformers mimi:1197-1203, 1262-1281): x [n_vectors, dim]; tables HOST array of n_levels device
 * ptrs [codebook_size, dim]; level l: idx = argmin_e sum_d (res - table_l[e])^2 (lowest index on ties), res -= table_l[idx].
 * idx_out[l * idx_level_stride + v] int32; resid_out [n_vectors, dim] (final residual) or NULL. */
int q3t_rvq_encode(const float* x, const float* const* tables_host, int n_vectors, int n_levels, int codebook_size, int dim,
                   long long idx_level_stride, int* idx_out, float* resid_out, void* stream);
/* (weighted) mean and std over time per channel: w [B, T, C] weights summing to 1 over T, or NULL for 1/T;
 * std = sqrt(max(sum w (x - mean)^2, eps))  (attentive statistics pooling, qwen2_5_omni:2654-2679) */
int q3t_time_stats(const float* x, const float* w, int B, int T, int C, float eps, float* mean, float* stdv, void* stream);
/* softmax over the time axis per (b, c) */
int q3t_softmax_time(const float* x, int B, int T, int C, float* y, void* stream);
/* op 0: out = a + b;  op 1: out = tanh(a);  op 2: out = a * b[item, c] + r  (b [n / per_item, C] channel gates, r or NULL) */
int q3t_eltwise(int op, const float* a, const float* b, const float* r, long long n, int C, long long per_item, float* out, void* stream);
/* mel front end tail: spec [rows, ld] = re (cols 0..n_freq-1) | im (cols n_freq..2 n_freq-1) of the STFT frames ->
 * out [rows, n_mels] = log(max(fb^T sqrt(re^2 + im^2 + 1e-9), 1e-5)),  fb [n_freq, n_mels] */
int q3t_mel(const float* spec, long long rows, int ld, int n_freq, const float* fb, int n_mels, float* out, void* stream);
/* LayerNorm over the channel axis with weight and bias */
int q3t_layernorm(const float* x, const float* w, const float* b, long long rows, int C, float eps, float* y, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* Q3TTS_B200_H */
