"""Host side of the B200 generation engine: owns device memory (weights, paged KV, state), builds the
C-ABI argument blocks once, captures the talker step and the whole frame into CUDA graphs and replays
them.  PyTorch is plumbing here (allocations, streams, graphs); every arithmetic kernel lives in
libq3tts_b200.so.

Mirrors what `mlx_audio`'s Qwen3-TTS `Model` does between `load_model` and the end of `generate`
(reference call sites: src/qwen3_tts/io.py:111-112, sessions/custom.py:163-170).
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import lib as L
from .config import ModelConfig
from .weights import WeightStore, pack_w8, quantize_w8


def _rope_inv_freq(head_dim: int, theta: float, device) -> torch.Tensor:
    # identical expression to the oracle (fp32 pow on the host)
    inv = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.float32) / head_dim))
    return inv.to(device)


class _Keep:
    """Keeps every tensor referenced by a ctypes struct alive."""

    def __init__(self):
        self.t: List[torch.Tensor] = []

    def __call__(self, t: torch.Tensor) -> int:
        assert t.is_cuda and t.is_contiguous()
        self.t.append(t)
        return t.data_ptr()


class TalkerEngine:
    """Talker + code predictor on one GPU for a fixed batch of `B` lock-step sequences."""

    def __init__(self, cfg: ModelConfig, ws: WeightStore, device: str = "cuda", batch: int = 1, max_frames: int = 512,
                 max_ctx: int = 2048, attn_nsplit: Optional[int] = None, keep_cp_logits: bool = False, max_trailing: int = 1,
                 use_mega: bool = True, prefill: str = "auto", prefill_gemm_rows: int = 32, kv_pages: Optional[int] = None):
        """kv_pages: size of the talker's K/V page POOL (pages of 16 tokens) when the block table is managed by a page
        allocator (serving.ContinuousBatcher): page 0 is a scratch page idle slots point at, the table starts all-zero.  None:
        every sequence owns max_ctx / 16 consecutive pages (identity table, lock-step batches)."""
        self.lib = L.load()
        self.cfg, self.dev, self.B = cfg, torch.device(device), batch
        if attn_nsplit is None:
            # context slices per (kv head, sequence) of the decode attention: few sequences need many CTAs; from 8 sequences on
            # the slices of a head form a thread-block cluster (<= 8) and four fill the machine (csrc/attn_decode.cu)
            attn_nsplit = 16 if batch <= 2 else (8 if batch < 32 else 4)
        self.max_frames, self.max_ctx, self.max_trailing = max_frames, max_ctx, max_trailing
        self.keep = _Keep()
        t, c = cfg.talker, cfg.cp
        self.G = c.num_code_groups
        dev = self.dev
        f32 = dict(device=dev, dtype=torch.float32)
        i32 = dict(device=dev, dtype=torch.int32)

        # ---- weights -------------------------------------------------------------------------
        self.w_bytes = 0

        def fp(name):
            return ws.fp[name].to(dev, torch.float32).contiguous()

        def w8(names: Sequence[str], bias_name: Optional[str] = None, interleave8: bool = False) -> L.W8:
            # fuse / interleave / pack WHERE THE CODES LIVE (host store: host work, no device launches before the first real
            # kernel; device store: on the device), then move the finished tiles
            trips = [ws.q[n] for n in names]
            q = torch.cat([x[0] for x in trips], 0)
            s = torch.cat([x[1] for x in trips], 0)
            b = torch.cat([x[2] for x in trips], 0)
            if interleave8:
                # fused gate/up: rows 16j..16j+7 = gate rows 8j.., rows 16j+8..16j+15 = the matching up rows, so one
                # 16-row weight tile holds both operands of silu(gate) * up (SwiGLU runs in the GEMV epilogue)
                half = q.shape[0] // 2
                idx = torch.arange(half, device=q.device).view(-1, 8)
                perm = torch.cat([idx, idx + half], 1).reshape(-1)
                q, s, b = q[perm].contiguous(), s[perm].contiguous(), b[perm].contiguous()
            blob = pack_w8(q, s, b).to(dev)
            self.w_bytes += blob.numel()
            o = L.W8()
            o.w, o.N, o.K = self.keep(blob), q.shape[0], q.shape[1]
            o.lin_bias = self.keep(fp(bias_name)) if bias_name else 0
            return o

        def stack(prefix: str, sc, n_pages: int, pages_per_seq: int, nsplit: int) -> Tuple[L.Stack, object]:
            layers = (L.Layer * sc.num_layers)()
            for i in range(sc.num_layers):
                p = f"{prefix}.layers.{i}"
                ly = layers[i]
                ly.input_norm = self.keep(fp(p + ".input_norm.weight"))
                ly.qkv = w8([p + ".q_proj", p + ".k_proj", p + ".v_proj"])
                ly.q_norm = self.keep(fp(p + ".q_norm.weight"))
                ly.k_norm = self.keep(fp(p + ".k_norm.weight"))
                ly.o = w8([p + ".o_proj"])
                ly.post_norm = self.keep(fp(p + ".post_norm.weight"))
                ly.gate_up = w8([p + ".gate_proj", p + ".up_proj"], interleave8=True)
                ly.down = w8([p + ".down_proj"])
            st = L.Stack()
            st.hidden, st.n_layers, st.n_heads, st.n_kv_heads = sc.hidden_size, sc.num_layers, sc.num_heads, sc.num_kv_heads
            st.head_dim, st.inter, st.eps = sc.head_dim, sc.intermediate_size, sc.rms_norm_eps
            st.layers_host = C.cast(layers, C.POINTER(L.Layer))
            raw = bytes(memoryview(layers).cast("B"))
            ldev = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)
            st.layers_dev = self.keep(ldev)
            st.final_norm = self.keep(fp(prefix + ".norm.weight"))
            st.inv_freq = self.keep(_rope_inv_freq(sc.head_dim, sc.rope_theta, dev))
            layer_elems = n_pages * 2 * sc.num_kv_heads * L.KV_PAGE * sc.head_dim
            pool = torch.zeros(sc.num_layers * layer_elems, device=dev, dtype=torch.bfloat16)
            st.kv_pool = self.keep(pool)
            st.kv_layer_stride_bytes = layer_elems * 2
            if prefix == "talker" and kv_pages is not None:
                tbl = torch.zeros(self.B, pages_per_seq, **i32)
            else:
                tbl = torch.arange(self.B * pages_per_seq, **i32).reshape(self.B, pages_per_seq).contiguous()
            if prefix == "talker":
                self.talker_tbl = tbl
            st.block_tbl, st.max_pages, st.attn_nsplit = self.keep(tbl), pages_per_seq, nsplit
            return st, layers

        pps = (max_ctx + L.KV_PAGE - 1) // L.KV_PAGE
        self.talker_stack, self._tl = stack("talker", t, self.B * pps if kv_pages is None else kv_pages + 1, pps, attn_nsplit)
        cp_pps = (self.G + 1 + L.KV_PAGE - 1) // L.KV_PAGE
        self.cp_stack, self._cl = stack("cp", c, self.B * cp_pps, cp_pps, 1)

        fa = L.FrameArgs()
        self.fa = fa
        fa.B = self.B
        fa.talker, fa.cp = self.talker_stack, self.cp_stack
        fa.codec_head = w8(["talker.codec_head"])
        fa.talker_vocab = t.vocab_size
        self.codec_embedding = fp("talker.codec_embedding")
        fa.codec_embedding = self.keep(self.codec_embedding)
        fa.cp_proj = w8(["cp.proj"], "cp.proj.bias")
        self.cp_embeddings = [fp(f"cp.embeddings.{g}") for g in range(self.G - 1)]
        self._cp_emb_host = (L.vp * (self.G - 1))(*[self.keep(e) for e in self.cp_embeddings])
        fa.cp_embeddings_host = C.cast(self._cp_emb_host, C.POINTER(L.vp))
        self._cp_emb_dev = torch.tensor([e.data_ptr() for e in self.cp_embeddings], device=dev, dtype=torch.int64)
        fa.cp_embeddings_dev = self.keep(self._cp_emb_dev)
        self._cp_heads = (L.W8 * (self.G - 1))(*[w8([f"cp.heads.{g}"]) for g in range(self.G - 1)])
        fa.cp_heads_host = C.cast(self._cp_heads, C.POINTER(L.W8))
        self._cp_heads_dev = torch.frombuffer(bytearray(bytes(memoryview(self._cp_heads).cast("B"))), dtype=torch.uint8).to(dev)
        fa.cp_heads_dev = self.keep(self._cp_heads_dev)
        fa.cp_vocab, fa.n_groups = c.vocab_size, self.G
        self._cp_proj_tabs = None               # built lazily (batch 1, persistent kernel): see _ensure_cp_proj_rows
        # text side (prefill only)
        self.text_embedding = fp("talker.text_embedding")
        self.tp_fc1 = w8(["talker.text_projection.fc1"], "talker.text_projection.fc1.bias")
        self.tp_fc2 = w8(["talker.text_projection.fc2"], "talker.text_projection.fc2.bias")

        # ---- state -------------------------------------------------------------------------------
        B, H, Hc, V, Vc, G = self.B, t.hidden_size, c.hidden_size, t.vocab_size, c.vocab_size, self.G
        self.x = torch.zeros(B, H, **f32)
        self.hidden = torch.zeros(B, H, **f32)
        self.logits = torch.zeros(B, V, **f32)
        self.cp_logits = torch.zeros((G - 1) if keep_cp_logits else 1, B, Vc, **f32)
        self.xc = torch.zeros(B, Hc, **f32)
        qkvd = max(t.q_dim + 2 * t.kv_dim, c.q_dim + 2 * c.kv_dim)
        self.qkv = torch.zeros(B, qkvd, **f32)
        self.attn = torch.zeros(B, max(t.q_dim, c.q_dim), **f32)
        self.gu = torch.zeros(B, 2 * max(t.intermediate_size, c.intermediate_size), **f32)
        rep = max(t.num_heads // t.num_kv_heads, c.num_heads // c.num_kv_heads)
        self.attn_work = torch.zeros(B * max(t.num_kv_heads, c.num_kv_heads) * max(attn_nsplit, 1) * rep *
                                     (max(t.head_dim, c.head_dim) + 2), **f32)
        self.attn_counters = torch.zeros(B * max(t.num_kv_heads, c.num_kv_heads), **i32)
        self.pos = torch.zeros(B, **i32)
        self.cp_pos = torch.arange(G + 1, **i32).repeat_interleave(B).contiguous()
        self.step = torch.zeros(1, **i32)
        self.cur_codes = torch.zeros(B, G, **i32)
        self.codes = torch.zeros(B, max_frames, G, **i32)
        self.own_codes = torch.zeros(B, max_frames, G, **i32)
        self.forced = torch.zeros(B, max_frames, G, **i32)
        self.seen = torch.zeros(B, (V + 31) // 32, device=dev, dtype=torch.int32)
        self.done = torch.zeros(B, **i32)
        self.trailing = torch.zeros(B, max_trailing, H, **f32)
        fa.x, fa.hidden, fa.logits, fa.cp_logits = map(self.keep, (self.x, self.hidden, self.logits, self.cp_logits))
        fa.keep_cp_logits = int(keep_cp_logits)
        fa.xc, fa.qkv, fa.attn, fa.gu = map(self.keep, (self.xc, self.qkv, self.attn, self.gu))
        fa.attn_work, fa.attn_counters = self.keep(self.attn_work), self.keep(self.attn_counters)
        fa.pos, fa.cp_pos, fa.step = self.keep(self.pos), self.keep(self.cp_pos), self.keep(self.step)
        fa.cur_codes, fa.codes, fa.own_codes = self.keep(self.cur_codes), self.keep(self.codes), self.keep(self.own_codes)
        fa.max_frames = max_frames
        fa.seen, fa.done = self.keep(self.seen), self.keep(self.done)
        fa.trailing, fa.n_trailing = self.keep(self.trailing), max_trailing
        fa.forced_codes = 0
        # persistent kernels (batch 1): exchange buffers of 64-bit {value, tag} words + [tag counter, error code]
        nbytes = int(self.lib.q3t_ll_work_bytes(C.byref(self.talker_stack), C.byref(self.cp_stack), max(V, Vc)))
        self.ll_work = torch.zeros(nbytes // 8, device=dev, dtype=torch.int64)
        self.ll_state = torch.zeros(4, **i32)
        fa.ll_work, fa.ll_work_bytes, fa.ll_state, fa.ll_timing = self.keep(self.ll_work), nbytes, self.keep(self.ll_state), 0
        fa.use_mega = int(use_mega and self.B == 1)
        # tcgen05 GEMM path (more than two rows per contraction): bf16 activation scratch
        kmax = max(t.hidden_size, t.intermediate_size, c.hidden_size, c.intermediate_size, t.q_dim, c.q_dim)
        # split bf16 rows [hi(K) | lo(K)]: the tcgen05 GEMM carries every operand as hi + lo (csrc/w8_gemm_tc.cu)
        self.gemm_xb = torch.empty(2 * max(B, 1) * kmax, device=dev, dtype=torch.bfloat16)
        fa.gemm_xb = self.keep(self.gemm_xb)
        self.gemm_xb2 = torch.empty(2 * max(B, 1) * kmax, device=dev, dtype=torch.bfloat16)
        fa.gemm_xb2 = self.keep(self.gemm_xb2)
        nmax = max(2 * t.intermediate_size, 2 * c.intermediate_size, V, Vc, t.q_dim + 2 * t.kv_dim)
        self.gemm_ws = torch.empty(8 * max(B, 1) * nmax, **f32)
        self.gemm_counters = torch.zeros(1024, **i32)
        fa.gemm_ws, fa.gemm_ws_floats, fa.gemm_counters = self.keep(self.gemm_ws), self.gemm_ws.numel(), self.keep(self.gemm_counters)
        # row statistics of the deferred RMSNorm between the GEMMs of the batched path (include/q3tts_b200.h: q3t_gemm_args.y_rowss)
        self.gemm_rowss = torch.zeros(32 * max(B, 1), **f32)
        if not os.environ.get("Q3T_NO_NORM_DEFER"):
            fa.gemm_rowss = self.keep(self.gemm_rowss)
        # prompt rows go through the tcgen05 W8 GEMM (bf16 operands, the logit tolerance of BASELINE.json) when there are
        # enough of them; "decode" pins the token-by-token path through the decode kernels (exact-integer contractions)
        assert prefill in ("auto", "gemm", "decode")
        self.prefill_mode, self.prefill_gemm_rows = prefill, prefill_gemm_rows
        self.gemm_prefill = self.B > 2 or prefill == "gemm"
        self.set_sampling()
        self._graphs: Dict[str, torch.cuda.CUDAGraph] = {}
        self.use_graphs = True
        self.launches_per_frame = None
        self.launches: Dict[str, int] = {}

    # ---- configuration ---------------------------------------------------------------------------
    def set_sampling(self, do_sample: bool = False, temperature: float = 0.9, top_k: int = 50, top_p: float = 1.0,
                     repetition_penalty: float = 1.05, min_new_tokens: int = 2, seed: int = 0,
                     cp_do_sample: Optional[bool] = None, cp_temperature: float = 0.9, cp_top_k: int = 50,
                     cp_top_p: float = 1.0):
        """Greedy parity definition: argmax after the suppress mask, penalty 1.0, no
        min_new_tokens.  Sampling defaults follow SURVEY App. A."""
        t = self.cfg.talker
        sp = self.fa.talker_sp
        sp.do_sample = int(do_sample)
        sp.temperature, sp.top_k, sp.top_p = temperature, top_k, top_p
        sp.repetition_penalty = repetition_penalty if do_sample else 1.0
        sp.min_new_tokens = min_new_tokens if do_sample else 0
        sp.suppress_lo, sp.suppress_hi, sp.eos_id = t.vocab_size - 1024, t.vocab_size, t.codec_eos_id
        sp.seed = seed
        cs = self.fa.cp_sp
        cs.do_sample = int(do_sample if cp_do_sample is None else cp_do_sample)
        cs.temperature, cs.top_k, cs.top_p, cs.repetition_penalty = cp_temperature, cp_top_k, cp_top_p, 1.0
        cs.min_new_tokens, cs.suppress_lo, cs.suppress_hi, cs.eos_id, cs.seed = 0, -1, -1, -1, seed + 1
        # the sampling block is baked into the captured launches: re-capture only when a parameter actually changed
        # (Model.generate calls this for every utterance and every long-text segment)
        key = tuple(getattr(x, f) for x in (sp, cs) for f, _ in L.Sampling._fields_)
        if key != getattr(self, "_sampling_key", None):
            self._sampling_key = key
            self._graphs = {}

    def set_mega(self, on: bool):
        """Switch between the persistent stack-pass kernel (batch 1) and the one-kernel-per-contraction path."""
        self.fa.use_mega = int(on and self.B == 1)
        self._graphs = {}

    def set_forced(self, forced: Optional[torch.Tensor]):
        """Teacher forcing for parity tests: forced [B, T, G] int."""
        if forced is None:
            self.fa.forced_codes = 0
        else:
            T = forced.shape[1]
            self.forced.zero_()
            self.forced[:, :T] = forced.to(self.dev, torch.int32)
            self.fa.forced_codes = self.forced.data_ptr()
        self._graphs = {}

    def _ensure_cp_proj_rows(self):
        """Projected embedding tables (q3t_frame_args.cp_proj_rows_dev; persistent kernel and batched frame alike): the input of
        code-predictor pass g is cp_proj(table_g[code]), a function of one sampled code, so the projection of every table
        row is computed once here with the same W8 GEMV the kernel would run (31.7 k rows; 0.13 GB + 0.52 GB for the q|k|v
        tables at full size) and each pass starts from lookups instead of two contraction phases.  Q3T_CP_PROJ_TABLES=0 keeps the in-kernel projection."""
        if self._cp_proj_tabs is not None or os.environ.get("Q3T_CP_PROJ_TABLES", "1") == "0":
            return
        tabs = [self.codec_embedding] + self.cp_embeddings[: self.G - 2]
        self._cp_proj_tabs = []
        for tab in tabs:
            y = torch.empty(tab.shape[0], self.cfg.cp.hidden_size, device=self.dev, dtype=torch.float32)
            self._gemv_rows(self.fa.cp_proj, tab, y)
            self._cp_proj_tabs.append(y)
        self._cp_proj_dev = torch.tensor([t.data_ptr() for t in self._cp_proj_tabs], device=self.dev, dtype=torch.int64)
        # ... and the first layer's q|k|v of RMSNorm(projected row): the first layer of those passes starts at the attention
        self._cp_qkv0_tabs = []
        ly0 = self._cl[0]
        for tab in self._cp_proj_tabs:
            y = torch.empty(tab.shape[0], ly0.qkv.N, device=self.dev, dtype=torch.float32)
            self._gemv_rows(ly0.qkv, tab, y, prologue=L.PRO_RMSNORM, norm_w=int(ly0.input_norm), eps=self.cp_stack.eps)
            self._cp_qkv0_tabs.append(y)
        self._cp_qkv0_dev = torch.tensor([t.data_ptr() for t in self._cp_qkv0_tabs], device=self.dev, dtype=torch.int64)
        torch.cuda.synchronize()
        self.fa.cp_proj_rows_dev = self._cp_proj_dev.data_ptr()
        if os.environ.get("Q3T_CP_QKV0_TABLES", "1") != "0":
            self.fa.cp_qkv0_rows_dev = self._cp_qkv0_dev.data_ptr()

    # ---- embeddings for the prefill (W8 GEMVs, two rows per launch) ------------------------------------
    def text_embed(self, ids: torch.Tensor) -> torch.Tensor:
        """P(ids) = fc2(silu(fc1(text_embedding[ids]))) -> [n, H] on device."""
        ids = ids.to(self.dev).long().reshape(-1)
        e = self.text_embedding[ids].contiguous()
        n, Ht = e.shape
        h = torch.empty(n, Ht, device=self.dev, dtype=torch.float32)
        o = torch.empty(n, self.cfg.talker.hidden_size, device=self.dev, dtype=torch.float32)
        self._gemv_rows(self.tp_fc1, e, h, act=L.ACT_SILU)
        self._gemv_rows(self.tp_fc2, h, o)
        return o

    def _gemv_rows(self, w: L.W8, x: torch.Tensor, y: torch.Tensor, act: int = 0, prologue: int = 0,
                   norm_w: Optional[torch.Tensor] = None, eps: float = 0.0, resid: Optional[torch.Tensor] = None):
        n = x.shape[0]
        a = L.GemvArgs()
        a.w, a.M, a.prologue = w, min(2, n), prologue
        a.x, a.x_stride = x.data_ptr(), x.stride(0)
        a.norm_w, a.eps = (norm_w if isinstance(norm_w, int) else L.ptr(norm_w)), eps
        a.act = act
        if resid is not None:
            a.resid, a.resid_stride = resid.data_ptr(), resid.stride(0)
        a.y, a.y_stride = y.data_ptr(), y.stride(0)
        L.check(self.lib.q3t_w8_gemv_rows(C.byref(a), n, L.stream_ptr()), "w8_gemv_rows")   # two rows per launch, looped in C

    # ---- generation --------------------------------------------------------------------------------
    def reset(self):
        for t in (self.pos, self.step, self.seen, self.done, self.codes, self.own_codes, self.attn_counters):
            t.zero_()

    def _talker_step(self, want_logits: bool):
        L.check(self.lib.q3t_talker_step(C.byref(self.fa), int(want_logits), L.stream_ptr()), "talker_step")

    def _frame(self):
        L.check(self.lib.q3t_frame(C.byref(self.fa), L.stream_ptr()), "frame")

    def _ensure_graphs(self):
        """Capture the talker step (with / without logits) and the whole frame once.  The warm-up run
        mutates state, so this is only called right before a reset()."""
        if self._graphs or not self.use_graphs:
            return
        fns = {"step": lambda: self._talker_step(False), "step_logits": lambda: self._talker_step(True),
               "frame": self._frame}
        for fn in fns.values():          # warm-up outside capture (lazy module load, func attributes)
            fn()
        torch.cuda.synchronize()
        self.reset()
        for key, fn in fns.items():
            n0 = self.lib.q3t_launch_count()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            self._graphs[key] = g
            self.launches[key] = int(self.lib.q3t_launch_count() - n0)
            if key == "frame":
                self.launches_per_frame = self.launches[key]
        torch.cuda.synchronize()

    def _run(self, key: str):
        if self.use_graphs:
            self._graphs[key].replay()
        elif key == "frame":
            self._frame()
        else:
            self._talker_step(key == "step_logits")

    def prefill(self, embeds: torch.Tensor, lengths: Optional[Sequence[int]] = None,
                trailing: Optional[torch.Tensor] = None):
        """embeds [B, Lmax, H] (RIGHT-aligned when lengths differ); trailing [B, n_tr, H] whose last row is the
        tts_pad embedding (rows past n_tr repeat it).  v1 prefill feeds one token per talker step."""
        B, Lmax, H = embeds.shape
        assert B == self.B
        embeds = embeds.to(self.dev, torch.float32)
        lengths = list(lengths) if lengths is not None else [Lmax] * B
        assert max(lengths) == Lmax and Lmax + self.max_frames <= self.max_ctx
        self._ensure_cp_proj_rows()
        self._ensure_graphs()
        self.reset()
        if trailing is None:
            trailing = torch.zeros(B, 1, H)
        n_tr = trailing.shape[1]
        assert n_tr <= self.max_trailing, "raise max_trailing for streaming text"
        tr = trailing.to(self.dev, torch.float32)
        self.trailing[:, :n_tr] = tr
        self.trailing[:, n_tr:] = tr[:, -1:]
        if self.prefill_mode != "decode" and (self.gemm_prefill or sum(lengths) >= self.prefill_gemm_rows):
            self._prefill_gemm(embeds, lengths)
        else:
            off = torch.tensor([Lmax - l for l in lengths], device=self.dev, dtype=torch.int32)
            for tkn in range(Lmax):
                self.pos.copy_((tkn - off).clamp_(min=0))
                self.x.copy_(embeds[:, tkn])
                self._run("step_logits" if tkn == Lmax - 1 else "step")
        self.pos.copy_(torch.tensor(lengths, device=self.dev, dtype=torch.int32))

    def _prefill_gemm(self, embeds: torch.Tensor, lengths: Sequence[int]):
        """All prompt tokens of all sequences as ONE set of GEMM rows per layer (tcgen05 W8 GEMM), no padding rows:
        row m = (sequence, position).  Two attention passes per layer: write every K/V row, then causal attention."""
        t = self.cfg.talker
        B, Lmax, H = embeds.shape
        dev = self.dev
        rows = torch.cat([embeds[b, Lmax - l:] for b, l in enumerate(lengths)], 0).contiguous()
        rows = self.prefill_rows(rows, lengths, list(range(B)))
        last = torch.tensor([sum(lengths[:b + 1]) - 1 for b in range(B)], device=dev)
        self.x.copy_(rows[last])
        L.check(self.lib.q3t_talker_tail(C.byref(self.fa), L.stream_ptr()), "talker_tail")
        torch.cuda.synchronize()          # the temporaries must outlive the enqueued kernels

    def prefill_rows(self, rows: torch.Tensor, lengths: Sequence[int], slots: Sequence[int]) -> torch.Tensor:
        """rows [M, H] = the prompt tokens of len(lengths) sequences, concatenated; sequence i occupies block-table row
        slots[i].  Writes their K/V pages and returns the residual stream after the last layer (pre final norm) [M, H]."""
        t = self.cfg.talker
        dev = self.dev
        rows = rows.to(dev, torch.float32).contiguous()
        M = rows.shape[0]
        assert M == sum(lengths)
        pos = torch.cat([torch.arange(l, dtype=torch.int32) for l in lengths]).to(dev)
        seq = torch.cat([torch.full((l,), b, dtype=torch.int32) for b, l in zip(slots, lengths)]).to(dev)
        qkvd, rep = t.q_dim + 2 * t.kv_dim, t.num_heads // t.num_kv_heads
        f32 = dict(device=dev, dtype=torch.float32)
        qkv = torch.empty(M, qkvd, **f32)
        attn = torch.empty(M, t.q_dim, **f32)
        gu = torch.empty(M, 2 * t.intermediate_size, **f32)
        xb = torch.empty(2 * M * max(t.hidden_size, t.intermediate_size, t.q_dim), device=dev, dtype=torch.bfloat16)   # split rows
        work = torch.empty(M * t.num_kv_heads * rep * (t.head_dim + 2), **f32)
        counters = torch.zeros(M * t.num_kv_heads, device=dev, dtype=torch.int32)
        a = L.PrefillArgs()
        a.f, a.M = C.pointer(self.fa), M
        a.x, a.pos, a.seq_of_row = rows.data_ptr(), pos.data_ptr(), seq.data_ptr()
        a.qkv, a.attn, a.gu, a.xb = qkv.data_ptr(), attn.data_ptr(), gu.data_ptr(), xb.data_ptr()
        a.attn_work, a.attn_counters = work.data_ptr(), counters.data_ptr()
        # runs of <= 32 consecutive rows of one sequence: one CTA of the tensor-core prefill attention each
        blocks, r0 = [], 0
        for l in lengths:
            blocks += [(r0 + o, min(32, l - o)) for o in range(0, l, 32)]
            r0 += l
        blk = torch.tensor(blocks, dtype=torch.int32, device=dev).contiguous()
        xb2 = torch.empty(2 * M * max(t.intermediate_size, t.q_dim), device=dev, dtype=torch.bfloat16)
        if not os.environ.get("Q3T_PREFILL_ATTN_PER_ROW"):
            a.blocks, a.n_blocks = blk.data_ptr(), len(blocks)
            if not os.environ.get("Q3T_NO_BF16_CHAIN"):
                a.xb2 = xb2.data_ptr()
                if not os.environ.get("Q3T_NO_NORM_DEFER"):
                    rowss = torch.empty(32 * M, **f32)
                    a.rowss = rowss.data_ptr()
        L.check(self.lib.q3t_talker_prefill(C.byref(a), L.stream_ptr()), "talker_prefill")
        torch.cuda.synchronize()          # the temporaries above must outlive the enqueued kernels
        return rows

    def generate(self, n_frames: int, check_every: int = 16) -> torch.Tensor:
        """Runs up to n_frames frames (stops early once every sequence sampled EOS). Returns codes [B, T, G]."""
        assert n_frames <= self.max_frames
        for f in range(n_frames):
            self._run("frame")
            if check_every and (f + 1) % check_every == 0 and f + 1 < n_frames:
                if bool(self.done.all().item()):
                    n_frames = f + 1
                    break
        return self.codes[:, :n_frames]
