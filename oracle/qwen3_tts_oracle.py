"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product path.

PyTorch-CPU fp32 restatement of the Qwen3-TTS-12Hz generation hot path that the reference
repo executes through its un-vendored dependency ``mlx-audio==0.3.1`` (+ ``mlx==0.30.3``;
pinned at /root/reference/pyproject.toml:38-41, called from src/qwen3_tts/io.py:111-112 and
sessions/custom.py:163-170, design.py:76-81, clone.py:218-224).

PARITY UNPINNED: the reference's own tests hold no golden vector, known-answer test or audio
fixture for this path (tests/test_sessions_smoke.py:6-11 checks import-ability only), and neither
mlx nor the upstream QwenLM package can be installed offline.  The restatement is therefore pinned
against the structurally identical classes that ARE on disk in transformers 5.5.0 (the "cousins",
see tests/test_oracle_vs_cousins.py):
    T = transformers/models
    talker / code-predictor layer   T/qwen3/modeling_qwen3.py:50-334
    code predictor model + heads    T/qwen3_omni_moe/modeling_qwen3_omni_moe.py:2352-2718
    frame loop (next-input sum)     T/qwen3_omni_moe/modeling_qwen3_omni_moe.py:3243-3279
    prefill layout                  T/qwen3_omni_moe/modeling_qwen3_omni_moe.py:3838-3893 + SURVEY App. C
    sampler                         transformers/generation/logits_process.py:236,302,469,536,1865
    split RVQ decode                T/mimi/modeling_mimi.py:1176-1349
    codec decoder                   T/qwen3_omni_moe/modeling_qwen3_omni_moe.py:3283-3790

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


# ======================================================================================
# primitives
# ======================================================================================
def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    """qwen3/modeling_qwen3.py:59-64 (fp32 variance, weight multiplied last)."""
    x = x.float()
    var = x.pow(2).mean(-1, keepdim=True)
    return w * (x * torch.rsqrt(var + eps))


def rope_cos_sin(positions: torch.Tensor, head_dim: int, theta: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """qwen3/modeling_qwen3.py:137-181: inv_freq = theta^(-2i/d); emb = cat(freqs, freqs)."""
    inv = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.float32) / head_dim))
    fr = positions.float()[:, None] * inv[None, :]
    emb = torch.cat([fr, fr], dim=-1)
    return emb.cos(), emb.sin()


def rotate_half(x: torch.Tensor) -> torch.Tensor:
    h = x.shape[-1] // 2
    return torch.cat([-x[..., h:], x[..., :h]], dim=-1)


def apply_rope(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """x [..., T, heads, D]; cos/sin [T, D]."""
    return x * cos[:, None, :] + rotate_half(x) * sin[:, None, :]


def kv_round(x: torch.Tensor, kv_dtype: Optional[torch.dtype]) -> torch.Tensor:
    """The reference runs its KV cache in the model dtype (bf16).  The device path stores K/V pages in
    bf16; passing kv_dtype=torch.bfloat16 makes the oracle round at the same point."""
    return x if kv_dtype is None else x.to(kv_dtype).float()


# ======================================================================================
# dense Qwen3 decoder stack (talker: 28 x 2048, code predictor: 5 x 1024)
# ======================================================================================
class DecoderStack:
    def __init__(self, w: Dict[str, torch.Tensor], prefix: str, n_layers: int, n_heads: int, n_kv: int,
                 head_dim: int, eps: float, theta: float, kv_dtype: Optional[torch.dtype] = None):
        self.w, self.p = w, prefix
        self.L, self.H, self.Hkv, self.D = n_layers, n_heads, n_kv, head_dim
        self.eps, self.theta, self.kv_dtype = eps, theta, kv_dtype
        self.reset()

    def reset(self):
        self.k: List[Optional[torch.Tensor]] = [None] * self.L   # [T, Hkv, D]
        self.v: List[Optional[torch.Tensor]] = [None] * self.L
        self.pos = 0

    def forward(self, x: torch.Tensor, final_norm: bool = True) -> torch.Tensor:
        """x [T, hidden] new tokens (appended to the cache).  Returns hidden [T, hidden]."""
        w, p = self.w, self.p
        T = x.shape[0]
        pos = torch.arange(self.pos, self.pos + T)
        cos, sin = rope_cos_sin(pos, self.D, self.theta)
        rep = self.H // self.Hkv
        for i in range(self.L):
            lp = f"{p}.layers.{i}"
            h = rms_norm(x, w[lp + ".input_norm.weight"], self.eps)
            q = (h @ w[lp + ".q_proj.weight"].T).view(T, self.H, self.D)
            k = (h @ w[lp + ".k_proj.weight"].T).view(T, self.Hkv, self.D)
            v = (h @ w[lp + ".v_proj.weight"].T).view(T, self.Hkv, self.D)
            q = apply_rope(rms_norm(q, w[lp + ".q_norm.weight"], self.eps), cos, sin)
            k = apply_rope(rms_norm(k, w[lp + ".k_norm.weight"], self.eps), cos, sin)
            k, v = kv_round(k, self.kv_dtype), kv_round(v, self.kv_dtype)
            self.k[i] = k if self.k[i] is None else torch.cat([self.k[i], k], 0)
            self.v[i] = v if self.v[i] is None else torch.cat([self.v[i], v], 0)
            K = self.k[i].repeat_interleave(rep, dim=1)         # [S, H, D]
            V = self.v[i].repeat_interleave(rep, dim=1)
            S = K.shape[0]
            scores = torch.einsum("thd,shd->hts", q, K) * (self.D ** -0.5)
            causal = torch.arange(S)[None, :] <= pos[:, None]
            scores = scores.masked_fill(~causal[None], float("-inf"))
            pr = torch.softmax(scores, dim=-1)                   # fp32 softmax (qwen3:213)
            a = torch.einsum("hts,shd->thd", pr, V).reshape(T, self.H * self.D)
            x = x + a @ w[lp + ".o_proj.weight"].T
            h = rms_norm(x, w[lp + ".post_norm.weight"], self.eps)
            g = h @ w[lp + ".gate_proj.weight"].T
            u = h @ w[lp + ".up_proj.weight"].T
            x = x + (F.silu(g) * u) @ w[lp + ".down_proj.weight"].T
        self.pos += T
        return rms_norm(x, w[p + ".norm.weight"], self.eps) if final_norm else x


# ======================================================================================
# sampler  (order = generation/utils.py:1088 -> 1129 -> 1172 -> 1213 -> 1217 -> 1221)
# ======================================================================================
@dataclass
class SamplingParams:
    do_sample: bool = False
    temperature: float = 1.0
    top_k: int = 0
    top_p: float = 1.0
    repetition_penalty: float = 1.0
    min_new_tokens: int = 0
    suppress_lo: int = -1          # suppress ids in [suppress_lo, suppress_hi) except eos_id
    suppress_hi: int = -1
    eos_id: int = -1


def process_logits(logits: torch.Tensor, sp: SamplingParams, history: Sequence[int], n_generated: int) -> torch.Tensor:
    """logits [V] fp32 -> processed scores [V] (before softmax / argmax)."""
    s = logits.clone().float()
    if sp.repetition_penalty != 1.0 and len(history):
        ids = torch.tensor(sorted(set(int(h) for h in history)), dtype=torch.long)
        sc = s[ids]
        s[ids] = torch.where(sc < 0, sc * sp.repetition_penalty, sc / sp.repetition_penalty)  # logits_process.py:302
    if sp.eos_id >= 0 and n_generated < sp.min_new_tokens:
        s[sp.eos_id] = float("-inf")                                                       # logits_process.py:164
    if sp.suppress_lo >= 0:
        keep = s[sp.eos_id].clone() if sp.suppress_lo <= sp.eos_id < sp.suppress_hi else None
        s[sp.suppress_lo:sp.suppress_hi] = float("-inf")                                   # logits_process.py:1865
        if keep is not None:
            s[sp.eos_id] = keep
    if not sp.do_sample:
        return s
    if sp.temperature != 1.0:
        s = s / sp.temperature                                                             # logits_process.py:236
    if sp.top_k and sp.top_k > 0:
        kk = min(sp.top_k, s.numel())
        thr = torch.topk(s, kk).values[-1]
        s = s.masked_fill(s < thr, float("-inf"))                                          # logits_process.py:536
    if sp.top_p < 1.0:
        srt, idx = torch.sort(s, descending=False)
        cum = torch.softmax(srt, dim=-1).cumsum(-1)
        rm = cum <= (1.0 - sp.top_p)
        rm[-1:] = False                                                                    # min_tokens_to_keep=1
        s = s.masked_fill(torch.zeros_like(rm).scatter(0, idx, rm), float("-inf"))         # logits_process.py:469
    return s


def draw(scores: torch.Tensor, sp: SamplingParams, u: Optional[float] = None) -> int:
    """argmax (lowest index wins ties) or inverse-CDF categorical draw with a supplied uniform `u`
    (RNG streams cannot match across frameworks: stochastic parity feeds identical uniforms)."""
    if not sp.do_sample:
        return int(torch.argmax(scores))
    p = torch.softmax(scores.float(), dim=-1)
    cdf = p.cumsum(-1)
    u = float(torch.rand(())) if u is None else u
    idx = int(torch.searchsorted(cdf, torch.tensor(u * float(cdf[-1]), dtype=cdf.dtype), right=True))
    idx = min(idx, scores.numel() - 1)
    while p[idx] == 0 and idx > 0:   # never return a masked id
        idx -= 1
    return idx


# ======================================================================================
# model = talker + code predictor (+ prefill builder + frame loop)
# ======================================================================================
class OracleModel:
    def __init__(self, cfg, weights: Dict[str, torch.Tensor], kv_dtype: Optional[torch.dtype] = None):
        self.cfg, self.w = cfg, weights
        t, c = cfg.talker, cfg.cp
        self.talker = DecoderStack(weights, "talker", t.num_layers, t.num_heads, t.num_kv_heads, t.head_dim,
                                   t.rms_norm_eps, t.rope_theta, kv_dtype)
        self.cp = DecoderStack(weights, "cp", c.num_layers, c.num_heads, c.num_kv_heads, c.head_dim,
                               c.rms_norm_eps, c.rope_theta, kv_dtype)

    # ---- embeddings ---------------------------------------------------------------
    def text_embed(self, ids: torch.Tensor) -> torch.Tensor:
        """P(x) = text_projection(text_embedding(x)): Linear+bias, SiLU, Linear+bias."""
        w = self.w
        e = w["talker.text_embedding"][ids.long()]
        h = F.silu(e @ w["talker.text_projection.fc1.weight"].T + w["talker.text_projection.fc1.bias"])
        return h @ w["talker.text_projection.fc2.weight"].T + w["talker.text_projection.fc2.bias"]

    def codec_embed(self, ids: torch.Tensor) -> torch.Tensor:
        return self.w["talker.codec_embedding"][ids.long()]

    # ---- prefill layout (cousin qwen3_omni_moe:3838-3893) ---------
    def ref_code_embeds(self, ref_codes: torch.Tensor) -> torch.Tensor:
        """[T_ref, G] codes of a reference clip -> [T_ref, H]: per frame the same 16-way embedding sum the frame loop feeds back
        (group 0 from the talker's codec table, groups 1.. from the code predictor's tables, added in order g = 0..15)."""
        w = self.w
        acc = w["talker.codec_embedding"][ref_codes[:, 0].long()]
        for g in range(1, ref_codes.shape[1]):
            acc = acc + w[f"cp.embeddings.{g - 1}"][ref_codes[:, g].long()]
        return acc

    def build_prefill(self, text_ids: Sequence[int], instruct_ids: Optional[Sequence[int]] = None,
                      speaker: Optional[str] = None, language: Optional[str] = None,
                      speaker_vec: Optional[torch.Tensor] = None, streaming: bool = False,
                      ref_codes: Optional[torch.Tensor] = None, ref_text_ids: Optional[Sequence[int]] = None
                      ) -> Tuple[torch.Tensor, torch.Tensor]:
        """text_ids = chat-templated ids: ids[:3] role prefix, ids[3:-5] body, ids[-5:] template tail.
        Returns (prefill embeds [L, H], trailing text embeds [n_trailing, H] whose LAST row is tts_pad).

        ICL (Base model + reference clip, SURVEY App. C last line; reference call site sessions/clone.py:218-224): with
        `ref_codes` [T_ref, G] (+ `ref_text_ids` = chat-templated "<|im_start|>assistant\n{ref_text}<|im_end|>\n", body =
        ids[3:-2]) the text side becomes P(ref body ++ text body) ++ eos and is paired position by position with the codec side
        E(codec_bos) ++ [sum_g emb_g(ref_codes[t])]: streaming pairs them (the longer side's rest trails / is padded with
        tts_pad), non-streaming lays the text block (on codec_pad) before the code block (on tts_pad)."""
        cfg, t = self.cfg, self.cfg.talker
        ids = torch.as_tensor(list(text_ids), dtype=torch.long)
        pad, bos, eos = self.text_embed(torch.tensor([cfg.tts_pad_token_id, cfg.tts_bos_token_id,
                                                      cfg.tts_eos_token_id]))
        if language is not None and language.lower() in t.codec_language_id:
            prefix = [t.codec_think_id, t.codec_think_bos_id, t.codec_language_id[language.lower()],
                      t.codec_think_eos_id]
        else:
            prefix = [t.codec_nothink_id, t.codec_think_bos_id, t.codec_think_eos_id]
        parts = [self.codec_embed(torch.tensor(prefix))]
        if speaker_vec is not None:
            parts.append(speaker_vec.float().view(1, -1))
        elif speaker is not None:
            parts.append(self.codec_embed(torch.tensor([t.spk_id[speaker.lower()]])))
        parts.append(self.codec_embed(torch.tensor([t.codec_pad_id, t.codec_bos_id])))
        codec_seq = torch.cat(parts, 0)
        n = codec_seq.shape[0]
        head = self.text_embed(ids[:3])
        mid = torch.cat([pad.expand(n - 2, -1), bos[None]], 0) + codec_seq[:-1]
        body_ids = ids[3:-5]
        segs = []
        if instruct_ids is not None and len(instruct_ids):
            segs.append(self.text_embed(torch.as_tensor(list(instruct_ids), dtype=torch.long)))
        segs += [head, mid]
        if ref_codes is not None:
            ref_body = torch.as_tensor(list(ref_text_ids), dtype=torch.long)[3:-2] if ref_text_ids is not None \
                else torch.zeros(0, dtype=torch.long)
            text_all = torch.cat([self.text_embed(torch.cat([ref_body, body_ids])), eos[None]], 0)          # [T1, H]
            codec_all = torch.cat([self.codec_embed(torch.tensor([t.codec_bos_id])), self.ref_code_embeds(ref_codes)], 0)
            t1, t2 = text_all.shape[0], codec_all.shape[0]
            if streaming:
                if t1 > t2:
                    segs.append(text_all[:t2] + codec_all)
                    trailing = torch.cat([text_all[t2:], pad[None]], 0)
                else:
                    segs.append(torch.cat([text_all, pad[None].expand(t2 - t1, -1)], 0) + codec_all)
                    trailing = pad[None]
            else:
                segs.append(text_all + self.codec_embed(torch.full((t1,), t.codec_pad_id)))
                segs.append(codec_all + pad[None])
                trailing = pad[None]
        elif not streaming:
            body = torch.cat([self.text_embed(body_ids), eos[None]], 0) + \
                self.codec_embed(torch.full((len(body_ids) + 1,), t.codec_pad_id))
            tail = pad[None] + self.codec_embed(torch.tensor([t.codec_bos_id]))
            segs += [body, tail]
            trailing = pad[None]
        else:
            first = self.text_embed(body_ids[:1]) + codec_seq[-1:]
            segs.append(first)
            trailing = torch.cat([self.text_embed(body_ids[1:]), eos[None], pad[None]], 0)
        return torch.cat(segs, 0), trailing

    # ---- talker ---------------------------------------------------------------------
    def talker_forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """x [T, H] -> (normed hidden [T, H], codec_head logits [T, V])."""
        h = self.talker.forward(x)
        return h, h @ self.w["talker.codec_head.weight"].T

    def talker_sampling(self, sp: Optional[SamplingParams] = None) -> SamplingParams:
        """Greedy parity definition: argmax after the suppress mask, penalty 1.0."""
        t = self.cfg.talker
        sp = sp or SamplingParams()
        sp.suppress_lo, sp.suppress_hi, sp.eos_id = t.vocab_size - 1024, t.vocab_size, t.codec_eos_id
        return sp

    # ---- code predictor frame (cousin :3243-3279 driver, :2669-2718 step) --------------
    def cp_frame(self, talker_hidden: torch.Tensor, code0: int, sp: Optional[SamplingParams] = None,
                 uniforms: Optional[Sequence[float]] = None, forced: Optional[Sequence[int]] = None,
                 return_logits: bool = False):
        """talker_hidden [H_talker] (post final norm), code0 -> codes[1..15], sum of the 16 embeddings."""
        w, c = self.w, self.cfg.cp
        sp = sp or SamplingParams()
        proj = lambda e: e @ w["cp.proj.weight"].T + w["cp.proj.bias"]
        self.cp.reset()
        e0 = w["talker.codec_embedding"][code0]
        acc = e0.clone()                                   # a8: sequential fp32 sum g = 0..15
        x = torch.stack([proj(talker_hidden), proj(e0)], 0)
        h = self.cp.forward(x)[-1]
        codes, all_logits = [], []
        for g in range(c.num_code_groups - 1):
            logits = h @ w[f"cp.heads.{g}.weight"].T
            all_logits.append(logits)
            s = process_logits(logits, sp, (), g)
            code = draw(s, sp, None if uniforms is None else uniforms[g])
            if forced is not None:
                code = int(forced[g])
            codes.append(code)
            e = w[f"cp.embeddings.{g}"][code]
            acc = acc + e
            if g < c.num_code_groups - 2:
                h = self.cp.forward(proj(e)[None])[-1]
        if return_logits:
            return codes, acc, torch.stack(all_logits)
        return codes, acc

    # ---- full frame loop ----------------------------------------------------------------
    def generate(self, prefill: torch.Tensor, trailing: torch.Tensor, max_frames: int,
                 talker_sp: Optional[SamplingParams] = None, cp_sp: Optional[SamplingParams] = None,
                 forced_codes: Optional[torch.Tensor] = None, record: bool = False, uniforms=None):
        """Returns codes [T, 16] (int64).  `forced_codes` [T,16] teacher-forces every sampled code while
        still recording what the oracle itself would have picked (used for teacher-forced parity).
        `uniforms(step, group) -> float in [0,1)` supplies the random number of every stochastic draw (RNG streams
        cannot match across frameworks: stochastic parity feeds both sides the same numbers)."""
        self.talker.reset()
        tsp = self.talker_sampling(talker_sp)
        h, logits = self.talker_forward(prefill)
        h, logits = h[-1], logits[-1]
        out, rec = [], {"talker_logits": [], "cp_logits": [], "own_codes": [], "margins": []}
        hist: List[int] = []
        for step in range(max_frames):
            s = process_logits(logits, tsp, hist, step)
            own0 = draw(s, tsp, None if uniforms is None else uniforms(step, 0))
            code0 = own0 if forced_codes is None else int(forced_codes[step, 0])
            if code0 == tsp.eos_id:
                break
            hist.append(code0)
            forced = None if forced_codes is None else [int(v) for v in forced_codes[step, 1:]]
            csp = cp_sp or SamplingParams()
            us = None if uniforms is None else [uniforms(step, g + 1) for g in range(self.cfg.cp.num_code_groups - 1)]
            rest, acc, cpl = self.cp_frame(h, code0, cp_sp, us, forced, True)
            if not csp.do_sample or us is not None:      # what the oracle itself picks from these logits (forced or not)
                own_rest = [draw(process_logits(cpl[g], csp, (), g), csp, None if us is None else us[g])
                            for g in range(cpl.shape[0])]
            else:
                own_rest = list(rest)
            out.append([code0] + list(rest))
            if record:
                top2 = torch.topk(s, 2).values
                rec.setdefault("talker_scores", []).append(s.clone())
                rec["talker_logits"].append(logits.clone())
                rec["cp_logits"].append(cpl)
                rec["own_codes"].append([own0] + list(own_rest))
                rec["margins"].append(float(top2[0] - top2[1]))
            nxt = acc + (trailing[step] if step < trailing.shape[0] - 1 else trailing[-1])
            h, logits = self.talker_forward(nxt[None])
            h, logits = h[-1], logits[-1]
        codes = torch.tensor(out, dtype=torch.long).reshape(-1, self.cfg.cp.num_code_groups)
        return (codes, rec) if record else codes


# ======================================================================================
# speech-tokenizer decoder (codes -> 24 kHz waveform)
# ======================================================================================
def rvq_decode(w: Dict[str, torch.Tensor], cfg, codes: torch.Tensor, split: bool = False):
    """codes [B, 16, T] -> [B, rvq_out_dim, T]  (mimi:1191-1195, 1216-1218, 1282-1293, 1340-1349).
    Gather + left-to-right fp32 sum is bit-exact work; the 1x1 projections are FP GEMMs."""
    k = cfg.codec
    outs, sums = [], []
    for grp, lo, hi in (("semantic", 0, k.num_semantic), ("acoustic", k.num_semantic, k.num_quantizers)):
        q = None
        for i in range(lo, hi):
            cb = w[f"codec.rvq.{grp}.codebooks.{i - lo}.embed_sum"] / \
                w[f"codec.rvq.{grp}.codebooks.{i - lo}.cluster_usage"].clamp(min=1e-5)[:, None]
            e = cb[codes[:, i].long()]                      # [B, T, D]
            q = e if q is None else q + e
        sums.append(q)
        outs.append(q @ w[f"codec.rvq.{grp}.out_proj.weight"].T)   # 1x1 conv, no bias
    y = (outs[0] + outs[1]).transpose(1, 2)
    return (y, sums) if split else y


def causal_conv1d(x, weight, bias, dilation=1, groups=1):
    """qwen3_omni_moe:3283-3316 (stride 1): left-pad (k-1)*dilation zeros."""
    ksz = weight.shape[-1]
    return F.conv1d(F.pad(x, ((ksz - 1) * dilation, 0)), weight, bias, dilation=dilation, groups=groups)


def causal_tconv1d(x, weight, bias, stride, trim="both"):
    """qwen3_omni_moe:3319-3331: ConvTranspose1d then trim (k - stride) samples."""
    ksz = weight.shape[-1]
    y = F.conv_transpose1d(x, weight, bias, stride=stride)
    pad = ksz - stride
    if pad == 0:
        return y
    if trim == "both":
        return y[..., pad: y.shape[-1] - pad]
    return y[..., : y.shape[-1] - pad]


def snake_beta(x, alpha, beta):
    """qwen3_omni_moe:3645-3683: x + sin^2(x e^alpha) / (e^beta + 1e-9)."""
    a = torch.exp(alpha)[None, :, None]
    b = torch.exp(beta)[None, :, None]
    return x + (1.0 / (b + 1e-9)) * torch.sin(x * a).pow(2)


def codec_transformer(w, cfg, x):
    """x [B, T, tf_hidden] -> [B, T, tf_hidden] (qwen3_omni_moe:3494-3642; sliding window 72 incl. self,
    masking_utils.py:90-99; no q/k norm :3396-3397; LayerScale :3479-3491)."""
    k = cfg.codec
    B, T, _ = x.shape
    pos = torch.arange(T)
    cos, sin = rope_cos_sin(pos, k.tf_head_dim, k.tf_rope_theta)
    i, j = pos[:, None], pos[None, :]
    mask = (j <= i) & (j > i - k.sliding_window)
    for l in range(k.tf_layers):
        p = f"codec.tf.layers.{l}"
        h = rms_norm(x, w[p + ".input_norm.weight"], k.tf_rms_eps)
        q = (h @ w[p + ".q_proj.weight"].T).view(B, T, k.tf_heads, k.tf_head_dim)
        kk = (h @ w[p + ".k_proj.weight"].T).view(B, T, k.tf_heads, k.tf_head_dim)
        v = (h @ w[p + ".v_proj.weight"].T).view(B, T, k.tf_heads, k.tf_head_dim)
        q = q * cos[None, :, None, :] + rotate_half(q) * sin[None, :, None, :]
        kk = kk * cos[None, :, None, :] + rotate_half(kk) * sin[None, :, None, :]
        sc = torch.einsum("bthd,bshd->bhts", q, kk) * (k.tf_head_dim ** -0.5)
        sc = sc.masked_fill(~mask[None, None], float("-inf"))
        a = torch.einsum("bhts,bshd->bthd", torch.softmax(sc, -1), v).reshape(B, T, -1)
        x = x + w[p + ".attn_scale"] * (a @ w[p + ".o_proj.weight"].T)
        h = rms_norm(x, w[p + ".post_norm.weight"], k.tf_rms_eps)
        m = (F.silu(h @ w[p + ".gate_proj.weight"].T) * (h @ w[p + ".up_proj.weight"].T)) @ w[p + ".down_proj.weight"].T
        x = x + w[p + ".mlp_scale"] * m
    return rms_norm(x, w["codec.tf.norm.weight"], k.tf_rms_eps)


def convnext_block(w, p: str, x: torch.Tensor) -> torch.Tensor:
    """qwen3_omni_moe:3334-3366.  x [B, C, T]."""
    y = causal_conv1d(x, w[p + ".dw.weight"], w[p + ".dw.bias"], groups=x.shape[1])
    y = F.layer_norm(y.transpose(1, 2), (x.shape[1],), w[p + ".ln.weight"], w[p + ".ln.bias"], 1e-6)
    y = F.gelu(y @ w[p + ".pw1.weight"].T + w[p + ".pw1.bias"])      # exact-erf GELU (:3345)
    y = y @ w[p + ".pw2.weight"].T + w[p + ".pw2.bias"]
    return x + (w[p + ".gamma"] * y).transpose(1, 2)


def decoder_block(w, cfg, p: str, x: torch.Tensor, rate: int) -> torch.Tensor:
    """qwen3_omni_moe:3686-3727: SnakeBeta -> transposed conv (k = 2r, stride r) -> 3 dilated residual units."""
    k = cfg.codec
    x = snake_beta(x, w[p + ".snake.alpha"], w[p + ".snake.beta"])
    x = causal_tconv1d(x, w[p + ".tconv.weight"], w[p + ".tconv.bias"], rate, k.transconv_trim)
    for j, d in enumerate((1, 3, 9)):
        u = f"{p}.units.{j}"
        y = snake_beta(x, w[u + ".snake1.alpha"], w[u + ".snake1.beta"])
        y = causal_conv1d(y, w[u + ".conv1.weight"], w[u + ".conv1.bias"], dilation=d)
        y = snake_beta(y, w[u + ".snake2.alpha"], w[u + ".snake2.beta"])
        y = causal_conv1d(y, w[u + ".conv2.weight"], w[u + ".conv2.bias"])
        x = x + y
    return x


def codec_forward(w, cfg, codes: torch.Tensor, stages: Optional[dict] = None) -> torch.Tensor:
    """One vocoder call: codes [B, 16, T] -> wav [B, 1, out_len(T)] (qwen3_omni_moe:3766-3778 with the
    Qwen3-TTS front end: split-RVQ decode -> causal conv k3 -> in-proj -> transformer -> out-proj)."""
    k = cfg.codec
    x = rvq_decode(w, cfg, codes)                                              # [B, 512, T]
    x = causal_conv1d(x, w["codec.pre_conv.weight"], w["codec.pre_conv.bias"])   # [B, 1024, T]
    if stages is not None:
        stages["pre_conv"] = x
    h = x.transpose(1, 2) @ w["codec.tf.in_proj.weight"].T + w["codec.tf.in_proj.bias"]
    h = codec_transformer(w, cfg, h)
    x = (h @ w["codec.tf.out_proj.weight"].T + w["codec.tf.out_proj.bias"]).transpose(1, 2)
    if stages is not None:
        stages["transformer"] = x
    for i, r in enumerate(k.upsampling_ratios):
        p = f"codec.up.{i}"
        x = causal_tconv1d(x, w[p + ".tconv.weight"], w[p + ".tconv.bias"], r, k.transconv_trim)
        x = convnext_block(w, p + ".cnx", x)
    if stages is not None:
        stages["upsample"] = x
    x = causal_conv1d(x, w["codec.dec.conv_in.weight"], w["codec.dec.conv_in.bias"])
    for i, r in enumerate(k.upsample_rates):
        x = decoder_block(w, cfg, f"codec.dec.blocks.{i}", x, r)
        if stages is not None:
            stages[f"block{i}"] = x
    x = snake_beta(x, w["codec.dec.snake_out.alpha"], w["codec.dec.snake_out.beta"])
    x = causal_conv1d(x, w["codec.dec.conv_out.weight"], w["codec.dec.conv_out.bias"])
    return x.clamp(-1, 1)


def codec_chunked_decode(w, cfg, codes: torch.Tensor) -> torch.Tensor:
    """qwen3_omni_moe:3780-3790: 300-frame chunks with 25 frames of left context, context samples dropped."""
    k = cfg.codec
    wavs, start, T = [], 0, codes.shape[-1]
    while start < T:
        end = min(start + k.chunk_size, T)
        ctx = k.left_context if start - k.left_context > 0 else start
        wav = codec_forward(w, cfg, codes[..., start - ctx:end])
        wavs.append(wav[..., ctx * k.hop:])
        start = end
    return torch.cat(wavs, -1)
