"""Seeded random-init weight factory + affine W8-g64 quantiser + HBM packing.

Checkpoints are not available offline, so parity runs on random-init weights of the named
architecture (BASELINE.json north_star).  The factory produces, for every quantised linear,
BOTH the uint8/scale/bias triple the device streams AND the de-quantised fp32 matrix the CPU
oracle multiplies with, so both sides compute on numerically identical weights
(quantise once on the host, never re-quantise on device).

Quantiser = MLX affine quantisation as used by the reference's checkpoints
(`mlx-community/*-8bit`, reference src/qwen3_tts/config.py:17,26,35; format in SURVEY
Appendix D): per output row, per group of 64 input elements, w ~= scale*q + bias, q in [0,255].
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from .config import ModelConfig, QUANT_GROUP

W8Triple = Tuple[torch.Tensor, torch.Tensor, torch.Tensor]  # q uint8 [N,K], scale bf16 [N,K/g], bias bf16 [N,K/g]


# --------------------------------------------------------------------------------------
# quantiser
# --------------------------------------------------------------------------------------
def quantize_w8(w: torch.Tensor, group: int = QUANT_GROUP) -> W8Triple:
    """Affine 8-bit group quantisation.  `w` is [N, K] float."""
    n, k = w.shape
    assert k % group == 0, (n, k)
    wg = w.float().reshape(n, k // group, group)
    w_max = wg.amax(-1)
    w_min = wg.amin(-1)
    eps = 1e-7
    scale = ((w_max - w_min) / 255.0).clamp_min(eps)
    side = w_min.abs() > w_max.abs()
    scale = torch.where(side, scale, -scale)
    edge = torch.where(side, w_min, w_max)
    q0 = torch.round(edge / scale)
    nz = q0 != 0
    scale = torch.where(nz, edge / torch.where(nz, q0, torch.ones_like(q0)), scale)
    bias = torch.where(nz, edge, torch.zeros_like(edge))
    # scales/biases are stored in the model dtype (bf16); quantise against the STORED values
    scale_b = scale.to(torch.bfloat16)
    bias_b = bias.to(torch.bfloat16)
    s = scale_b.float().unsqueeze(-1)
    b = bias_b.float().unsqueeze(-1)
    s_safe = torch.where(s == 0, torch.ones_like(s), s)
    q = torch.round((wg - b) / s_safe).clamp_(0, 255).to(torch.uint8)
    return q.reshape(n, k), scale_b, bias_b


def dequantize_w8(q: torch.Tensor, scale: torch.Tensor, bias: torch.Tensor, group: int = QUANT_GROUP) -> torch.Tensor:
    """fp32 matrix the oracle multiplies with: scale*q + bias evaluated in fp32."""
    n, k = q.shape
    qg = q.reshape(n, k // group, group).float()
    return (qg * scale.float().unsqueeze(-1) + bias.float().unsqueeze(-1)).reshape(n, k)


# --------------------------------------------------------------------------------------
# HBM layout of a W8 matrix: "fragment-ordered tiles"
#
# A tile = 16 output rows x 256 input columns (4 quantisation groups) = 4096 B of codes
# followed by 256 B of (scale, bias) metadata = 4352 contiguous bytes, tiles ordered
# [row_tile][k_chunk].  Inside a tile the codes are permuted so that ONE warp-wide 128-bit
# load (lane L reads bytes [16L, 16L+16)) yields exactly the four A-operand registers of an
# `mma.sync.m16n8k32.u8.s8` instruction: no shared-memory staging, no shuffles.
#   group j4 (0..3) , mma j (0..1)  ->  512 B block at (j4*2 + j) * 512
#   lane L: g = L>>2, t = L&3 ; register i (0..3) ; byte b (0..3)
#       row = g + 8*(i&1) ;  k = 64*j4 + 32*j + 16*(i>>1) + 4*t + b
#   byte offset inside block = 16*L + 4*i + b
# Metadata: 16 rows x [s0 s1 s2 s3 b0 b1 b2 b3] bf16 (16 B per row).
# --------------------------------------------------------------------------------------
TILE_ROWS = 16
TILE_K = 256
TILE_Q_BYTES = TILE_ROWS * TILE_K
TILE_META_BYTES = TILE_ROWS * 16
TILE_BYTES = TILE_Q_BYTES + TILE_META_BYTES


def _frag_index() -> torch.Tensor:
    """[4096] gather index: packed byte p of a tile <- source element (row*256 + k)."""
    idx = torch.empty(TILE_Q_BYTES, dtype=torch.long)
    p = torch.arange(TILE_Q_BYTES)
    blk = p // 512
    j4, j = blk // 2, blk % 2
    r = p % 512
    lane, i, b = r // 16, (r % 16) // 4, r % 4
    g, t = lane // 4, lane % 4
    row = g + 8 * (i & 1)
    k = 64 * j4 + 32 * j + 16 * (i >> 1) + 4 * t + b
    idx[:] = row * TILE_K + k
    return idx


_FRAG_IDX: Optional[torch.Tensor] = None


def pack_w8(q: torch.Tensor, scale: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """[N,K] uint8 + [N,K/64] bf16 x2  ->  flat uint8 blob of (N/16)*(K/256) tiles."""
    global _FRAG_IDX
    n, k = q.shape
    assert n % TILE_ROWS == 0 and k % TILE_K == 0, f"W8 matrix [{n},{k}] must be a multiple of 16x256"
    if _FRAG_IDX is None:
        _FRAG_IDX = _frag_index()
    idx = _FRAG_IDX.to(q.device)
    nrt, nkc = n // TILE_ROWS, k // TILE_K
    tiles = q.reshape(nrt, TILE_ROWS, nkc, TILE_K).permute(0, 2, 1, 3).reshape(nrt, nkc, TILE_Q_BYTES)
    qp = tiles[:, :, idx]                                                # [nrt, nkc, 4096]
    s = scale.reshape(nrt, TILE_ROWS, nkc, 4).permute(0, 2, 1, 3)        # [nrt, nkc, 16, 4]
    b = bias.reshape(nrt, TILE_ROWS, nkc, 4).permute(0, 2, 1, 3)
    meta = torch.cat([s, b], dim=-1).contiguous().view(torch.uint8).reshape(nrt, nkc, TILE_META_BYTES)
    return torch.cat([qp, meta], dim=-1).reshape(-1).contiguous()


def unpack_w8(blob: torch.Tensor, n: int, k: int) -> W8Triple:
    """Inverse of pack_w8 (host-side check of the layout)."""
    global _FRAG_IDX
    if _FRAG_IDX is None:
        _FRAG_IDX = _frag_index()
    nrt, nkc = n // TILE_ROWS, k // TILE_K
    t = blob.reshape(nrt, nkc, TILE_BYTES)
    qp, meta = t[..., :TILE_Q_BYTES], t[..., TILE_Q_BYTES:]
    tiles = torch.empty_like(qp)
    tiles[:, :, _FRAG_IDX.to(blob.device)] = qp
    q = tiles.reshape(nrt, nkc, TILE_ROWS, TILE_K).permute(0, 2, 1, 3).reshape(n, k)
    m = meta.contiguous().view(torch.bfloat16).reshape(nrt, nkc, TILE_ROWS, 8)
    s = m[..., :4].permute(0, 2, 1, 3).reshape(n, k // 64)
    b = m[..., 4:].permute(0, 2, 1, 3).reshape(n, k // 64)
    return q.contiguous(), s.contiguous(), b.contiguous()


# --------------------------------------------------------------------------------------
# factory
# --------------------------------------------------------------------------------------
class WeightStore:
    """name -> tensor.  `fp[name]` is fp32 (de-quantised for W8 linears); `q[name]` the W8 triple."""

    def __init__(self, cfg: ModelConfig):
        self.cfg = cfg
        self.fp: Dict[str, torch.Tensor] = {}
        self.q: Dict[str, W8Triple] = {}

    def __getitem__(self, name: str) -> torch.Tensor:
        return self.fp[name]


def _stack_layers(add_lin, add_norm, prefix: str, n_layers: int, hidden: int, q_dim: int, kv_dim: int,
                  head_dim: int, inter: int):
    for i in range(n_layers):
        p = f"{prefix}.layers.{i}"
        add_norm(f"{p}.input_norm", hidden)
        add_lin(f"{p}.q_proj", q_dim, hidden)
        add_lin(f"{p}.k_proj", kv_dim, hidden)
        add_lin(f"{p}.v_proj", kv_dim, hidden)
        add_norm(f"{p}.q_norm", head_dim)
        add_norm(f"{p}.k_norm", head_dim)
        add_lin(f"{p}.o_proj", hidden, q_dim)
        add_norm(f"{p}.post_norm", hidden)
        add_lin(f"{p}.gate_proj", inter, hidden)
        add_lin(f"{p}.up_proj", inter, hidden)
        add_lin(f"{p}.down_proj", hidden, inter)
    add_norm(f"{prefix}.norm", hidden)


def expected_shapes(cfg: ModelConfig) -> "WeightStore":
    """The store of `cfg` with every tensor on the META device: names, shapes and which linears are W8, nothing
    materialised (the checkpoint loader checks shapes and conv layouts against it)."""
    parts = ("talker", "cp", "codec") + (("enc", "spk") if cfg.tts_model_type == "base" else ())
    return make_weights(cfg, device="meta", keep_fp=True, keep_q=True, parts=parts)


def make_weights(cfg: ModelConfig, seed: int = 0, device: str = "cpu", keep_fp: bool = True,
                 keep_q: bool = True, head_std: float = 0.02, parts=("talker", "cp", "codec")) -> WeightStore:
    """Seeded N(0, 0.02^2) init of every tensor on the hot path (SURVEY 8d "Synthetic inputs").

    Norm weights are drawn around 1 (not exactly 1) and SnakeBeta alpha/beta ~ N(0, 0.1^2), LayerScale
    0.01, ConvNeXt gamma 1 so that every branch of the arithmetic is exercised by the parity tests.
    """
    ws = WeightStore(cfg)
    meta = str(device) == "meta"
    gen = None if meta else torch.Generator(device=device)
    if gen is not None:
        gen.manual_seed(seed)

    def randn(*shape, std=0.02):
        if meta:
            return torch.empty(*shape, device="meta", dtype=torch.float32)
        return torch.randn(*shape, generator=gen, device=device, dtype=torch.float32) * std

    def add_lin(name, n, k, quant=True, std=0.02, bias=False):
        w = randn(n, k, std=std)
        if meta and quant and k % cfg.quant_group == 0:
            g = cfg.quant_group
            ws.q[name] = (torch.empty(n, k, device="meta", dtype=torch.uint8), torch.empty(n, k // g, device="meta", dtype=torch.bfloat16),
                          torch.empty(n, k // g, device="meta", dtype=torch.bfloat16))
            ws.fp[name + ".weight"] = w
        elif quant and k % cfg.quant_group == 0:
            trip = quantize_w8(w, cfg.quant_group)
            if keep_q:
                ws.q[name] = trip
            if keep_fp:
                ws.fp[name + ".weight"] = dequantize_w8(*trip, cfg.quant_group)
        else:
            ws.fp[name + ".weight"] = w
        if bias:
            ws.fp[name + ".bias"] = randn(n, std=0.02)

    def add_norm(name, n):
        ws.fp[name + ".weight"] = 1.0 + randn(n, std=0.05)

    t, c, k = cfg.talker, cfg.cp, cfg.codec
    if "talker" in parts:
        ws.fp["talker.text_embedding"] = randn(t.text_vocab_size, t.text_hidden_size)
        add_lin("talker.text_projection.fc1", t.text_hidden_size, t.text_hidden_size, bias=True)
        add_lin("talker.text_projection.fc2", t.hidden_size, t.text_hidden_size, bias=True)
        ws.fp["talker.codec_embedding"] = randn(t.vocab_size, t.hidden_size)
        _stack_layers(add_lin, add_norm, "talker", t.num_layers, t.hidden_size, t.q_dim, t.kv_dim, t.head_dim,
                      t.intermediate_size)
        add_lin("talker.codec_head", t.vocab_size, t.hidden_size, std=head_std)
    if "cp" in parts:
        add_lin("cp.proj", c.hidden_size, c.embed_dim, bias=True)
        for g in range(c.num_code_groups - 1):
            ws.fp[f"cp.embeddings.{g}"] = randn(c.vocab_size, c.embed_dim)
        _stack_layers(add_lin, add_norm, "cp", c.num_layers, c.hidden_size, c.q_dim, c.kv_dim, c.head_dim,
                      c.intermediate_size)
        for g in range(c.num_code_groups - 1):
            add_lin(f"cp.heads.{g}", c.vocab_size, c.hidden_size, std=head_std)
    if "codec" in parts:
        _make_codec(ws, cfg, randn)
    if "enc" in parts:
        _make_speech_encoder(ws, cfg, randn)
    if "spk" in parts:
        _make_speaker_encoder(ws, cfg, randn)
    return ws


def _make_speech_encoder(ws: WeightStore, cfg: ModelConfig, randn):
    """Speech-tokenizer encoder (Mimi layout, transformers mimi:454-496, 926-1141, 1296-1340), fp32."""
    e, fp = cfg.enc, ws.fp

    def conv(name, cout, cin, ksz, bias=True):
        fp[name + ".weight"] = randn(cout, cin, ksz, std=(1.0 / (cin * ksz)) ** 0.5)
        if bias:
            fp[name + ".bias"] = randn(cout, std=0.02)

    def lin(name, n, kk):
        fp[name + ".weight"] = randn(n, kk, std=(1.0 / kk) ** 0.5)

    conv("enc.conv_in", e.num_filters, 1, e.kernel_size)
    dim = e.num_filters
    for i, r in enumerate(reversed(e.ratios)):
        p = f"enc.stages.{i}"
        conv(p + ".res.conv1", dim // e.compress, dim, e.residual_kernel_size)
        conv(p + ".res.conv2", dim, dim // e.compress, 1)
        conv(p + ".down", 2 * dim, dim, 2 * r)
        dim *= 2
    conv("enc.conv_out", e.hidden_size, dim, e.last_kernel_size)
    hd = e.tf_heads * e.tf_head_dim
    for l in range(e.tf_layers):
        p = f"enc.tf.layers.{l}"
        for nm in ("input_norm", "post_norm"):
            fp[f"{p}.{nm}.weight"] = 1.0 + randn(e.hidden_size, std=0.05)
            fp[f"{p}.{nm}.bias"] = randn(e.hidden_size, std=0.02)
        lin(p + ".q_proj", hd, e.hidden_size); lin(p + ".k_proj", hd, e.hidden_size); lin(p + ".v_proj", hd, e.hidden_size)
        lin(p + ".o_proj", e.hidden_size, hd)
        lin(p + ".fc1", e.tf_intermediate, e.hidden_size); lin(p + ".fc2", e.hidden_size, e.tf_intermediate)
        for nm in ("attn_scale", "mlp_scale"):
            # LayerScale is 0.01 at init (mimi:499-512); 0.3 here so that the branches matter in parity tests
            fp[f"{p}.{nm}"] = 0.3 * (1.0 + randn(e.hidden_size, std=0.05))
    conv("enc.downsample", e.hidden_size, e.hidden_size, 4, bias=False)
    for grp, nq in (("semantic", e.num_semantic), ("acoustic", e.num_quantizers - e.num_semantic)):
        lin(f"enc.rvq.{grp}.in_proj", e.codebook_dim, e.hidden_size)
        for i in range(nq):
            # residual levels shrink: later codebooks are drawn smaller, as trained RVQ codebooks are
            fp[f"enc.rvq.{grp}.codebooks.{i}.embed_sum"] = randn(e.codebook_size, e.codebook_dim, std=1.0 * (0.7 ** i))
            fp[f"enc.rvq.{grp}.codebooks.{i}.cluster_usage"] = randn(e.codebook_size, std=0.0) + 1.0


def _make_speaker_encoder(ws: WeightStore, cfg: ModelConfig, randn):
    """ECAPA-TDNN (transformers qwen2_5_omni:2499-2790), fp32."""
    sc, fp = cfg.spk, ws.fp

    def conv(name, cout, cin, ksz):
        fp[name + ".weight"] = randn(cout, cin, ksz, std=(1.0 / (cin * ksz)) ** 0.5)
        fp[name + ".bias"] = randn(cout, std=0.05)

    ch = sc.channels
    conv("spk.blocks.0.conv", ch[0], sc.n_mels, sc.kernel_sizes[0])
    for i in range(1, len(ch) - 1):
        p = f"spk.blocks.{i}"
        conv(p + ".tdnn1.conv", ch[i], ch[i - 1], 1)
        w = ch[i] // sc.res2net_scale
        for j in range(sc.res2net_scale - 1):
            conv(f"{p}.res2net.{j}.conv", w, w, sc.kernel_sizes[i])
        conv(p + ".tdnn2.conv", ch[i], ch[i], 1)
        conv(p + ".se.conv1", sc.se_channels, ch[i], 1)
        conv(p + ".se.conv2", ch[i], sc.se_channels, 1)
    conv("spk.mfa.conv", ch[-1], ch[-1], sc.kernel_sizes[-1])
    conv("spk.asp.tdnn.conv", sc.attention_channels, 3 * ch[-1], 1)
    conv("spk.asp.conv", ch[-1], sc.attention_channels, 1)
    conv("spk.fc", sc.enc_dim, 2 * ch[-1], 1)


def _make_codec(ws: WeightStore, cfg: ModelConfig, randn):
    k = cfg.codec
    fp = ws.fp

    def conv(name, cout, cin, ksz, std=None):
        std = std if std is not None else (1.0 / (cin * ksz)) ** 0.5
        fp[name + ".weight"] = randn(cout, cin, ksz, std=std)
        fp[name + ".bias"] = randn(cout, std=0.02)

    def tconv(name, cin, cout, ksz, gain=1.0):
        # nn.ConvTranspose1d layout [Cin, Cout, k]; two taps overlap per output sample
        fp[name + ".weight"] = randn(cin, cout, ksz, std=gain * (1.0 / (2 * cin)) ** 0.5)
        fp[name + ".bias"] = randn(cout, std=0.02)

    def lin(name, n, kk, bias=False, std=None):
        fp[name + ".weight"] = randn(n, kk, std=std if std is not None else (1.0 / kk) ** 0.5)
        if bias:
            fp[name + ".bias"] = randn(n, std=0.02)

    def norm(name, n, bias=False):
        fp[name + ".weight"] = 1.0 + randn(n, std=0.05)
        if bias:
            fp[name + ".bias"] = randn(n, std=0.02)

    def snake(name, n):
        fp[name + ".alpha"] = randn(n, std=0.1)
        fp[name + ".beta"] = randn(n, std=0.1)

    # split RVQ: 1 semantic + 15 acoustic codebooks, one 1x1 out-projection each (mimi:1252-1260)
    for grp, nq in (("semantic", k.num_semantic), ("acoustic", k.num_quantizers - k.num_semantic)):
        for i in range(nq):
            fp[f"codec.rvq.{grp}.codebooks.{i}.embed_sum"] = randn(k.codebook_size, k.codebook_dim, std=1.0)
            fp[f"codec.rvq.{grp}.codebooks.{i}.cluster_usage"] = torch.ones(k.codebook_size, device=fp[
                f"codec.rvq.{grp}.codebooks.{i}.embed_sum"].device)
        lin(f"codec.rvq.{grp}.out_proj", k.rvq_out_dim, k.codebook_dim)
    conv("codec.pre_conv", k.latent_dim, k.rvq_out_dim, 3)
    lin("codec.tf.in_proj", k.tf_hidden, k.latent_dim, bias=True)
    hd = k.tf_heads * k.tf_head_dim
    for i in range(k.tf_layers):
        p = f"codec.tf.layers.{i}"
        norm(p + ".input_norm", k.tf_hidden)
        lin(p + ".q_proj", hd, k.tf_hidden)
        lin(p + ".k_proj", hd, k.tf_hidden)
        lin(p + ".v_proj", hd, k.tf_hidden)
        lin(p + ".o_proj", k.tf_hidden, hd)
        fp[p + ".attn_scale"] = torch.full((k.tf_hidden,), k.layer_scale, device=fp[p + ".o_proj.weight"].device) \
            * (1.0 + randn(k.tf_hidden, std=0.05))
        norm(p + ".post_norm", k.tf_hidden)
        lin(p + ".gate_proj", k.tf_intermediate, k.tf_hidden)
        lin(p + ".up_proj", k.tf_intermediate, k.tf_hidden)
        lin(p + ".down_proj", k.tf_hidden, k.tf_intermediate)
        fp[p + ".mlp_scale"] = torch.full((k.tf_hidden,), k.layer_scale, device=fp[p + ".o_proj.weight"].device) \
            * (1.0 + randn(k.tf_hidden, std=0.05))
    norm("codec.tf.norm", k.tf_hidden)
    lin("codec.tf.out_proj", k.latent_dim, k.tf_hidden, bias=True)
    for i, r in enumerate(k.upsampling_ratios):
        p = f"codec.up.{i}"
        tconv(p + ".tconv", k.latent_dim, k.latent_dim, r)
        fp[p + ".cnx.dw.weight"] = randn(k.latent_dim, 1, 7, std=(1.0 / 7) ** 0.5)
        fp[p + ".cnx.dw.bias"] = randn(k.latent_dim, std=0.02)
        norm(p + ".cnx.ln", k.latent_dim, bias=True)
        lin(p + ".cnx.pw1", 4 * k.latent_dim, k.latent_dim, bias=True)
        lin(p + ".cnx.pw2", k.latent_dim, 4 * k.latent_dim, bias=True)
        fp[p + ".cnx.gamma"] = 1.0 + randn(k.latent_dim, std=0.05)
    conv("codec.dec.conv_in", k.decoder_dim, k.latent_dim, 7)
    ch = k.decoder_dim
    for i, r in enumerate(k.upsample_rates):
        p = f"codec.dec.blocks.{i}"
        snake(p + ".snake", ch)
        tconv(p + ".tconv", ch, ch // 2, 2 * r, gain=0.6)
        ch //= 2
        for j in range(3):
            u = f"{p}.units.{j}"
            snake(u + ".snake1", ch)
            conv(u + ".conv1", ch, ch, 7)
            snake(u + ".snake2", ch)
            conv(u + ".conv2", ch, ch, 1, std=0.3 * (1.0 / ch) ** 0.5)
    snake("codec.dec.snake_out", ch)
    conv("codec.dec.conv_out", 1, ch, 7, std=0.08 * (1.0 / (7 * ch)) ** 0.5)
