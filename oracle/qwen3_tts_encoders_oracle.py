"""CPU ORACLE (reference-clip side) -- TEST INFRASTRUCTURE ONLY.  Never imported by the product path.

PyTorch-CPU fp32 restatement of what the reference's voice-cloning call (`generate_audio(..., ref_audio=<wav>, ref_text=)`,
/root/reference/src/qwen3_tts/sessions/clone.py:218-224) makes its un-vendored dependency (`mlx-audio==0.3.1`,
/root/reference/pyproject.toml:39) compute from the reference clip before the frame loop starts:
  * speech-tokenizer ENCODER: 24 kHz wav -> [16, T] RVQ codes at 12.5 Hz.  Qwen3-TTS-Tokenizer-12Hz encodes with a Mimi model
    (SEANet conv encoder -> 8-layer causal transformer -> stride-2 downsample -> split residual VQ, first 16 quantizers);
  * speaker encoder: log-mel spectrogram (128 bins) -> ECAPA-TDNN -> one vector of the talker's hidden size.

PARITY UNPINNED by the reference (no vectors, no fixtures for this path; neither mlx-audio nor the upstream QwenLM package
installs offline).  Pinned instead against the structurally identical classes on disk in transformers 5.5.0
(tests/test_oracle_vs_cousins.py):
    T = transformers/models
    MimiConv1d padding rule          T/mimi/modeling_mimi.py:214-352
    SEANet encoder + resnet block    T/mimi/modeling_mimi.py:412-496
    transformer layer (LayerNorm, LayerScale, GELU MLP, RoPE, causal sliding window 250)   T/mimi/modeling_mimi.py:499-995, masking_utils.py:90-99
    RVQ encode (cdist argmin, residual) + split quantizer   T/mimi/modeling_mimi.py:1176-1340
    encode pipeline                   T/mimi/modeling_mimi.py:1455-1484
    ECAPA-TDNN                        T/qwen2_5_omni/modeling_qwen2_5_omni.py:2499-2790
    mel filter bank (slaney)          transformers/audio_utils.py mel_filter_bank
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F


# ======================================================================================================================
# speech-tokenizer encoder (Mimi)
# ======================================================================================================================
def mimi_conv1d(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], stride: int = 1, dilation: int = 1,
                pad_mode: str = "constant") -> torch.Tensor:
    """mimi:273-352 (causal): left pad = effective kernel - stride, right pad = what completes the last frame."""
    k = (weight.shape[-1] - 1) * dilation + 1
    pad_total = k - stride
    length = x.shape[-1]
    n_frames = math.ceil((length - k + pad_total) / stride + 1) - 1
    extra = n_frames * stride + k - pad_total - length
    x = F.pad(x, (pad_total, extra), mode=pad_mode)
    return F.conv1d(x, weight, bias, stride=stride, dilation=dilation)


def seanet_encoder(w: Dict[str, torch.Tensor], ecfg, x: torch.Tensor, stages: Optional[dict] = None) -> torch.Tensor:
    """mimi:454-496.  x [B, 1, n] -> [B, hidden, n / prod(ratios)]."""
    x = mimi_conv1d(x, w["enc.conv_in.weight"], w["enc.conv_in.bias"])
    for i, r in enumerate(reversed(ecfg.ratios)):
        p = f"enc.stages.{i}"
        y = F.elu(x)                                                                         # resnet block, mimi:412-451
        y = mimi_conv1d(y, w[p + ".res.conv1.weight"], w[p + ".res.conv1.bias"])
        y = F.elu(y)
        y = mimi_conv1d(y, w[p + ".res.conv2.weight"], w[p + ".res.conv2.bias"])
        x = x + y
        x = mimi_conv1d(F.elu(x), w[p + ".down.weight"], w[p + ".down.bias"], stride=r)
        if stages is not None:
            stages[f"stage{i}"] = x
    return mimi_conv1d(F.elu(x), w["enc.conv_out.weight"], w["enc.conv_out.bias"])


def _rope(x: torch.Tensor, theta: float) -> torch.Tensor:
    """x [B, T, H, D], rotate_half convention, fp32 cos/sin (mimi:515-612)."""
    T, D = x.shape[1], x.shape[-1]
    inv = 1.0 / (theta ** (torch.arange(0, D, 2, dtype=torch.float32) / D))
    fr = torch.arange(T, dtype=torch.float32)[:, None] * inv[None]
    emb = torch.cat([fr, fr], -1)
    cos, sin = emb.cos()[None, :, None], emb.sin()[None, :, None]
    h = D // 2
    return x * cos + torch.cat([-x[..., h:], x[..., :h]], -1) * sin


def mimi_transformer(w: Dict[str, torch.Tensor], ecfg, x: torch.Tensor) -> torch.Tensor:
    """mimi:926-1141.  x [B, T, hidden]; pre-LN layers with LayerScale, no final norm."""
    B, T, _ = x.shape
    H, D = ecfg.tf_heads, ecfg.tf_head_dim
    i, j = torch.arange(T)[:, None], torch.arange(T)[None, :]
    mask = (j <= i) & (j > i - ecfg.sliding_window)
    for l in range(ecfg.tf_layers):
        p = f"enc.tf.layers.{l}"
        h = F.layer_norm(x, (x.shape[-1],), w[p + ".input_norm.weight"], w[p + ".input_norm.bias"], ecfg.norm_eps)
        q = _rope((h @ w[p + ".q_proj.weight"].T).view(B, T, H, D), ecfg.rope_theta)
        k = _rope((h @ w[p + ".k_proj.weight"].T).view(B, T, H, D), ecfg.rope_theta)
        v = (h @ w[p + ".v_proj.weight"].T).view(B, T, H, D)
        sc = torch.einsum("bthd,bshd->bhts", q, k) * (D ** -0.5)
        sc = sc.masked_fill(~mask[None, None], float("-inf"))
        a = torch.einsum("bhts,bshd->bthd", torch.softmax(sc, -1), v).reshape(B, T, H * D)
        x = x + w[p + ".attn_scale"] * (a @ w[p + ".o_proj.weight"].T)
        h = F.layer_norm(x, (x.shape[-1],), w[p + ".post_norm.weight"], w[p + ".post_norm.bias"], ecfg.norm_eps)
        x = x + w[p + ".mlp_scale"] * (F.gelu(h @ w[p + ".fc1.weight"].T) @ w[p + ".fc2.weight"].T)
    return x


def rvq_encode(w: Dict[str, torch.Tensor], ecfg, emb: torch.Tensor, n_q: int, record: Optional[dict] = None) -> torch.Tensor:
    """mimi:1197-1203 (cdist argmin), 1262-1281 (residual), 1311-1340 (split: both groups quantise the SAME embeddings).
    emb [B, hidden, T] -> codes [B, n_q, T] int64.  `record` collects the per-level residuals and top-2 distance gaps."""
    out = []
    for grp, lo, hi in (("semantic", 0, ecfg.num_semantic), ("acoustic", ecfg.num_semantic, n_q)):
        res = torch.einsum("oc,bct->bto", w[f"enc.rvq.{grp}.in_proj.weight"], emb)             # 1x1 conv, no bias
        for i in range(hi - lo):
            cb = w[f"enc.rvq.{grp}.codebooks.{i}.embed_sum"] / w[f"enc.rvq.{grp}.codebooks.{i}.cluster_usage"].clamp(min=1e-5)[:, None]
            flat = res.reshape(-1, res.shape[-1])
            d = torch.cdist(flat[None].float(), cb[None].float(), p=2)[0]
            idx = d.argmin(-1)
            if record is not None:
                top2 = torch.topk(d, 2, dim=-1, largest=False).values
                record.setdefault("residual", []).append(res.clone())
                record.setdefault("gap", []).append((top2[:, 1] - top2[:, 0]).view(res.shape[:-1]))
                record.setdefault("dist0", []).append(top2[:, 0].view(res.shape[:-1]))
            out.append(idx.view(res.shape[:-1]))
            res = res - cb[idx].view_as(res)
    return torch.stack(out, 1)


def speech_encode(w: Dict[str, torch.Tensor], ecfg, wav: torch.Tensor, record: Optional[dict] = None) -> torch.Tensor:
    """mimi:1455-1484.  wav [B, n] at 24 kHz -> codes [B, valid_quantizers, T]."""
    x = seanet_encoder(w, ecfg, wav[:, None], record)
    if record is not None:
        record["seanet"] = x
    x = mimi_transformer(w, ecfg, x.transpose(1, 2)).transpose(1, 2)
    if record is not None:
        record["transformer"] = x
    x = mimi_conv1d(x, w["enc.downsample.weight"], None, stride=2, pad_mode="replicate")
    if record is not None:
        record["embeddings"] = x
    return rvq_encode(w, ecfg, x, ecfg.valid_quantizers, record)


# ======================================================================================================================
# speaker encoder: log-mel front end + ECAPA-TDNN
# ======================================================================================================================
def hz_to_mel_slaney(f):
    f = torch.as_tensor(f, dtype=torch.float64)
    lin = 3.0 * f / 200.0
    log = 15.0 + torch.log(f.clamp(min=1e-10) / 1000.0) * (27.0 / math.log(6.4))
    return torch.where(f >= 1000.0, log, lin)


def mel_to_hz_slaney(m):
    m = torch.as_tensor(m, dtype=torch.float64)
    lin = 200.0 * m / 3.0
    log = 1000.0 * torch.exp(math.log(6.4) / 27.0 * (m - 15.0))
    return torch.where(m >= 15.0, log, lin)


def mel_filter_bank(n_fft: int, n_mels: int, sr: int, fmin: float, fmax: float) -> torch.Tensor:
    """librosa.filters.mel(htk=False, norm='slaney') == transformers.audio_utils.mel_filter_bank(norm='slaney',
    mel_scale='slaney').  Returns [n_fft/2+1, n_mels] float32."""
    n_freq = n_fft // 2 + 1
    fft_freqs = torch.linspace(0, sr / 2, n_freq, dtype=torch.float64)
    mel_pts = torch.linspace(float(hz_to_mel_slaney(fmin)), float(hz_to_mel_slaney(fmax)), n_mels + 2, dtype=torch.float64)
    hz = mel_to_hz_slaney(mel_pts)
    fdiff = hz[1:] - hz[:-1]
    slopes = hz[None, :] - fft_freqs[:, None]
    down = -slopes[:, :-2] / fdiff[:-1]
    up = slopes[:, 2:] / fdiff[1:]
    fb = torch.clamp(torch.minimum(down, up), min=0.0)
    fb = fb * (2.0 / (hz[2:n_mels + 2] - hz[:n_mels]))[None]
    return fb.float()


def log_mel(wav: torch.Tensor, scfg) -> torch.Tensor:
    """BigVGAN / HiFi-GAN style mel front end [U]: reflect-pad (n_fft - hop)/2, STFT (hann, center=False), magnitude
    sqrt(re^2 + im^2 + 1e-9), slaney mel filter bank, log(clamp(., 1e-5)).  wav [B, n] -> [B, frames, n_mels]."""
    pad = (scfg.n_fft - scfg.hop) // 2
    y = F.pad(wav[:, None], (pad, pad), mode="reflect")[:, 0]
    spec = torch.stft(y, scfg.n_fft, hop_length=scfg.hop, win_length=scfg.win, window=torch.hann_window(scfg.win),
                      center=False, normalized=False, onesided=True, return_complex=True)
    mag = torch.sqrt(spec.real.pow(2) + spec.imag.pow(2) + 1e-9)                              # [B, n_freq, frames]
    mel = torch.einsum("fm,bft->btm", mel_filter_bank(scfg.n_fft, scfg.n_mels, scfg.sample_rate, scfg.fmin, scfg.fmax), mag)
    return torch.log(mel.clamp(min=1e-5))


def _tdnn(w, name: str, x: torch.Tensor, dilation: int = 1, relu: bool = True) -> torch.Tensor:
    """qwen2_5_omni:2499-2521: Conv1d(padding='same', padding_mode='reflect') + ReLU.  x [B, C, T]."""
    weight, bias = w[name + ".weight"], w[name + ".bias"]
    k = weight.shape[-1]
    total = dilation * (k - 1)
    left = total // 2
    if total:
        x = F.pad(x, (left, total - left), mode="reflect")
    y = F.conv1d(x, weight, bias, dilation=dilation)
    return F.relu(y) if relu else y


def ecapa_forward(w: Dict[str, torch.Tensor], scfg, mel: torch.Tensor, stages: Optional[dict] = None) -> torch.Tensor:
    """qwen2_5_omni:2717-2790.  mel [B, frames, n_mels] -> [B, enc_dim]."""
    x = mel.transpose(1, 2)
    feats: List[torch.Tensor] = []
    x = _tdnn(w, "spk.blocks.0.conv", x, scfg.dilations[0])
    for i in range(1, len(scfg.channels) - 1):
        p = f"spk.blocks.{i}"
        res = x
        x = _tdnn(w, p + ".tdnn1.conv", x)
        parts, prev = [], None
        for j, part in enumerate(torch.chunk(x, scfg.res2net_scale, dim=1)):               # Res2Net, :2524-2560
            if j == 0:
                prev = part
            elif j == 1:
                prev = _tdnn(w, f"{p}.res2net.{j - 1}.conv", part, scfg.dilations[i])
            else:
                prev = _tdnn(w, f"{p}.res2net.{j - 1}.conv", part + prev, scfg.dilations[i])
            parts.append(prev)
        x = _tdnn(w, p + ".tdnn2.conv", torch.cat(parts, 1))
        s = x.mean(2, keepdim=True)                                                          # squeeze-excitation, :2563-2596
        s = F.relu(F.conv1d(s, w[p + ".se.conv1.weight"], w[p + ".se.conv1.bias"]))
        s = torch.sigmoid(F.conv1d(s, w[p + ".se.conv2.weight"], w[p + ".se.conv2.bias"]))
        x = x * s + res
        feats.append(x)
        if stages is not None:
            stages[f"block{i}"] = x
    x = _tdnn(w, "spk.mfa.conv", torch.cat(feats, 1), scfg.dilations[-1])
    if stages is not None:
        stages["mfa"] = x
    # attentive statistics pooling, :2599-2679 (single full-length sequence per item)
    T = x.shape[-1]
    mean = x.mean(2)
    std = torch.sqrt(((x - mean[..., None]).pow(2).mean(2)).clamp(1e-12))
    att_in = torch.cat([x, mean[..., None].expand(-1, -1, T), std[..., None].expand(-1, -1, T)], 1)
    att = _tdnn(w, "spk.asp.tdnn.conv", att_in)
    att = F.conv1d(torch.tanh(att), w["spk.asp.conv.weight"], w["spk.asp.conv.bias"])
    att = torch.softmax(att, dim=2)
    mean = (att * x).sum(2)
    std = torch.sqrt((att * (x - mean[..., None]).pow(2)).sum(2).clamp(1e-12))
    pooled = torch.cat([mean, std], 1)[..., None]
    return F.conv1d(pooled, w["spk.fc.weight"], w["spk.fc.bias"])[..., 0]


def speaker_embed(w: Dict[str, torch.Tensor], scfg, wav: torch.Tensor) -> torch.Tensor:
    return ecapa_forward(w, scfg, log_mel(wav, scfg))
