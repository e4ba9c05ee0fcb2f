"""Host side of the reference-clip encoders (voice cloning, reference call site sessions/clone.py:218-224):
`SpeechEncoder`  24 kHz wav -> [16, T] RVQ codes (Mimi: SEANet conv stem, 8-layer causal transformer, stride-2 downsample,
                 split residual VQ; SURVEY 8f-2, cousin transformers mimi/modeling_mimi.py:454-496, 926-1141, 1296-1340, 1455-1484)
`SpeakerEncoder` 24 kHz wav -> log-mel -> ECAPA-TDNN -> one vector of the talker width (cousin
                 transformers qwen2_5_omni/modeling_qwen2_5_omni.py:2499-2790).
Every convolution / linear is a q3t_tapgemm call (FP32 pipe: a nearest-codebook search follows, so no TF32 rounding); a strided
convolution is a 2-tap GEMM over the input viewed as [T / stride, stride * C]; the STFT is a 4-tap GEMM against the windowed
DFT basis.  The remaining operators are kernels of csrc/encoders.cu.  PyTorch only pads, reshapes and concatenates here.
Runs once per reference clip; nothing of it is on the per-frame path.
"""
from __future__ import annotations

import ctypes as C
import math
import wave
from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from . import lib as L
from .config import ModelConfig
from .weights import WeightStore


def read_wav(path: str, target_sr: int) -> torch.Tensor:
    """Mono float32 [-1, 1] at `target_sr`.  The reference converts every clip to 24 kHz mono PCM16 before the call
    (io.py:231-286), so anything else is rejected rather than resampled behind the caller's back."""
    with wave.open(path, "rb") as w:
        sr, nch, sw, n = w.getframerate(), w.getnchannels(), w.getsampwidth(), w.getnframes()
        raw = w.readframes(n)
    if sw != 2:
        raise ValueError(f"{path}: expected 16-bit PCM (the reference converts clips with ffmpeg/afconvert first), got {8 * sw}-bit")
    x = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    if nch > 1:
        x = x.reshape(-1, nch).mean(1)
    if sr != target_sr:
        raise ValueError(f"{path}: sample rate {sr} != {target_sr} (reference clips are converted to 24 kHz mono, io.py:243-247)")
    return torch.from_numpy(x.copy())


class _Ops:
    """tap-GEMM + small operator wrappers shared by both encoders (fp32, time-major [B, T, C])."""

    def __init__(self, device):
        self.lib, self.dev = L.load(), torch.device(device)

    def tap(self, A: torch.Tensor, W: torch.Tensor, bias, shifts: List[int], rows: int, act: int = L.ACT_NONE, want_raw: bool = True,
            scale=None, resid=None):
        """A [B, T, Cin]; W [taps, N, Cin] -> (raw [B, rows, N] | None, act(raw) | None)."""
        B, T, Cin = A.shape
        assert A.is_contiguous() and W.is_contiguous() and W.shape[2] == Cin and Cin % 4 == 0
        N = W.shape[1]
        a = L.TapGemmArgs()
        a.A, a.B, a.T_in, a.Cin = A.data_ptr(), B, T, Cin
        a.W, a.bias, a.taps = W.data_ptr(), L.ptr(bias), W.shape[0]
        for i, s in enumerate(shifts):
            a.shift[i] = s
        a.up, a.Cout, a.T_out_rows = 1, N, rows
        a.scale, a.resid = L.ptr(scale), L.ptr(resid)
        raw = torch.empty(B, rows, N, device=self.dev) if want_raw else None
        out_act = torch.empty(B, rows, N, device=self.dev) if act != L.ACT_NONE else None
        a.out_raw, a.out_act, a.act, a.force_fp32 = L.ptr(raw), L.ptr(out_act), act, 1
        L.check(self.lib.q3t_tapgemm(C.byref(a), L.stream_ptr()), "tapgemm")
        return raw, out_act

    def eltwise(self, op: int, a: torch.Tensor, b: Optional[torch.Tensor] = None, r: Optional[torch.Tensor] = None) -> torch.Tensor:
        out = torch.empty_like(a)
        Cn = a.shape[-1]
        L.check(self.lib.q3t_eltwise(op, a.data_ptr(), L.ptr(b), L.ptr(r), a.numel(), Cn, a.numel() // a.shape[0], out.data_ptr(),
                                     L.stream_ptr()), "eltwise")
        return out

    def time_stats(self, x: torch.Tensor, w: Optional[torch.Tensor], eps: float) -> Tuple[torch.Tensor, torch.Tensor]:
        B, T, Cn = x.shape
        mean, std = torch.empty(B, Cn, device=self.dev), torch.empty(B, Cn, device=self.dev)
        L.check(self.lib.q3t_time_stats(x.data_ptr(), L.ptr(w), B, T, Cn, eps, mean.data_ptr(), std.data_ptr(), L.stream_ptr()), "time_stats")
        return mean, std

    def layernorm(self, x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float) -> torch.Tensor:
        y = torch.empty_like(x)
        L.check(self.lib.q3t_layernorm(x.data_ptr(), w.data_ptr(), b.data_ptr(), x.numel() // x.shape[-1], x.shape[-1], eps, y.data_ptr(),
                                       L.stream_ptr()), "layernorm")
        return y


def _conv_taps(W: torch.Tensor) -> torch.Tensor:
    """nn.Conv1d weight [Cout, Cin, k] -> tap-major [k, Cout, Cin]."""
    return W.permute(2, 0, 1).contiguous()


class SpeechEncoder(_Ops):
    def __init__(self, cfg: ModelConfig, ws: WeightStore, device: str = "cuda"):
        super().__init__(device)
        self.cfg, self.e = cfg, cfg.enc
        e = self.e
        w = {n: t.to(self.dev, torch.float32).contiguous() for n, t in ws.fp.items() if n.startswith("enc.")}
        if "enc.conv_in.weight" not in w:
            raise ValueError("this checkpoint holds no speech-tokenizer encoder (needed for ref_audio)")
        self.w = w
        win = w["enc.conv_in.weight"]                                            # [F, 1, k]: one input channel, padded to 4
        self.conv_in = torch.zeros(win.shape[2], win.shape[0], 4, device=self.dev)
        self.conv_in[:, :, 0] = win[:, 0, :].t()
        self.stages = []
        for i, r in enumerate(reversed(e.ratios)):
            p = f"enc.stages.{i}"
            wd = w[p + ".down.weight"]                                            # [2C, C, 2r]
            # strided conv as a 2-tap GEMM over x viewed as [T / r, r * C]: tap 0 = kernel positions 0..r-1, tap 1 = r..2r-1
            lo = wd[:, :, :r].permute(0, 2, 1).reshape(wd.shape[0], -1)
            hi = wd[:, :, r:].permute(0, 2, 1).reshape(wd.shape[0], -1)
            self.stages.append(dict(r=r, c1=_conv_taps(w[p + ".res.conv1.weight"]), b1=w[p + ".res.conv1.bias"],
                                    c2=_conv_taps(w[p + ".res.conv2.weight"]), b2=w[p + ".res.conv2.bias"],
                                    down=torch.stack([lo, hi]).contiguous(), bd=w[p + ".down.bias"]))
        self.conv_out = _conv_taps(w["enc.conv_out.weight"])
        self.layers = []
        for l in range(e.tf_layers):
            p = f"enc.tf.layers.{l}"
            self.layers.append(dict(
                n1=(w[p + ".input_norm.weight"], w[p + ".input_norm.bias"]), n2=(w[p + ".post_norm.weight"], w[p + ".post_norm.bias"]),
                qkv=torch.cat([w[p + ".q_proj.weight"], w[p + ".k_proj.weight"], w[p + ".v_proj.weight"]], 0)[None].contiguous(),
                o=w[p + ".o_proj.weight"][None].contiguous(), s1=w[p + ".attn_scale"],
                fc1=w[p + ".fc1.weight"][None].contiguous(), fc2=w[p + ".fc2.weight"][None].contiguous(), s2=w[p + ".mlp_scale"]))
        self.inv_freq = (1.0 / (e.rope_theta ** (torch.arange(0, e.tf_head_dim, 2, dtype=torch.float32) / e.tf_head_dim))).to(self.dev)
        wd = w["enc.downsample.weight"]                                           # [H, H, 4], stride 2
        self.down = torch.stack([wd[:, :, :2].permute(0, 2, 1).reshape(wd.shape[0], -1),
                                 wd[:, :, 2:].permute(0, 2, 1).reshape(wd.shape[0], -1)]).contiguous()
        self.rvq = []
        for grp, lo_, hi_ in (("semantic", 0, e.num_semantic), ("acoustic", e.num_semantic, e.valid_quantizers)):
            tabs = [(w[f"enc.rvq.{grp}.codebooks.{i}.embed_sum"] /
                     w[f"enc.rvq.{grp}.codebooks.{i}.cluster_usage"].clamp(min=1e-5)[:, None]).contiguous() for i in range(hi_ - lo_)]
            self.rvq.append(dict(proj=w[f"enc.rvq.{grp}.in_proj.weight"][None].contiguous(), tabs=tabs,
                                 ptrs=(L.vp * len(tabs))(*[t.data_ptr() for t in tabs])))

    # ---- MimiConv1d (mimi:214-352) pieces ---------------------------------------------------------------------------------
    def _causal_conv(self, x, W, bias, act=L.ACT_NONE, want_raw=True, resid=None):
        """stride 1: left pad k - 1 = the tap-GEMM's zero fill; the right 'extra' padding is 0 for stride 1."""
        k = W.shape[0]
        return self.tap(x, W, bias, [-(k - 1 - j) for j in range(k)], x.shape[1], act, want_raw, resid=resid)

    def _strided_conv(self, x, W2, bias, stride: int, act=L.ACT_NONE, want_raw=True, pad_mode: str = "constant"):
        """kernel 2 * stride: left pad = stride, right pad completes the last frame (mimi:273-285)."""
        B, T, Cn = x.shape
        k = 2 * stride
        n_frames = math.ceil((T - k + stride) / stride + 1) - 1
        extra = n_frames * stride + k - stride - T
        xp = F.pad(x.transpose(1, 2), (stride, extra), mode=pad_mode).transpose(1, 2).contiguous()      # [B, (n_out + 1) * stride, C]
        n_out = xp.shape[1] // stride - 1
        xv = xp.view(B, n_out + 1, stride * Cn)
        return self.tap(xv, W2, bias, [0, 1], n_out, act, want_raw)

    def embeddings(self, wav: torch.Tensor, stages: Optional[dict] = None) -> torch.Tensor:
        """wav [B, n] -> pre-quantiser embeddings [B, T, hidden] at 12.5 Hz."""
        e = self.e
        B, n = wav.shape
        x4 = torch.zeros(B, n, 4, device=self.dev)
        x4[:, :, 0] = wav.to(self.dev, torch.float32)
        x, a = self._causal_conv(x4, self.conv_in, self.w["enc.conv_in.bias"], act=L.ACT_ELU)
        for i, st in enumerate(self.stages):
            _, h = self._causal_conv(a, st["c1"], st["b1"], act=L.ACT_ELU, want_raw=False)      # ELU -> conv k3 -> ELU
            x, a = self._causal_conv(h, st["c2"], st["b2"], act=L.ACT_ELU, resid=x)             # conv k1 + shortcut; ELU for the next conv
            x, a = self._strided_conv(a, st["down"], st["bd"], st["r"], act=L.ACT_ELU)
            if stages is not None:
                stages[f"stage{i}"] = x
        x, _ = self._causal_conv(a, self.conv_out, self.w["enc.conv_out.bias"])
        if stages is not None:
            stages["seanet"] = x
        T = x.shape[1]
        H, D = e.tf_heads, e.tf_head_dim
        assert (B * T * H) % 4 == 0
        for ly in self.layers:
            hn = self.layernorm(x, ly["n1"][0], ly["n1"][1], e.norm_eps)
            qkv, _ = self.tap(hn, ly["qkv"], None, [0], T)
            att = torch.empty(B, T, H * D, device=self.dev)
            L.check(self.lib.q3t_window_attn(qkv.data_ptr(), self.inv_freq.data_ptr(), B, T, H, D, e.sliding_window, att.data_ptr(),
                                             L.stream_ptr()), "window_attn")
            x, _ = self.tap(att, ly["o"], None, [0], T, scale=ly["s1"], resid=x)
            hn = self.layernorm(x, ly["n2"][0], ly["n2"][1], e.norm_eps)
            _, g = self.tap(hn, ly["fc1"], None, [0], T, act=L.ACT_GELU, want_raw=False)
            x, _ = self.tap(g, ly["fc2"], None, [0], T, scale=ly["s2"], resid=x)
        if stages is not None:
            stages["transformer"] = x
        x, _ = self._strided_conv(x, self.down, None, 2, pad_mode="replicate")
        return x

    def quantize(self, emb: torch.Tensor) -> torch.Tensor:
        """emb [B, T, hidden] -> codes [B, valid_quantizers, T] int32 (both RVQ groups quantise the same embeddings)."""
        B, T, _ = emb.shape
        nq = self.e.valid_quantizers
        idx = torch.empty(nq, B * T, device=self.dev, dtype=torch.int32)
        lv = 0
        for grp in self.rvq:
            z, _ = self.tap(emb.contiguous(), grp["proj"], None, [0], T)
            n = len(grp["tabs"])
            L.check(self.lib.q3t_rvq_encode(z.data_ptr(), grp["ptrs"], B * T, n, self.e.codebook_size, self.e.codebook_dim, B * T,
                                            idx[lv:].data_ptr(), 0, L.stream_ptr()), "rvq_encode")
            lv += n
        return idx.view(nq, B, T).permute(1, 0, 2).contiguous()

    def encode(self, wav: torch.Tensor) -> torch.Tensor:
        return self.quantize(self.embeddings(wav))


class SpeakerEncoder(_Ops):
    def __init__(self, cfg: ModelConfig, ws: WeightStore, device: str = "cuda"):
        super().__init__(device)
        self.cfg, self.s = cfg, cfg.spk
        s = self.s
        w = {n: t.to(self.dev, torch.float32).contiguous() for n, t in ws.fp.items() if n.startswith("spk.")}
        if "spk.fc.weight" not in w:
            raise ValueError("this checkpoint holds no speaker encoder (needed for ref_audio)")
        self.w = w
        self.t = {n[: -len(".weight")]: _conv_taps(t) for n, t in w.items() if n.endswith(".weight")}
        # STFT as a strided convolution against the hann-windowed DFT basis: frames of n_fft samples, hop apart =
        # n_fft / hop taps over the signal viewed as [n / hop, hop]; outputs re | im, padded to a multiple of 16 columns
        assert s.n_fft % s.hop == 0 and s.win == s.n_fft and s.hop % 4 == 0
        nf = s.n_fft // 2 + 1
        self.nf, self.ld = nf, (2 * nf + 15) // 16 * 16
        k = torch.arange(s.n_fft, dtype=torch.float64)
        win = torch.hann_window(s.win, periodic=True, dtype=torch.float64)
        ang = 2 * math.pi * torch.arange(nf, dtype=torch.float64)[:, None] * k[None, :] / s.n_fft
        basis = torch.zeros(self.ld, s.n_fft, dtype=torch.float64)
        basis[:nf] = torch.cos(ang) * win
        basis[nf:2 * nf] = -torch.sin(ang) * win
        taps = s.n_fft // s.hop
        self.dft = basis.float().view(self.ld, taps, s.hop).permute(1, 0, 2).contiguous().to(self.dev)       # [taps, ld, hop]
        self.fb = _mel_filter_bank(s.n_fft, s.n_mels, s.sample_rate, s.fmin, s.fmax).to(self.dev).contiguous()

    def log_mel(self, wav: torch.Tensor) -> torch.Tensor:
        """wav [B, n] -> [B, frames, n_mels] (reflect pad (n_fft - hop) / 2, hann STFT, |.|, slaney mel, log clamp)."""
        s = self.s
        B, n = wav.shape
        pad = (s.n_fft - s.hop) // 2
        y = F.pad(wav.to(self.dev, torch.float32)[:, None], (pad, pad), mode="reflect")[:, 0]
        frames = (y.shape[1] - s.n_fft) // s.hop + 1
        rows = frames + s.n_fft // s.hop - 1
        yv = y[:, : rows * s.hop].contiguous().view(B, rows, s.hop)
        spec, _ = self.tap(yv, self.dft, None, list(range(s.n_fft // s.hop)), frames)                       # [B, frames, ld]
        mel = torch.empty(B, frames, s.n_mels, device=self.dev)
        L.check(self.lib.q3t_mel(spec.data_ptr(), B * frames, self.ld, self.nf, self.fb.data_ptr(), s.n_mels, mel.data_ptr(), L.stream_ptr()), "mel")
        return mel

    def _tdnn(self, name: str, x: torch.Tensor, dilation: int = 1, act: int = L.ACT_RELU) -> torch.Tensor:
        """Conv1d(padding='same', padding_mode='reflect') + activation (qwen2_5_omni:2499-2521)."""
        W = self.t[name]
        k = W.shape[0]
        total = dilation * (k - 1)
        left = total // 2
        T = x.shape[1]
        if total:
            x = F.pad(x.transpose(1, 2), (left, total - left), mode="reflect").transpose(1, 2).contiguous()
        raw, a = self.tap(x, W, self.w[name + ".bias"], [j * dilation for j in range(k)], T, act, want_raw=(act == L.ACT_NONE))
        return raw if act == L.ACT_NONE else a

    def ecapa(self, mel: torch.Tensor, stages: Optional[dict] = None) -> torch.Tensor:
        s = self.s
        x = self._tdnn("spk.blocks.0.conv", mel.contiguous(), s.dilations[0])
        feats = []
        for i in range(1, len(s.channels) - 1):
            p = f"spk.blocks.{i}"
            res = x
            x = self._tdnn(p + ".tdnn1.conv", x)
            parts, prev = [], None
            for j, part in enumerate(torch.chunk(x, s.res2net_scale, dim=2)):
                part = part.contiguous()
                if j == 0:
                    prev = part
                elif j == 1:
                    prev = self._tdnn(f"{p}.res2net.{j - 1}.conv", part, s.dilations[i])
                else:
                    prev = self._tdnn(f"{p}.res2net.{j - 1}.conv", self.eltwise(0, part, prev), s.dilations[i])
                parts.append(prev)
            x = self._tdnn(p + ".tdnn2.conv", torch.cat(parts, 2).contiguous())
            m, _ = self.time_stats(x, None, 0.0)                                             # squeeze (mean over time)
            g = self._tdnn(p + ".se.conv1", m[:, None, :].contiguous())
            g = self._tdnn(p + ".se.conv2", g, act=L.ACT_SIGMOID)
            x = self.eltwise(2, x, g[:, 0].contiguous(), res)                                   # excite + residual
            feats.append(x)
            if stages is not None:
                stages[f"block{i}"] = x
        x = self._tdnn("spk.mfa.conv", torch.cat(feats, 2).contiguous(), s.dilations[-1])
        if stages is not None:
            stages["mfa"] = x
        B, T, Cn = x.shape
        mean, std = self.time_stats(x, None, 1e-12)
        att_in = torch.cat([x, mean[:, None].expand(-1, T, -1), std[:, None].expand(-1, T, -1)], 2).contiguous()
        att = self.eltwise(1, self._tdnn("spk.asp.tdnn.conv", att_in))                         # tanh(relu(conv))
        att = self._tdnn("spk.asp.conv", att, act=L.ACT_NONE)
        sm = torch.empty_like(att)
        L.check(self.lib.q3t_softmax_time(att.data_ptr(), B, T, Cn, sm.data_ptr(), L.stream_ptr()), "softmax_time")
        mean, std = self.time_stats(x, sm, 1e-12)
        pooled = torch.cat([mean, std], 1)[:, None, :].contiguous()
        return self._tdnn("spk.fc", pooled, act=L.ACT_NONE)[:, 0]

    def embed(self, wav: torch.Tensor) -> torch.Tensor:
        return self.ecapa(self.log_mel(wav))


def _mel_filter_bank(n_fft: int, n_mels: int, sr: int, fmin: float, fmax: float) -> torch.Tensor:
    """Slaney-scale, slaney-normalised triangular filters [n_fft/2+1, n_mels] (librosa.filters.mel(htk=False)); computed in
    float64 on the host once at load."""
    def hz2mel(f):
        f = np.asarray(f, dtype=np.float64)
        return np.where(f >= 1000.0, 15.0 + np.log(np.maximum(f, 1e-10) / 1000.0) * (27.0 / np.log(6.4)), 3.0 * f / 200.0)

    def mel2hz(m):
        m = np.asarray(m, dtype=np.float64)
        return np.where(m >= 15.0, 1000.0 * np.exp(np.log(6.4) / 27.0 * (m - 15.0)), 200.0 * m / 3.0)

    freqs = np.linspace(0.0, sr / 2.0, n_fft // 2 + 1)
    hz = mel2hz(np.linspace(hz2mel(fmin), hz2mel(fmax), n_mels + 2))
    fdiff = np.diff(hz)
    slopes = hz[None, :] - freqs[:, None]
    fb = np.maximum(0.0, np.minimum(-slopes[:, :-2] / fdiff[:-1], slopes[:, 2:] / fdiff[1:]))
    fb *= (2.0 / (hz[2:n_mels + 2] - hz[:n_mels]))[None]
    return torch.from_numpy(fb.astype(np.float32))
