"""Host side of the speech-tokenizer decoder (RVQ codes -> 24 kHz waveform) on B200.

Weight re-layout for the tap-GEMM operator happens once at load; every arithmetic op is a kernel of
libq3tts_b200.so (csrc/codec.cu).  Structure follows the cousin `Code2Wav` forward + `chunked_decode`
(transformers qwen3_omni_moe/modeling_qwen3_omni_moe.py:3766-3790) with the Qwen3-TTS front end
(split-RVQ decode, pre-conv, in/out projections).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional

import torch

from . import lib as L
from .config import ModelConfig
from .weights import WeightStore


class _Tap:
    """One tap-GEMM layer: W [taps, N, Cin] + bias + shifts."""

    def __init__(self, W: torch.Tensor, bias: Optional[torch.Tensor], shifts: List[int], up: int, cout: int,
                 rows_delta: int = 0, device=None, f16: bool = False):
        # vocoder layers also keep an fp16 copy of the (unrounded) weights: between those layers the activations travel as
        # fp16 - 11 significant bits, one more than the TF32 read of fp32 data, at half the bytes (csrc/tapgemm_tc.cu, a_f16)
        self.W16 = (W.contiguous().to(torch.float16).to(device) if device is not None else W.contiguous().to(torch.float16)) if f16 else None
        # weights are rounded to TF32 (10-bit mantissa, round-to-nearest-even) once: the tcgen05 tap-GEMM reads its
        # operands as TF32 by truncation, so pre-rounded weights make that truncation exact (csrc/tapgemm_tc.cu)
        Wi = W.contiguous().view(torch.int32)
        W = ((Wi + 0x0FFF + ((Wi >> 13) & 1)) & ~0x1FFF).view(torch.float32)
        # re-layout and rounding run where the checkpoint tensors live (a host store costs no device launches); the finished
        # operands move to the device once
        W = W.contiguous().to(device) if device is not None else W.contiguous()
        bias = bias.to(device) if (bias is not None and device is not None) else bias
        self.W, self.bias, self.shifts, self.up, self.cout, self.rows_delta = W, bias, shifts, up, cout, rows_delta
        self.taps, self.cin = W.shape[0], W.shape[2]


class CodecDecoder:
    def __init__(self, cfg: ModelConfig, ws: WeightStore, device: str = "cuda"):
        self.lib = L.load()
        self.cfg, self.k, self.dev = cfg, cfg.codec, torch.device(device)
        self.f16 = os.environ.get("Q3T_CODEC_F16", "1") != "0"     # fp16 activations between the vocoder's tensor-core layers
        k = self.k
        w: Dict[str, torch.Tensor] = {n: t.to(torch.float32) for n, t in ws.fp.items() if n.startswith("codec.")}   # source device
        dev = self.dev
        D = lambda t: t.to(dev).contiguous()

        def conv(name: str, dilation: int = 1) -> _Tap:
            W = w[name + ".weight"]                      # [Cout, Cin, k]
            ks = W.shape[2]
            return _Tap(W.permute(2, 0, 1), w[name + ".bias"], [-(ks - 1 - j) * dilation for j in range(ks)], 1, W.shape[0], device=dev,
                        f16=name.startswith("codec.dec."))

        def tconv(name: str, stride: int) -> _Tap:
            W = w[name + ".weight"]                      # [Cin, Cout, k]
            cin, cout, ks = W.shape
            t0 = W[:, :, :stride].permute(2, 1, 0).reshape(stride * cout, cin)
            f16 = name.startswith("codec.dec.")
            if ks == stride:
                return _Tap(t0[None], w[name + ".bias"], [0], stride, cout, device=dev, f16=f16)
            assert ks == 2 * stride
            t1 = W[:, :, stride:].permute(2, 1, 0).reshape(stride * cout, cin)
            if k.transconv_trim == "both":               # out[q*s+p] = x[q+1] W[p] + x[q] W[p+s]   (cousin :3319-3331)
                return _Tap(torch.stack([t0, t1]), w[name + ".bias"], [1, 0], stride, cout, rows_delta=-1, device=dev, f16=f16)
            return _Tap(torch.stack([t0, t1]), w[name + ".bias"], [0, -1], stride, cout, device=dev, f16=f16)

        def lin(W: torch.Tensor, bias: Optional[torch.Tensor] = None) -> _Tap:
            return _Tap(W[None], bias, [0], 1, W.shape[0], device=dev)

        def snake(name: str):
            # evaluated in fp64 and rounded once: the same fp32 constants whether the store lives on the host or on the device
            return (D(torch.exp(w[name + ".alpha"].double()).float()), D((1.0 / (torch.exp(w[name + ".beta"].double()) + 1e-9)).float()))

        # RVQ tables: embed_sum / clamp(usage, 1e-5)  (mimi:1191-1195), one per quantizer
        self.tables = []
        for grp, nq in (("semantic", k.num_semantic), ("acoustic", k.num_quantizers - k.num_semantic)):
            for i in range(nq):
                p = f"codec.rvq.{grp}.codebooks.{i}"
                self.tables.append(D(w[p + ".embed_sum"] / w[p + ".cluster_usage"].clamp(min=1e-5)[:, None]))
        self._tab_ptrs = (L.vp * len(self.tables))(*[t.data_ptr() for t in self.tables])
        self.rvq_sem = lin(w["codec.rvq.semantic.out_proj.weight"])
        self.rvq_ac = lin(w["codec.rvq.acoustic.out_proj.weight"])
        self.pre_conv = conv("codec.pre_conv")
        self.tf_in = lin(w["codec.tf.in_proj.weight"], w["codec.tf.in_proj.bias"])
        self.tf_layers = []
        for i in range(k.tf_layers):
            p = f"codec.tf.layers.{i}"
            gu = torch.stack([w[p + ".gate_proj.weight"], w[p + ".up_proj.weight"]], 1).reshape(-1, k.tf_hidden)
            self.tf_layers.append(dict(
                n1=D(w[p + ".input_norm.weight"]),
                qkv=lin(torch.cat([w[p + ".q_proj.weight"], w[p + ".k_proj.weight"], w[p + ".v_proj.weight"]], 0)),
                o=lin(w[p + ".o_proj.weight"]), s1=D(w[p + ".attn_scale"]),
                n2=D(w[p + ".post_norm.weight"]), gu=lin(gu), down=lin(w[p + ".down_proj.weight"]),
                s2=D(w[p + ".mlp_scale"])))
        self.tf_norm = D(w["codec.tf.norm.weight"])
        self.tf_out = lin(w["codec.tf.out_proj.weight"], w["codec.tf.out_proj.bias"])
        self.tf_inv_freq = (1.0 / (k.tf_rope_theta ** (torch.arange(0, k.tf_head_dim, 2, dtype=torch.float32) /
                                                      k.tf_head_dim))).to(self.dev)
        self.ups = []
        for i, r in enumerate(k.upsampling_ratios):
            p = f"codec.up.{i}"
            self.ups.append(dict(tconv=tconv(p + ".tconv", r), dw_w=D(w[p + ".cnx.dw.weight"].reshape(k.latent_dim, -1)),
                                 dw_b=D(w[p + ".cnx.dw.bias"]), ln_w=D(w[p + ".cnx.ln.weight"]), ln_b=D(w[p + ".cnx.ln.bias"]),
                                 pw1=lin(w[p + ".cnx.pw1.weight"], w[p + ".cnx.pw1.bias"]),
                                 pw2=lin(w[p + ".cnx.pw2.weight"], w[p + ".cnx.pw2.bias"]), gamma=D(w[p + ".cnx.gamma"])))
        self.conv_in = conv("codec.dec.conv_in")
        self.blocks = []
        for i, r in enumerate(k.upsample_rates):
            p = f"codec.dec.blocks.{i}"
            units = []
            for j, d in enumerate((1, 3, 9)):
                u = f"{p}.units.{j}"
                units.append(dict(s1=snake(u + ".snake1"), c1=conv(u + ".conv1", d), s2=snake(u + ".snake2"),
                                  c2=conv(u + ".conv2")))
            self.blocks.append(dict(snake=snake(p + ".snake"), tconv=tconv(p + ".tconv", r), units=units))
        self.snake_out = snake("codec.dec.snake_out")
        self.conv_out = conv("codec.dec.conv_out")
        self.conv_out_w = D(w["codec.dec.conv_out.weight"][0].t())      # [taps, C] unrounded fp32 (FP32-pipe kernel)

    # ---- operator wrappers ---------------------------------------------------------------------------------
    def _shape_args(self, layer: _Tap, B: int, T: int) -> "L.TapGemmArgs":
        a = L.TapGemmArgs()
        a.B, a.T_in, a.Cin, a.taps = B, T, layer.cin, layer.taps
        a.up, a.Cout, a.T_out_rows = layer.up, layer.cout, T + layer.rows_delta
        return a

    def _want16(self, layer: _Tap, B: int, T: int) -> bool:
        """May the producer of this layer's input write fp16?  Only if the layer will run on the tcgen05 kernel."""
        return self.f16 and layer.W16 is not None and bool(self.lib.q3t_tapgemm_tc_eligible(C.byref(self._shape_args(layer, B, T))))

    def _tap(self, layer: _Tap, A: torch.Tensor, scale=None, resid=None, want_raw=True, act=L.ACT_NONE, act_ab=None, act16=False):
        """A [B, T, Cin] (fp32, or fp16 from a producer that was asked for it) -> (raw [B, rows*up, Cout] fp32 or None, act or
        None; act is fp16 when act16 is set and this layer runs on the tcgen05 kernel)."""
        B, T, Cin = A.shape
        assert Cin == layer.cin and A.is_contiguous()
        rows = T + layer.rows_delta
        N = layer.up * layer.cout
        a = self._shape_args(layer, B, T)
        a16 = A.dtype == torch.float16
        a.A, a.a_f16 = A.data_ptr(), int(a16)
        a.W, a.bias = (layer.W16 if a16 else layer.W).data_ptr(), L.ptr(layer.bias)
        act16 = bool(act16 and act not in (L.ACT_NONE, L.ACT_SWIGLU_PAIR) and self.lib.q3t_tapgemm_tc_eligible(C.byref(a)))
        for i, s in enumerate(layer.shifts):
            a.shift[i] = s
        a.scale, a.resid = L.ptr(scale), L.ptr(resid)
        raw = torch.empty(B, rows * layer.up, layer.cout, device=self.dev, dtype=torch.float32) if want_raw else None
        out_act = None
        if act != L.ACT_NONE:
            shape = (B, rows, N // 2) if act == L.ACT_SWIGLU_PAIR else (B, rows * layer.up, layer.cout)
            out_act = torch.empty(*shape, device=self.dev, dtype=torch.float16 if act16 else torch.float32)
        a.out_raw, a.out_act, a.act, a.act_f16 = L.ptr(raw), L.ptr(out_act), act, int(act16)
        if act_ab is not None:
            a.act_a, a.act_b = act_ab[0].data_ptr(), act_ab[1].data_ptr()
        L.check(self.lib.q3t_tapgemm(C.byref(a), L.stream_ptr()), "tapgemm")
        return raw, out_act

    def _rmsnorm(self, x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
        y = torch.empty_like(x)
        L.check(self.lib.q3t_rmsnorm(x.data_ptr(), w.data_ptr(), y.data_ptr(), x.numel() // x.shape[-1], x.shape[-1],
                                     eps, L.stream_ptr()), "rmsnorm")
        return y

    def rvq_sums(self, codes: torch.Tensor):
        """codes [B, G, T] int -> (semantic sum [B,T,dim], acoustic sum [B,T,dim]); bit-exact gather + fp32 adds."""
        k = self.k
        codes = codes.to(self.dev, torch.int32).contiguous()
        B, G, T = codes.shape
        outs = []
        for lo, hi in ((0, k.num_semantic), (k.num_semantic, k.num_quantizers)):
            o = torch.empty(B, T, k.codebook_dim, device=self.dev, dtype=torch.float32)
            L.check(self.lib.q3t_rvq_gather_sum(codes.data_ptr(), self._tab_ptrs, B, G, T, lo, hi, k.codebook_dim,
                                                k.codebook_size, o.data_ptr(), L.stream_ptr()), "rvq_gather_sum")
            outs.append(o)
        return outs

    # ---- one vocoder call -----------------------------------------------------------------------------------
    def forward(self, codes: torch.Tensor, stages: Optional[dict] = None) -> torch.Tensor:
        """codes [B, 16, T] -> wav [B, out_len(T)] fp32 in [-1, 1]."""
        k = self.k
        sem, ac = self.rvq_sums(codes)
        y, _ = self._tap(self.rvq_sem, sem)
        x, _ = self._tap(self.rvq_ac, ac, resid=y)                        # [B, T, 512]
        x, _ = self._tap(self.pre_conv, x)                                # [B, T, 1024]
        if stages is not None:
            stages["pre_conv"] = x
        h, _ = self._tap(self.tf_in, x)
        B, T, _ = h.shape
        for ly in self.tf_layers:
            n = self._rmsnorm(h, ly["n1"], k.tf_rms_eps)
            qkv, _ = self._tap(ly["qkv"], n)
            a = torch.empty(B, T, k.tf_heads * k.tf_head_dim, device=self.dev, dtype=torch.float32)
            L.check(self.lib.q3t_window_attn(qkv.data_ptr(), self.tf_inv_freq.data_ptr(), B, T, k.tf_heads, k.tf_head_dim,
                                             k.sliding_window, a.data_ptr(), L.stream_ptr()), "window_attn")
            h, _ = self._tap(ly["o"], a, scale=ly["s1"], resid=h)
            n = self._rmsnorm(h, ly["n2"], k.tf_rms_eps)
            _, g = self._tap(ly["gu"], n, want_raw=False, act=L.ACT_SWIGLU_PAIR)
            h, _ = self._tap(ly["down"], g, scale=ly["s2"], resid=h)
        h = self._rmsnorm(h, self.tf_norm, k.tf_rms_eps)
        x, _ = self._tap(self.tf_out, h)                                  # [B, T, 1024]
        if stages is not None:
            stages["transformer"] = x
        for up in self.ups:
            x, _ = self._tap(up["tconv"], x)
            Bn, Tn, Cn = x.shape
            n = torch.empty_like(x)
            L.check(self.lib.q3t_dwconv_ln(x.data_ptr(), up["dw_w"].data_ptr(), up["dw_b"].data_ptr(), up["ln_w"].data_ptr(),
                                           up["ln_b"].data_ptr(), 1e-6, Bn, Tn, Cn, up["dw_w"].shape[1], n.data_ptr(),
                                           L.stream_ptr()), "dwconv_ln")
            _, g = self._tap(up["pw1"], n, want_raw=False, act=L.ACT_GELU)
            x, _ = self._tap(up["pw2"], g, scale=up["gamma"], resid=x)
        if stages is not None:
            stages["upsample"] = x
        # vocoder: every conv writes the SnakeBeta of its consumer in its epilogue - as fp16 wherever that consumer runs on the
        # tcgen05 kernel (all of them at the BASELINE shapes); the residual stream `u` stays fp32
        co = self.conv_out
        co_direct = co.cout == 1 and co.cin % 4 == 0 and co.shifts == [-(co.taps - 1 - j) for j in range(co.taps)]
        Bn, T, _ = x.shape
        _, act = self._tap(self.conv_in, x, want_raw=False, act=L.ACT_SNAKE, act_ab=self.blocks[0]["snake"],
                           act16=self._want16(self.blocks[0]["tconv"], Bn, T))
        for bi, blk in enumerate(self.blocks):
            tc = blk["tconv"]
            T = (T + tc.rows_delta) * tc.up
            u, act = self._tap(tc, act, act=L.ACT_SNAKE, act_ab=blk["units"][0]["s1"], act16=self._want16(blk["units"][0]["c1"], Bn, T))
            for ui, unit in enumerate(blk["units"]):
                _, a2 = self._tap(unit["c1"], act, want_raw=False, act=L.ACT_SNAKE, act_ab=unit["s2"], act16=self._want16(unit["c2"], Bn, T))
                if ui + 1 < len(blk["units"]):
                    nxt, nxt16 = blk["units"][ui + 1]["s1"], self._want16(blk["units"][ui + 1]["c1"], Bn, T)
                elif bi + 1 < len(self.blocks):
                    nxt, nxt16 = self.blocks[bi + 1]["snake"], self._want16(self.blocks[bi + 1]["tconv"], Bn, T)
                else:
                    nxt, nxt16 = self.snake_out, self.f16 and co_direct
                u, act = self._tap(unit["c2"], a2, resid=u, want_raw=True, act=L.ACT_SNAKE, act_ab=nxt, act16=nxt16)
            if stages is not None:
                stages[f"block{bi}"] = u
        Bn, Tn, Cn = act.shape
        out = torch.empty(Bn, Tn, device=self.dev, dtype=torch.float32)
        if co_direct:
            # one output channel is not GEMM-shaped: dedicated HBM-bound kernel, clamp fused
            fn = self.lib.q3t_conv_out_clamp_h if act.dtype == torch.float16 else self.lib.q3t_conv_out_clamp
            L.check(fn(act.data_ptr(), Bn, Tn, Cn, self.conv_out_w.data_ptr(), L.ptr(co.bias), co.taps, out.data_ptr(), L.stream_ptr()),
                    "conv_out_clamp")
            return out
        wav, _ = self._tap(co, act)                                         # [B, n, 1]
        L.check(self.lib.q3t_clamp_pcm16(wav.data_ptr(), wav.numel(), out.data_ptr(), 0, L.stream_ptr()), "clamp")
        return out

    def decode_interval(self, codes: torch.Tensor, start: int, end: int) -> torch.Tensor:
        """Samples of frames [start, end) of codes [B,16,T'] (T' >= end): one vocoder call over the interval plus
        `left_context` frames of history, the history's samples dropped (one step of the chunked decode below; the
        streaming path of BASELINE config 3 calls it once per interval as the frames arrive)."""
        k = self.k
        ctx = k.left_context if start - k.left_context > 0 else start
        wav = self.forward(codes[..., start - ctx:end])
        return wav[:, ctx * k.hop:]

    def decode(self, codes: torch.Tensor, chunk_size: Optional[int] = None) -> torch.Tensor:
        """Chunked decode (300-frame chunks, 25 frames of left context; cousin :3780-3790). codes [B,16,T] -> [B, n]."""
        k = self.k
        chunk = chunk_size or k.chunk_size
        wavs, start, T = [], 0, codes.shape[-1]
        while start < T:
            end = min(start + chunk, T)
            wavs.append(self.decode_interval(codes, start, end))
            start = end
        return torch.cat(wavs, -1)
