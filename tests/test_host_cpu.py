"""CPU tests: quantiser, HBM packing, C-ABI library exports, oracle vs committed golden fixtures, shim surface."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import qwen3_tts_oracle as O
from qwen3_tts_b200 import config as Cfg
from qwen3_tts_b200 import lib as L
from qwen3_tts_b200.weights import (TILE_BYTES, dequantize_w8, make_weights, pack_w8, quantize_w8, unpack_w8)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "small_golden.npz")


# ---- quantiser --------------------------------------------------------------------------
def test_quantizer_error_bound_and_edges():
    g = torch.Generator().manual_seed(0)
    w = torch.randn(64, 512, generator=g) * 0.02
    w[0, :64] = 0.0                       # all-zero group
    w[1, :64] = 0.5                       # constant group
    w[2, :64] = torch.linspace(-1, 0, 64)  # min edge dominates
    q, s, b = quantize_w8(w)
    assert q.dtype == torch.uint8 and s.dtype == torch.bfloat16 and s.shape == (64, 8)
    wd = dequantize_w8(q, s, b)
    step = s.float().abs().repeat_interleave(64, 1)
    # bf16 storage of scale/bias adds up to 2^-8 relative error on top of half a quantisation step
    assert ((wd - w).abs() <= 0.5 * step + 2 ** -7 * w.abs().amax() + 1e-8).all()
    assert torch.all(wd[0, :64] == 0)
    assert (wd[1, :64] - 0.5).abs().max() < 0.5 * 2 ** -7


def test_pack_unpack_roundtrip_and_fragment_order():
    g = torch.Generator().manual_seed(1)
    n, k = 48, 512
    q = torch.randint(0, 256, (n, k), generator=g, dtype=torch.uint8)
    s = torch.randn(n, k // 64, generator=g).bfloat16()
    b = torch.randn(n, k // 64, generator=g).bfloat16()
    blob = pack_w8(q, s, b)
    assert blob.numel() == (n // 16) * (k // 256) * TILE_BYTES
    q2, s2, b2 = unpack_w8(blob, n, k)
    assert torch.equal(q, q2) and torch.equal(s, s2) and torch.equal(b, b2)
    # spot-check the documented permutation: tile (rt=1, kc=1), group j4=2, mma j=1, lane 13, reg 3, byte 2
    rt, kc, j4, j, lane, i, byte = 1, 1, 2, 1, 13, 3, 2
    gq, t = lane >> 2, lane & 3
    row = rt * 16 + gq + 8 * (i & 1)
    col = kc * 256 + 64 * j4 + 32 * j + 16 * (i >> 1) + 4 * t + byte
    off = (rt * (k // 256) + kc) * TILE_BYTES + (j4 * 2 + j) * 512 + 16 * lane + 4 * i + byte
    assert int(blob[off]) == int(q[row, col])
    with pytest.raises(AssertionError):
        pack_w8(q[:40], s[:40], b[:40])


def test_config_roundtrip_and_sizes():
    cfg = Cfg.full("voice_design")
    assert Cfg.ModelConfig.from_dict(cfg.to_dict()).to_dict() == cfg.to_dict()
    t = cfg.talker
    per_layer = t.hidden_size * (t.q_dim + 2 * t.kv_dim) + t.q_dim * t.hidden_size + 3 * t.hidden_size * t.intermediate_size
    assert t.num_layers * per_layer + t.vocab_size * t.hidden_size == 1_415_577_600
    assert cfg.codec.hop == 1920 and cfg.codec.sample_rate == 24000
    assert set(s.lower() for v in [["Ryan", "Aiden", "Serena", "Vivian"], ["Uncle_Fu", "Dylan", "Eric"], ["Ono_Anna"],
                                   ["Sohee"]] for s in v) == set(t.spk_id)           # reference config.py:44-49


# ---- C ABI ------------------------------------------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "q3tts_b200.h")).read()
    declared = set(re.findall(r"\b(q3t_[a-z0-9_]+)\s*\(", header))
    assert declared == set(L.SYMBOLS), declared ^ set(L.SYMBOLS)
    if not os.path.exists(L.LIB_PATH):
        subprocess.run(["bash", os.path.join(ROOT, "qwen3-tts-apple-silicon_b200", "csrc", "build.sh")], check=True)
    lib = L.load()
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert lib.q3t_abi_version() == 2
    # struct mirrors must match the C layout: compile a tiny probe against the header
    probe = r'''
    #include "q3tts_b200.h"
    #include <stdio.h>
    int main(){ printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(q3t_w8), sizeof(q3t_gemv_args), sizeof(q3t_attn_args),
        sizeof(q3t_sampling), sizeof(q3t_sample_args), sizeof(q3t_layer), sizeof(q3t_stack), sizeof(q3t_frame_args),
        sizeof(q3t_tapgemm_args), sizeof(q3t_gemm_args), sizeof(q3t_attn_prefill_args), sizeof(q3t_prefill_args)); return 0; }'''
    src, exe = "/tmp/q3t_probe.c", "/tmp/q3t_probe"
    open(src, "w").write(probe)
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
    sizes = list(map(int, subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()))
    mirrors = [L.W8, L.GemvArgs, L.AttnArgs, L.Sampling, L.SampleArgs, L.Layer, L.Stack, L.FrameArgs, L.TapGemmArgs,
               L.GemmArgs, L.AttnPrefillArgs, L.PrefillArgs]
    assert sizes == [ctypes.sizeof(m) for m in mirrors]
    # ... and the fields round 2 appended sit where the C compiler puts them
    probe2 = r'''
    #include "q3tts_b200.h"
    #include <stddef.h>
    #include <stdio.h>
    int main(){ printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", offsetof(q3t_gemm_args, y_norm_w), offsetof(q3t_gemm_args, x_rowss_parts),
        offsetof(q3t_tapgemm_args, a_f16), offsetof(q3t_tapgemm_args, act_f16), offsetof(q3t_frame_args, active),
        offsetof(q3t_frame_args, gemm_rowss), offsetof(q3t_prefill_args, rowss), offsetof(q3t_sample_args, step_stride)); return 0; }'''
    open(src, "w").write(probe2)
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
    offs = list(map(int, subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()))
    assert offs == [L.GemmArgs.y_norm_w.offset, L.GemmArgs.x_rowss_parts.offset, L.TapGemmArgs.a_f16.offset, L.TapGemmArgs.act_f16.offset,
                    L.FrameArgs.active.offset, L.FrameArgs.gemm_rowss.offset, L.PrefillArgs.rowss.offset, L.SampleArgs.step_stride.offset]


def test_product_path_has_no_cpu_fallback_and_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "qwen3-tts-apple-silicon_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), f"{f} imports the oracle"
                assert "qwen3_tts_oracle" not in txt, f"{f} references the oracle module"
    if not torch.cuda.is_available():
        from qwen3_tts_b200.model import Model
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            Model(Cfg.small(), make_weights(Cfg.small(), parts=()), "cuda")


# ---- oracle vs committed golden fixtures ---------------------------------------------------------------------------
@pytest.fixture(scope="module")
def small():
    cfg = Cfg.small("custom_voice")
    return cfg, make_weights(cfg, seed=0, head_std=0.2)


def test_oracle_reproduces_golden(small):
    cfg, ws = small
    gold = np.load(GOLD)
    m = O.OracleModel(cfg, ws.fp, kv_dtype=torch.bfloat16)
    pre, tr = m.build_prefill(gold["text_ids"].tolist(), instruct_ids=[3, 1, 4, 1, 5], speaker="vivian", language="chinese")
    assert np.allclose(pre.numpy(), gold["prefill"], rtol=1e-5, atol=1e-6)
    codes, rec = m.generate(pre, tr, gold["codes"].shape[0], record=True)
    assert np.array_equal(codes.numpy(), gold["codes"])
    assert np.allclose(rec["talker_logits"][0].numpy(), gold["talker_logits0"], rtol=1e-4, atol=1e-5)
    cc = torch.from_numpy(gold["codec_codes"])
    _, sums = O.rvq_decode(ws.fp, cfg, cc, split=True)
    assert np.array_equal(sums[0].numpy(), gold["rvq_sem"]) and np.array_equal(sums[1].numpy(), gold["rvq_ac"])
    wav = O.codec_forward(ws.fp, cfg, cc)[0, 0]
    assert np.allclose(wav.numpy(), gold["wav"], rtol=1e-4, atol=1e-5)


def test_oracle_edge_cases(small):
    cfg, ws = small
    m = O.OracleModel(cfg, ws.fp)
    # empty text body, no speaker (voice_design style), auto language
    ids = [cfg.im_start_id, cfg.assistant_id, 10, cfg.im_end_id, 10, cfg.im_start_id, cfg.assistant_id, 10]
    pre, tr = m.build_prefill(ids)
    assert pre.shape[0] == 3 + 4 + 1 + 1 and tr.shape[0] == 1
    # streaming with a single body token: trailing = [eos, pad]
    pre_s, tr_s = m.build_prefill(ids[:3] + [7] + ids[3:], streaming=True)
    assert tr_s.shape[0] == 2
    # EOS stops generation: force it at frame 2
    forced = torch.zeros(4, 16, dtype=torch.long)
    forced[2, 0] = cfg.talker.codec_eos_id
    codes = m.generate(pre, tr, 4, forced_codes=forced)
    assert codes.shape == (2, 16)
    # one-frame clip through the codec; the trimmed transposed convs make it shorter than 1920 samples
    wav = O.codec_forward(ws.fp, cfg, torch.zeros(1, 16, 1, dtype=torch.long))
    assert wav.shape[-1] == cfg.codec.out_len(1) and wav.abs().max() <= 1


def test_oracle_prefill_is_causal_and_cache_consistent(small):
    cfg, ws = small
    t = cfg.talker
    x = torch.randn(7, t.hidden_size, generator=torch.Generator().manual_seed(3)) * 0.1
    a = O.DecoderStack(ws.fp, "talker", t.num_layers, t.num_heads, t.num_kv_heads, t.head_dim, t.rms_norm_eps, t.rope_theta)
    full = a.forward(x)
    b = O.DecoderStack(ws.fp, "talker", t.num_layers, t.num_heads, t.num_kv_heads, t.head_dim, t.rms_norm_eps, t.rope_theta)
    inc = torch.cat([b.forward(x[i:i + 1]) for i in range(7)])
    assert torch.allclose(full, inc, rtol=1e-4, atol=1e-5)


# ---- shim surface --------------------------------------------------------------------------------------------------
def test_shim_modules_export_the_reference_names():
    import inspect
    from mlx_audio.tts import generate as G, utils as U
    assert callable(U.load_model) and callable(G.generate_audio)
    params = inspect.signature(G.generate_audio).parameters
    for kw in ("model", "text", "voice", "instruct", "speed", "ref_audio", "ref_text", "output_path"):
        assert kw in params                       # custom.py:163-170, design.py:76-81, clone.py:218-224
    with pytest.raises(OSError):
        U.load_model("/nonexistent/model/dir")    # reference reports OSError as "Failed to load model" (io.py:115)
    with pytest.raises(ValueError):
        G.generate_audio(text="hi", model=None)


def test_write_wav_is_pcm16_mono_24k(tmp_path):
    import wave
    from qwen3_tts_b200.model import write_wav
    p = str(tmp_path / "audio_000.wav")
    write_wav(p, np.array([0.0, 0.5, -1.5, 1.0], dtype=np.float32), 24000)
    with wave.open(p) as w:
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (1, 2, 24000, 4)
        assert np.frombuffer(w.readframes(4), dtype="<i2").tolist() == [0, 16384, -32767, 32767]


def test_long_text_segmentation():
    """SURVEY 8f-4 / BASELINE config 5 ("2-min texts chunked"): whole sentences packed into segments of bounded length, in
    order, nothing dropped; over-long sentences cut at commas / spaces, unspaced runs at the limit."""
    from qwen3_tts_b200.text import segment_text, split_sentences
    t = ("Hello there. How are you today? I am fine! 你好。今天天气不错！Line one\nLine two; and more, with commas, and so on "
         "and so forth without any end in sight whatsoever")
    assert split_sentences(t)[:5] == ["Hello there.", "How are you today?", "I am fine!", "你好。", "今天天气不错！"]
    squash = lambda x: "".join(x.split())
    for m in (30, 60, 200):
        segs = segment_text(t, m)
        assert all(0 < len(x) <= m for x in segs)
        assert squash("".join(segs)) == squash(t)
    assert segment_text(t, 1000) == [t] and segment_text(t, 0) == [t] and segment_text("   ", 10) == []
    assert segment_text("a" * 95, 30) == ["a" * 30, "a" * 30, "a" * 30, "a" * 5]
    two_minutes = " ".join(f"This is sentence number {i} of a long narration." for i in range(60))
    segs = segment_text(two_minutes, 400)
    assert len(segs) >= 6 and all(s.endswith(".") for s in segs)
