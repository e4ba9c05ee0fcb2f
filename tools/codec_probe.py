"""Codec decoder alone (BASELINE config 2 shapes): codes [B, 16, T] -> 24 kHz wav.  usage: codec_probe.py [B] [T]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qwen3-tts-apple-silicon_b200"))
from qwen3_tts_b200 import config as Cfg
from qwen3_tts_b200.codec import CodecDecoder
from qwen3_tts_b200.weights import make_weights
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 375
cfg = Cfg.full()
ws = make_weights(cfg, seed=0, device="cuda", keep_fp=True, parts=("codec",))
dec = CodecDecoder(cfg, ws, "cuda")
codes = torch.randint(0, 2048, (B, 16, T), device="cuda", dtype=torch.int32)
wav = dec.decode(codes); torch.cuda.synchronize()
s, f = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record(); wav = dec.decode(codes); f.record(); torch.cuda.synchronize()
ms = s.elapsed_time(f)
print(f"codec B={B} T={T}: {ms:.1f} ms -> {B * T * 4.96 / ms:.1f} TFLOP/s, RTFx {B * T * 0.08 / (ms / 1e3):.0f}, wav {tuple(wav.shape)}")
