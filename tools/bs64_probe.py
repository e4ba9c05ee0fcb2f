"""Batch-64 serving probe (BASELINE config 4 shapes): GEMM prefill, batched frame graph, batched codec decode."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qwen3-tts-apple-silicon_b200"))
from qwen3_tts_b200 import config as Cfg
from qwen3_tts_b200.engine import TalkerEngine
from qwen3_tts_b200.codec import CodecDecoder
from qwen3_tts_b200.weights import make_weights
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
L = int(sys.argv[2]) if len(sys.argv) > 2 else 300
T = int(sys.argv[3]) if len(sys.argv) > 3 else 24
NS = int(sys.argv[4]) if len(sys.argv) > 4 else 4
cfg = Cfg.full("voice_design")
ws = make_weights(cfg, seed=0, device="cuda", keep_fp=False)
e = TalkerEngine(cfg, ws, "cuda", batch=B, max_frames=T + 8, max_ctx=((L + T + 8 + 15) // 16) * 16 + 16, attn_nsplit=NS)
e.set_sampling(do_sample=False)
emb = torch.randn(B, L, cfg.talker.hidden_size, device="cuda") * 0.02
ev = lambda: torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e.prefill(emb, None, None)          # warm-up incl. graph capture
torch.cuda.synchronize()
s, f = ev(), ev()
s.record(); e.prefill(emb, None, None); f.record(); torch.cuda.synchronize()
ms_pre = s.elapsed_time(f)
flops = 2.0 * 1.4157e9 * B * L
print(f"prefill B={B} L={L}: {ms_pre:.1f} ms  ({flops / ms_pre / 1e9:.0f} TFLOP/s incl. attention)   launches/frame {e.launches_per_frame}")
e.generate(4, check_every=0); torch.cuda.synchronize()
s.record(); codes = e.generate(T, check_every=0); f.record(); torch.cuda.synchronize()
ms_frame = s.elapsed_time(f) / T
print(f"frame (graph) B={B}: {ms_frame:.2f} ms/frame -> generation-only RTFx {B * 0.08 / (ms_frame / 1e3):.0f}")
codec = CodecDecoder(cfg, ws, "cuda")
cc = torch.randint(0, 2048, (B, 16, T), device="cuda", dtype=torch.int32)
wav = codec.decode(cc)
torch.cuda.synchronize()
s.record(); wav = codec.decode(cc); f.record(); torch.cuda.synchronize()
ms_codec = s.elapsed_time(f) / T
print(f"codec B={B} T={T}: {ms_codec:.2f} ms per frame-step ({B * 4.96 / ms_codec:.1f} TFLOP/s) wav {tuple(wav.shape)}")
print(f"end-to-end RTFx estimate at B={B} (excl. prefill): {B * 0.08 / ((ms_frame + ms_codec) / 1e3):.0f}")
