"""Debug aid: batch-64 leg of bench.py with range checks on the generated codes (optionally after a batch-1 run and/or
with freed memory poisoned by NaNs, to expose reads of uninitialised buffers)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qwen3-tts-apple-silicon_b200"))
from qwen3_tts_b200 import config as Cfg
from qwen3_tts_b200.engine import TalkerEngine
from qwen3_tts_b200.weights import make_weights

mode = sys.argv[1] if len(sys.argv) > 1 else "plain"
cfg = Cfg.full("voice_design")
if "bs1" in mode:
    ws = make_weights(cfg, seed=0, device="cuda", keep_fp=False, parts=("talker", "cp"))
    e1 = TalkerEngine(cfg, ws, "cuda", batch=1, max_frames=64, max_ctx=512)
    e1.set_sampling(do_sample=False)
    e1.prefill(torch.randn(1, 90, cfg.talker.hidden_size, device="cuda") * 0.02, None, None)
    c1 = e1.generate(32, check_every=0)
    torch.cuda.synchronize()
    print("bs1 codes max", int(c1.max()), "min", int(c1.min()))
    del e1, ws
if "poison" in mode:
    junk = torch.full((3 << 30,), float("nan"), device="cuda", dtype=torch.bfloat16)
    torch.cuda.synchronize()
    del junk
B, Lp, T = 64, 300, (int(sys.argv[2]) if len(sys.argv) > 2 else 24)
ws = make_weights(cfg, seed=0, device="cuda", keep_fp=False)
eng = TalkerEngine(cfg, ws, "cuda", batch=B, max_frames=T + 8, max_ctx=((Lp + T + 24) // 16) * 16, attn_nsplit=4)
eng.set_sampling(do_sample=False)
emb = torch.randn(B, Lp, cfg.talker.hidden_size, device="cuda") * 0.02
for it in range(2):
    eng.prefill(emb, None, None)
    torch.cuda.synchronize()
    print(f"run {it}: logits finite {bool(torch.isfinite(eng.logits).all())}  hidden finite {bool(torch.isfinite(eng.hidden).all())}")
    codes = eng.generate(T, check_every=0)
    torch.cuda.synchronize()
    bad = (codes[..., 0] >= 2048).nonzero()
    print(f"run {it}: codes max {int(codes.max())} min {int(codes.min())}  code0>=2048 at {bad[:4].tolist()}  logits finite {bool(torch.isfinite(eng.logits).all())}")
