"""Per-phase timeline of the persistent data-flow kernel (csrc/frame_ll.cu): globaltimer stamps written by thread 0 of
every CTA.  Also cross-checks the talker logits against the one-kernel-per-contraction path."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qwen3-tts-apple-silicon_b200"))
from qwen3_tts_b200 import config as Cfg, lib as L
from qwen3_tts_b200.engine import TalkerEngine
from qwen3_tts_b200.weights import make_weights

size = sys.argv[1] if len(sys.argv) > 1 else "full"
cfg = getattr(Cfg, size)()
ws = make_weights(cfg, seed=0, device="cuda", keep_fp=False, parts=("talker", "cp"))
e = TalkerEngine(cfg, ws, "cuda", batch=1, max_frames=64, max_ctx=1024)
del ws
lib = e.lib
NST = 2048
G = torch.cuda.get_device_properties(0).multi_processor_count
timing = torch.zeros(G * NST, dtype=torch.int64, device="cuda")
e.use_graphs = False
e.x.normal_(0, 0.02)
x0 = e.x.clone()
ctx = 300 if size == "full" else 40
# ---- correctness: persistent vs multi-kernel talker step
e.set_mega(False); e.use_graphs = False
e.pos.fill_(ctx); e.x.copy_(x0); e._talker_step(True); torch.cuda.synchronize()
ref = e.logits.clone(); refh = e.hidden.clone()
e.set_mega(True); e.use_graphs = False
e.pos.fill_(ctx); e.x.copy_(x0); e._talker_step(True); torch.cuda.synchronize()
print("state", e.ll_state.tolist())
rel = float((e.logits - ref).abs().max() / ref.abs().max())
relh = float((e.hidden - refh).abs().max() / refh.abs().max())
print(f"talker step persistent vs multi-kernel: logits rel {rel:.2e} hidden rel {relh:.2e}")
# ---- timeline of the talker step
e.fa.ll_timing = timing.data_ptr()
nl = cfg.talker.num_layers
names = ["qkv.pro", "qkv.gemv", "attn", "o.pro", "o.gemv", "gu.pro", "gu.gemv", "down.pro", "down.gemv"]
for it in range(3):
    timing.zero_(); e.pos.fill_(ctx); e.x.copy_(x0)
    e._talker_step(True); torch.cuda.synchronize()
t = timing.view(G, NST).cpu()
ns = 1 + 9 * nl + 1
print(f"ctx={ctx}: talker step, CTA0 first->last stamp {(t[0, ns - 1] - t[0, 0]).item() / 1e3:.1f} us; "
      f"max over CTAs {(t[:, ns - 1].max() - t[:, 0].min()).item() / 1e3:.1f} us")
for cta in (0, 1, 73, G - 1):
    d = (t[cta, 1:9 * nl + 1] - t[cta, 0:9 * nl]).view(nl, 9).float() / 1e3
    steady = d[1:].mean(0)
    print(f"cta {cta:3d}: " + "  ".join(f"{n}={v:.2f}" for n, v in zip(names, steady.tolist())) + f"  | layer {steady.sum():.2f} us")
# CUDA-event time of back-to-back steps
s, f = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e.fa.ll_timing = 0
e.pos.fill_(ctx)
reps = 50
s.record()
for _ in range(reps):
    e._talker_step(True)
f.record(); torch.cuda.synchronize()
us = s.elapsed_time(f) * 1e3 / reps
t_ = cfg.talker
params = t_.num_layers * (t_.hidden_size * (t_.q_dim + 2 * t_.kv_dim) + t_.q_dim * t_.hidden_size + 3 * t_.hidden_size * t_.intermediate_size) + t_.vocab_size * t_.hidden_size
print(f"talker step (eager, events): {us:.1f} us -> {params * 1.0625 / us / 1e3:.0f} GB/s weights-only")
# ---- whole frame
e.reset(); e.pos.fill_(ctx); e.x.copy_(x0); e._talker_step(True)
e.fa.ll_timing = timing.data_ptr()
for it in range(2):
    timing.zero_()
    e._frame(); torch.cuda.synchronize()
t = timing.view(G, NST).cpu()
n_used = int((t[0] > 0).sum())
print(f"frame: {n_used} stamps, CTA0 first->last {(t[0, n_used - 1] - t[0, 0]).item() / 1e3:.1f} us; state {e.ll_state.tolist()}; codes {e.cur_codes.tolist()}")
e.fa.ll_timing = 0
s.record()
for _ in range(20):
    e._frame()
f.record(); torch.cuda.synchronize()
print(f"frame (eager, events): {s.elapsed_time(f) * 1e3 / 20:.1f} us -> RTFx {80e3 / (s.elapsed_time(f) * 1e3 / 20):.1f}")
