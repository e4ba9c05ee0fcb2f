"""Per-phase timeline of the persistent stack-pass kernel (globaltimer stamps written by thread 0 of every CTA)."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qwen3-tts-apple-silicon_b200"))
from qwen3_tts_b200 import config as Cfg, lib as L
from qwen3_tts_b200.engine import TalkerEngine
from qwen3_tts_b200.weights import make_weights

cfg = Cfg.full()
ws = make_weights(cfg, seed=0, device="cuda", keep_fp=False, parts=("talker", "cp"))
e = TalkerEngine(cfg, ws, "cuda", batch=1, max_frames=64, max_ctx=1024)
del ws
lib = e.lib
NST = 1024
timing = torch.zeros(148 * NST, dtype=torch.int64, device="cuda")
a = L.StackPassArgs()
a.stack = e.talker_stack
a.head = e.fa.codec_head
a.pos, a.x_in, a.hidden_out, a.logits_out = e.pos.data_ptr(), e.x.data_ptr(), e.hidden.data_ptr(), e.logits.data_ptr()
a.work, a.counters, a.barrier, a.timing = e.mega_work.data_ptr(), e.attn_counters.data_ptr(), e.mega_barrier.data_ptr(), timing.data_ptr()
e.x.normal_(0, 0.02)
for ctx in (300,):
    e.pos.fill_(ctx)
    for it in range(3):
        timing.zero_()
        L.check(lib.q3t_stack_pass(C.byref(a), L.stream_ptr()), "stack_pass")
        torch.cuda.synchronize()
    t = timing.view(148, NST).cpu()
    names = ["qkv.pro", "qkv.gemv", "qkv.sync", "attn", "attn.sync", "o.pro", "o.gemv", "o.sync", "gu.pro", "gu.gemv",
             "gu.sync", "down.pro", "down.gemv", "down.sync"]
    nl = cfg.talker.num_layers
    print(f"ctx={ctx}: total (cta0 first->last stamp) {(t[0, 14 * nl] - t[0, 0]).item() / 1e3:.1f} us")
    for cta in (0, 73, 147):
        d = (t[cta, 1:14 * nl + 1] - t[cta, 0:14 * nl]).view(nl, 14).float() / 1e3     # us
        steady = d[2:].mean(0)
        print(f"cta {cta}: " + "  ".join(f"{n}={v:.2f}" for n, v in zip(names, steady.tolist())) + f"  | layer {steady.sum():.2f} us")
    # spread of arrival at barriers: max over CTAs minus min of the stamp BEFORE each sync
    idx = torch.tensor([1 + 14 * 5 + k for k in (1, 3, 6, 9, 12)])   # layer 5
    arr = t[:, idx].float() / 1e3
    print("layer-5 arrival spread at the 5 barriers (us):", (arr.max(0).values - arr.min(0).values).tolist())
