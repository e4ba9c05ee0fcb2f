"""Fine-grained stamps inside the attention phase (library built with -DLL_FINE): CTA 0, thread 0."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qwen3-tts-apple-silicon_b200"))
from qwen3_tts_b200 import config as Cfg
from qwen3_tts_b200.engine import TalkerEngine
from qwen3_tts_b200.weights import make_weights
size = sys.argv[1] if len(sys.argv) > 1 else "full"
cfg = getattr(Cfg, size)()
ws = make_weights(cfg, seed=0, device="cuda", keep_fp=False, parts=("talker", "cp"))
e = TalkerEngine(cfg, ws, "cuda", batch=1, max_frames=64, max_ctx=1024)
NST = 2048; G = torch.cuda.get_device_properties(0).multi_processor_count
timing = torch.zeros(G * NST, dtype=torch.int64, device="cuda")
e.use_graphs = False; e.x.normal_(0, 0.02); x0 = e.x.clone()
ctx = 300 if size == "full" else 40
e.fa.ll_timing = timing.data_ptr()
for it in range(3):
    timing.zero_(); e.pos.fill_(ctx); e.x.copy_(x0); e._talker_step(True); torch.cuda.synchronize()
t = timing.view(G, NST).cpu()
# attention CTA 0: 9 coarse + 7 fine stamps per layer = 16
names = ["qkv.pro", "qkv.gemv", "att.A(preload)", "att.B(q words)", "att.C(norm+rope)", "att.D(bar)", "att.E(scores)", "att.F(halfmerge)",
         "att.G(bar)", "att.end(merge+st)", "o.pro", "o.gemv", "gu.pro", "gu.gemv", "down.pro", "down.gemv"]
nl = cfg.talker.num_layers
d = (t[0, 1:16 * nl + 1] - t[0, 0:16 * nl]).view(nl, 16).float() / 1e3
steady = d[1:].mean(0)
for n, v in zip(names, steady.tolist()):
    print(f"{n:22s} {v:6.2f} us")
print("layer", float(steady.sum()))
