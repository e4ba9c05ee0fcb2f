"""Target program for ncu: full-size engine, one warm-up talker step, then (eagerly, no graph) one talker decode
step at ctx=300 and one whole frame.  Filter with -k regex:q3t to see only this repo's kernels."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qwen3-tts-apple-silicon_b200"))
from qwen3_tts_b200 import config as Cfg
from qwen3_tts_b200.engine import TalkerEngine
from qwen3_tts_b200.weights import make_weights

cfg = Cfg.full()
ws = make_weights(cfg, seed=0, device="cuda", keep_fp=False, parts=("talker", "cp"))
e = TalkerEngine(cfg, ws, "cuda", batch=1, max_frames=64, max_ctx=1024)
del ws
e.set_sampling(do_sample=False)
if os.environ.get("Q3T_NCU_TABLES"):   # code-predictor table rows as prefill() builds them in production; off by default because
    e._ensure_cp_proj_rows()           # the 32 k GEMV launches of the build would all be profiled by a w8_gem kernel filter
e.use_graphs = False
e.x.normal_(0, 0.02)
e.pos.fill_(300)
e._talker_step(True)          # warm-up (143 launches)
torch.cuda.synchronize()
e.pos.fill_(300)
e._talker_step(True)          # measured talker step
torch.cuda.synchronize()
e._frame()                    # measured frame (sample + 16 CP passes + next input + talker step)
torch.cuda.synchronize()
print("ok")
