"""Quick on-GPU probe (not a bench): per-shape W8 GEMV bandwidth, talker step and frame time at full size."""
import ctypes as C
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qwen3-tts-apple-silicon_b200"))
from qwen3_tts_b200 import config as Cfg, lib as L
from qwen3_tts_b200.weights import make_weights, TILE_BYTES

lib = L.load()
dev = torch.device("cuda")
print(torch.cuda.get_device_name(0))


def ev_time(fn, iters):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn(); torch.cuda.synchronize()
    s.record()
    for _ in range(iters):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


# ---- per-shape GEMV, cycling through enough distinct matrices to defeat the 126 MB L2
for n, k in [(4096, 2048), (2048, 2048), (12288, 2048), (2048, 6144), (3072, 2048), (4096, 1024), (1024, 2048)]:
    tiles = (n // 16) * (k // 256)
    nbytes = tiles * TILE_BYTES
    copies = max(2, int(400e6 // nbytes) + 1)
    blobs = [torch.randint(0, 255, (nbytes,), dtype=torch.uint8, device=dev) for _ in range(copies)]
    x = torch.randn(1, k, device=dev)
    y = torch.empty(1, n, device=dev)
    args = []
    for b in blobs:
        a = L.GemvArgs()
        a.w.w, a.w.N, a.w.K, a.M, a.prologue = b.data_ptr(), n, k, 1, L.PRO_RAW
        a.x, a.x_stride, a.y, a.y_stride = x.data_ptr(), k, y.data_ptr(), n
        args.append(a)
    def run_all():
        st = L.stream_ptr()
        for a in args:
            lib.q3t_w8_gemv(C.byref(a), st)
    run_all(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        run_all()
    ms = ev_time(g.replay, 20) / copies
    print(f"gemv N={n:6d} K={k:5d}  {ms*1e3:8.2f} us/launch (graph of {copies})  {nbytes/ms/1e6:8.1f} GB/s  ({nbytes/1e6:.1f} MB)")
    del blobs

# ---- full-size talker step / frame
from qwen3_tts_b200.engine import TalkerEngine
t0 = time.time()
cfg = Cfg.full()
ws = make_weights(cfg, seed=0, device="cuda", keep_fp=False, parts=("talker", "cp"))
print("weights on device in", round(time.time() - t0, 1), "s")
e = TalkerEngine(cfg, ws, "cuda", batch=1, max_frames=256, max_ctx=1024)
del ws
torch.cuda.empty_cache()
print("W8 bytes total", e.w_bytes / 1e9, "GB")
e.set_sampling(do_sample=False)
L0 = 90
emb = torch.randn(1, L0, cfg.talker.hidden_size, device=dev) * 0.02
t0 = time.time(); e.prefill(emb); torch.cuda.synchronize(); print("prefill(+graph capture)", round(time.time() - t0, 2), "s")
t0 = time.time(); e.prefill(emb); torch.cuda.synchronize(); print("prefill 90 tok", round((time.time() - t0) * 1e3, 1), "ms")
print("launches per frame", e.launches_per_frame)
ms = ev_time(lambda: e._run("step_logits"), 50)
e.pos.fill_(L0)
talker_bytes = sum(1 for _ in [0]) and (28 * (4096 * 2048 + 2048 * 2048 + 12288 * 2048 + 2048 * 6144) + 3072 * 2048) * 1.0625
print(f"talker step (graph): {ms*1e3:.1f} us -> {talker_bytes/ms/1e6:.0f} GB/s algorithmic")
e.use_graphs = False
ms2 = ev_time(lambda: e._run("step_logits"), 20)
print(f"talker step (eager launches): {ms2*1e3:.1f} us")
e.use_graphs = True
e.prefill(emb)
ms = ev_time(lambda: e._run("frame"), 50)
print(f"frame (graph): {ms*1e3:.1f} us -> RTFx {80.0/ms:.1f}")
