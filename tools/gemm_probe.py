"""Throughput of the tcgen05 W8 GEMM: HBM GB/s at decode sizes (M=64), TFLOP/s at prefill sizes."""
import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qwen3-tts-apple-silicon_b200"))
from qwen3_tts_b200 import lib as L
from qwen3_tts_b200.weights import pack_w8
lib = L.load()
dev = "cuda"
def blob(n, k):
    q = torch.randint(0, 256, (n, k), device=dev, dtype=torch.uint8)
    s = (torch.rand(n, k // 64, device=dev) * 1e-3).to(torch.bfloat16); b = (-s.float() * 128).to(torch.bfloat16)
    return pack_w8(q, s, b)
print(torch.cuda.get_device_name(0))
for (n, k) in ((2048, 256), (2048, 512), (2048, 1024), (4096, 2048), (2048, 2048), (12288, 2048), (2048, 6144), (3072, 2048)):
    for m in ((64, 0), (64, 1), (128, 0), (256, 0), (4096, 0), (19200, 0)):
        m, use_split = m
        nrep = 24 if m <= 256 else 4
        ws = [blob(n, k) for _ in range(nrep)]          # distinct weights per launch: nothing served from L2
        x = torch.randn(m, k, device=dev); y = torch.empty(m, n, device=dev); xb = torch.empty(2 * m * k, device=dev, dtype=torch.bfloat16)
        sws = torch.empty(8 * m * n, device=dev); cnt = torch.zeros(1024, device=dev, dtype=torch.int32)
        args = []
        for w in ws:
            a = L.GemmArgs(); o = L.W8(); o.w, o.N, o.K = w.data_ptr(), n, k
            a.w, a.M, a.prologue = o, m, L.PRO_RAW
            a.x, a.x_stride, a.y, a.y_stride, a.xb = x.data_ptr(), k, y.data_ptr(), n, xb.data_ptr()
            if use_split:
                a.splitk_ws, a.splitk_ws_floats, a.splitk_counters = sws.data_ptr(), sws.numel(), cnt.data_ptr()
            args.append(a)
        s = L.stream_ptr()
        for a in args: L.check(lib.q3t_w8_gemm(C.byref(a), s))
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for a in args: L.check(lib.q3t_w8_gemm(C.byref(a), L.stream_ptr()))
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / nrep
        print(f"N={n:6d} K={k:5d} M={m:6d} split={use_split}: {us:9.1f} us/launch (prep+gemm)  {n*k*1.0625/us/1e3:7.0f} GB/s weights  {2.0*m*n*k/us/1e6:7.1f} TFLOP/s")
