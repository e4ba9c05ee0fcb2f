import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qwen3-tts-apple-silicon_b200"))
from qwen3_tts_b200 import lib as L
from qwen3_tts_b200.weights import pack_w8
lib = L.load(); dev = "cuda"
def blob(n, k):
    q = torch.randint(0, 256, (n, k), device=dev, dtype=torch.uint8)
    s = (torch.rand(n, k // 64, device=dev) * 1e-3).to(torch.bfloat16); b = (-s.float() * 128).to(torch.bfloat16)
    return pack_w8(q, s, b)
# decode shapes of the talker (QKV, O, gate/up, down) and of the code predictor (QKV, O, gate/up, down) at 64 tokens with the split-K
# opt-in the engine passes (clusters of 2..8 CTAs per output tile), then a prompt-sized GEMM
ws = torch.zeros(8 * 64 * 12288, device=dev)
for (n, k, m) in ((4096, 2048, 64), (2048, 2048, 64), (12288, 2048, 64), (2048, 6144, 64), (4096, 1024, 64), (1024, 2048, 64), (6144, 1024, 64), (1024, 3072, 64), (4096, 2048, 4096)):
    w = blob(n, k)
    x = torch.randn(m, k, device=dev); y = torch.empty(m, n, device=dev); xb = torch.empty(2 * m * k, device=dev, dtype=torch.bfloat16)
    a = L.GemmArgs(); o = L.W8(); o.w, o.N, o.K = w.data_ptr(), n, k
    a.w, a.M, a.prologue = o, m, L.PRO_RAW
    a.x, a.x_stride, a.y, a.y_stride, a.xb = x.data_ptr(), k, y.data_ptr(), n, xb.data_ptr()
    if m <= 128:
        a.splitk_ws, a.splitk_ws_floats = ws.data_ptr(), ws.numel()
    for _ in range(3):
        L.check(lib.q3t_w8_gemm(C.byref(a), L.stream_ptr()))
    torch.cuda.synchronize()
print("ok")
