"""Where the time of one decode-sized W8 GEMM launch goes: %globaltimer stamps of the CTAs (0, 0, z) of a -DTC_TIMING build
(tools/build_variant_gemm.sh tct -DTC_TIMING; Q3T_LIB=.../libq3tts_b200_tct.so python tools/gemm_stamps.py)."""
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
names = ["start", "setup", "B-prod pdl", "MMA 1st", "MMA issued", "acc ready", "epi done", "end", "staged", "cluster sync"]
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
for (n, k, m, split) in ((4096, 2048, 64, 1), (2048, 2048, 64, 1), (12288, 2048, 64, 1), (2048, 6144, 64, 1), 
                         (3072, 1024, 64, 1), (6144, 1024, 64, 1), (1024, 3072, 64, 1)):
    w = blob(n, k)
    x = torch.randn(m, k, device=dev); y = torch.empty(m, n, device=dev); xb = torch.empty(2 * m * k, device=dev, dtype=torch.bfloat16)
    ws = torch.zeros(8 * m * n + 1024, device=dev); cnt = torch.zeros(1024, device=dev, dtype=torch.int32)
    a = L.GemmArgs(); o = L.W8(); o.w, o.N, o.K = w.data_ptr(), n, k
    a.w, a.M, a.prologue = o, m, L.PRO_RAW
    a.x, a.x_stride, a.y, a.y_stride, a.xb = x.data_ptr(), k, y.data_ptr(), n, xb.data_ptr()
    a.splitk_ws, a.splitk_ws_floats = ws.data_ptr(), (8 * m * n if split else 0)
    a.splitk_counters = cnt.data_ptr() if split else 0
    for it in range(3):
        flush.fill_(it)                  # weights cold in L2, like a layer's weights inside a frame
        L.check(lib.q3t_w8_gemm(C.byref(a), L.stream_ptr()))
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    st = ws[-1024:].view(torch.int64).cpu().tolist()       # [z][16], then the per-K-block detail of CTA (0, 0, 0)
    print(f"N={n} K={k} M={m} splitK={'on' if split else 'off'}")
    t0 = min(st[z * 16] for z in range(8) if st[z * 16])
    for z in range(8):
        r = st[z * 16: z * 16 + 16]
        if not r[0]:
            continue
        print(f"   z={z}: " + "  ".join(f"{nm}={(r[i] - t0) / 1e3:.2f}" for i, nm in enumerate(names) if r[i] > 0 and r[i] >= t0))
    dn = ["free", "raw", "stored", "full", "issued", "B tma"]
    for kb in range(16):
        r = st[128 + kb * 8: 128 + kb * 8 + 6]
        if r[3]:
            print(f"      kb {kb:2d}: " + "  ".join(f"{nm}={(r[i] - t0) / 1e3:.2f}" for i, nm in enumerate(dn) if r[i]))
