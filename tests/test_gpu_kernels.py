"""GPU parity of the individual kernels against the CPU oracle, called through the C ABI (ctypes)."""
import ctypes as C

import pytest
import torch

from oracle import qwen3_tts_oracle as O
from qwen3_tts_b200 import lib as L
from qwen3_tts_b200.weights import dequantize_w8, pack_w8, quantize_w8

pytestmark = pytest.mark.gpu


def _w8(n, k, seed, dev):
    g = torch.Generator().manual_seed(seed)
    w = torch.randn(n, k, generator=g) * 0.02
    q, s, b = quantize_w8(w)
    wd = dequantize_w8(q, s, b)
    o = L.W8()
    blob = pack_w8(q.to(dev), s.to(dev), b.to(dev))
    o.w, o.N, o.K = blob.data_ptr(), n, k
    return o, blob, wd


@pytest.mark.parametrize("n,k", [(2048, 2048), (4096, 2048), (12288, 2048), (2048, 6144), (3072, 2048), (1024, 256),
                                 (1024, 1024), (6144, 1024), (1024, 3072), (16, 256), (2064, 512)])
@pytest.mark.parametrize("m", [1, 2])
def test_w8_gemv_raw(cuda, n, k, m):
    lib = L.load()
    o, blob, wd = _w8(n, k, n + k, cuda)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(m, k, generator=g) * torch.rand(1, generator=g) * 3
    ref = (x.double() @ wd.double().T)
    xd = x.to(cuda)
    y = torch.full((m, n), float("nan"), device=cuda)
    a = L.GemvArgs()
    a.w, a.M, a.prologue = o, m, L.PRO_RAW
    a.x, a.x_stride, a.y, a.y_stride = xd.data_ptr(), k, y.data_ptr(), n
    L.check(lib.q3t_w8_gemv(C.byref(a), L.stream_ptr()), "gemv")
    torch.cuda.synchronize()
    err = (y.cpu().double() - ref).abs().max() / ref.abs().max()
    assert err < 2e-6, f"rel err {err}"


def _bf16(x):
    return x.to(torch.bfloat16).double()


@pytest.mark.parametrize("m,n,k", [(1, 128, 256), (5, 256, 512), (64, 2048, 2048), (64, 4096, 2048), (100, 2048, 6144),
                                   (300, 1024, 1024), (64, 3072, 2048), (17, 1024, 256)])
def test_w8_gemm_tcgen05_raw(cuda, m, n, k):
    """tcgen05/TMEM W8 GEMM with split-bf16 operands (hi + lo, three MMAs per K block): within 5e-5 of the UNROUNDED fp64
    product - a single bf16 rounding of the operands would be 200x worse (and 1.4e-2 on the logits of the 28-layer stack,
    see tests/test_gpu_fullsize.py)."""
    lib = L.load()
    o, blob, wd = _w8(n, k, n + k + m, cuda)
    g = torch.Generator().manual_seed(3 * m + 1)
    x = torch.randn(m, k, generator=g) * 0.7
    ref_bf = _bf16(x) @ _bf16(wd).T
    ref = x.double() @ wd.double().T
    xd = x.to(cuda)
    y = torch.full((m, n), float("nan"), device=cuda)
    xb = torch.empty(2 * m * k, device=cuda, dtype=torch.bfloat16)       # split rows [hi(K) | lo(K)]
    a = L.GemmArgs()
    a.w, a.M, a.prologue = o, m, L.PRO_RAW
    a.x, a.x_stride, a.y, a.y_stride, a.xb = xd.data_ptr(), k, y.data_ptr(), n, xb.data_ptr()
    L.check(lib.q3t_w8_gemm(C.byref(a), L.stream_ptr()), "gemm")
    torch.cuda.synchronize()
    yd = y.cpu().double()
    assert torch.isfinite(yd).all()
    assert (yd - ref).abs().max() / ref.abs().max() < 5e-5, "split-bf16 operands: expected ~1e-5 of the exact product"
    assert (yd - ref).abs().max() < (yd - ref_bf).abs().max() / 20, "not better than single-bf16 operands?"
    # split-K (decode-sized problems): partial tiles summed in split order by the last CTA to arrive; twice, to check
    # that the arrival counters re-arm themselves and that the result is bit-reproducible
    ws = torch.empty(8 * m * n, device=cuda)
    cnt = torch.zeros(1024, device=cuda, dtype=torch.int32)
    a.splitk_ws, a.splitk_ws_floats, a.splitk_counters = ws.data_ptr(), ws.numel(), cnt.data_ptr()
    outs = []
    for _ in range(2):
        y.fill_(float("nan"))
        L.check(lib.q3t_w8_gemm(C.byref(a), L.stream_ptr()), "gemm split-K")
        torch.cuda.synchronize()
        outs.append(y.clone())
    assert torch.equal(outs[0], outs[1])
    assert int(cnt.abs().sum()) == 0
    assert (outs[0].cpu().double() - ref).abs().max() / ref.abs().max() < 5e-5, "split-K"


def test_w8_gemm_tcgen05_prologues_epilogues(cuda):
    lib = L.load()
    n, k, m = 2048, 1024, 37
    o, blob, wd = _w8(n, k, 5, cuda)
    g = torch.Generator().manual_seed(12)
    # RMSNorm prologue + bias + residual (in place)
    x = torch.randn(m, k, generator=g)
    nw = 1 + 0.1 * torch.randn(k, generator=g)
    bias = torch.randn(n, generator=g)
    resid = torch.randn(m, n, generator=g)
    ref = O.rms_norm(x, nw, 1e-6).double() @ wd.double().T + bias.double() + resid.double()
    xd, nwd, bd, rd = x.to(cuda), nw.to(cuda), bias.to(cuda), resid.to(cuda)
    xb = torch.empty(m * 2 * k, device=cuda, dtype=torch.bfloat16)
    o.lin_bias = bd.data_ptr()
    a = L.GemmArgs()
    a.w, a.M, a.prologue = o, m, L.PRO_RMSNORM
    a.x, a.x_stride, a.norm_w, a.eps = xd.data_ptr(), k, nwd.data_ptr(), 1e-6
    a.resid, a.resid_stride, a.y, a.y_stride, a.xb = rd.data_ptr(), n, rd.data_ptr(), n, xb.data_ptr()
    L.check(lib.q3t_w8_gemm(C.byref(a), L.stream_ptr()), "gemm")
    torch.cuda.synchronize()
    assert (rd.cpu().double() - ref).abs().max() / ref.abs().max() < 1e-4
    # SwiGLU prologue (interleaved gate/up input) + SiLU epilogue
    o.lin_bias = 0
    gu = torch.randn(m, 2 * k, generator=g)
    act = torch.nn.functional.silu(gu[:, :k]) * gu[:, k:]
    ref = torch.nn.functional.silu(act.double() @ wd.double().T)
    gud = torch.cat([gu[:, :k].reshape(m, -1, 8), gu[:, k:].reshape(m, -1, 8)], 2).reshape(m, 2 * k).contiguous().to(cuda)
    y = torch.empty(m, n, device=cuda)
    a = L.GemmArgs()
    a.w, a.M, a.prologue, a.act = o, m, L.PRO_SWIGLU, L.ACT_SILU
    a.x, a.x_stride, a.y, a.y_stride, a.xb = gud.data_ptr(), 2 * k, y.data_ptr(), n, xb.data_ptr()
    L.check(lib.q3t_w8_gemm(C.byref(a), L.stream_ptr()), "gemm")
    torch.cuda.synchronize()
    assert (y.cpu().double() - ref).abs().max() / ref.abs().max() < 1e-4
    # SwiGLU OUTPUT epilogue: fused gate/up matrix with rows interleaved in blocks of 8
    half = n // 2
    wg, wu = wd[:half], wd[half:]
    idx = torch.arange(half).view(-1, 8)
    perm = torch.cat([idx, idx + half], 1).reshape(-1)
    q, s, b = quantize_w8(wd)            # wd is already on the W8 grid: quantising again is (nearly) lossless
    wd2 = dequantize_w8(q, s, b)
    o2 = L.W8()
    blob2 = pack_w8(q[perm].contiguous().to(cuda), s[perm].contiguous().to(cuda), b[perm].contiguous().to(cuda))
    o2.w, o2.N, o2.K = blob2.data_ptr(), n, k
    xr = torch.randn(m, k, generator=g)
    gate = xr.double() @ wd2[:half].double().T
    up = xr.double() @ wd2[half:].double().T
    ref = torch.nn.functional.silu(gate) * up
    y = torch.empty(m, half, device=cuda)
    xrd = xr.to(cuda)
    a = L.GemmArgs()
    a.w, a.M, a.prologue, a.swiglu_out = o2, m, L.PRO_RAW, 1
    a.x, a.x_stride, a.y, a.y_stride, a.xb = xrd.data_ptr(), k, y.data_ptr(), half, xb.data_ptr()
    L.check(lib.q3t_w8_gemm(C.byref(a), L.stream_ptr()), "gemm")
    torch.cuda.synchronize()
    assert (y.cpu().double() - ref).abs().max() / ref.abs().max() < 1e-4


@pytest.mark.parametrize("m,splitk", [(64, True), (37, False), (300, False)])
def test_w8_gemm_deferred_rmsnorm_pair(cuda, m, splitk):
    """Two GEMMs of a residual stream with the RMSNorm between them deferred (q3t_gemm_args.y_norm_w / x_rowss): the first
    writes y = W1 a + resid (fp32), split rows of y * w and per-128-feature sums of y^2; the second contracts those rows and
    scales by rsqrt(mean(y^2) + eps).  Against the fp64 arithmetic of the oracle's RMSNorm (O.rms_norm), plain and SwiGLU."""
    lib = L.load()
    hid, kin, n2 = 1024, 2048, 2048
    o1, blob1, wd1 = _w8(hid, kin, 21, cuda)
    g = torch.Generator().manual_seed(100 + m)
    a_in = torch.randn(m, kin, generator=g) * 0.5
    resid = torch.randn(m, hid, generator=g)
    nw = 1 + 0.2 * torch.randn(hid, generator=g)
    y_ref = a_in.double() @ wd1.double().T + resid.double()
    yn_ref = (y_ref * torch.rsqrt((y_ref * y_ref).mean(-1, keepdim=True) + 1e-6)) * nw.double()
    o2, blob2, wd2 = _w8(n2, hid, 22, cuda)
    z_ref = yn_ref @ wd2.double().T
    # the fused gate/up flavour of the consumer (rows interleaved in blocks of 8)
    half = n2 // 2
    idx = torch.arange(half).view(-1, 8)
    perm = torch.cat([idx, idx + half], 1).reshape(-1)
    q, sc, bi = quantize_w8(wd2)
    wd2b = dequantize_w8(q, sc, bi)
    o3 = L.W8()
    blob3 = pack_w8(q[perm].contiguous().to(cuda), sc[perm].contiguous().to(cuda), bi[perm].contiguous().to(cuda))
    o3.w, o3.N, o3.K = blob3.data_ptr(), n2, hid
    sw_ref = torch.nn.functional.silu(yn_ref @ wd2b[:half].double().T) * (yn_ref @ wd2b[half:].double().T)

    xd, rd, nwd = a_in.to(cuda), resid.to(cuda), nw.to(cuda)
    xb = torch.empty(2 * m * kin, device=cuda, dtype=torch.bfloat16)
    yb = torch.full((m, 2 * hid), float("nan"), device=cuda, dtype=torch.bfloat16)
    rowss = torch.full((m, hid // 128), float("nan"), device=cuda)
    ws = torch.empty(8 * m * max(hid, n2), device=cuda)
    cnt = torch.zeros(1024, device=cuda, dtype=torch.int32)
    a = L.GemmArgs()
    a.w, a.M, a.prologue = o1, m, L.PRO_RAW
    a.x, a.x_stride, a.xb = xd.data_ptr(), kin, xb.data_ptr()
    a.resid, a.resid_stride, a.y, a.y_stride = rd.data_ptr(), hid, rd.data_ptr(), hid            # in place, like the engine
    a.y_bf16, a.y_norm_w, a.y_rowss = yb.data_ptr(), nwd.data_ptr(), rowss.data_ptr()
    if splitk:
        a.splitk_ws, a.splitk_ws_floats, a.splitk_counters = ws.data_ptr(), ws.numel(), cnt.data_ptr()
    L.check(lib.q3t_w8_gemm(C.byref(a), L.stream_ptr()), "gemm producer")
    torch.cuda.synchronize()
    y = rd.cpu().double()
    assert (y - y_ref).abs().max() / y_ref.abs().max() < 5e-5
    ss = rowss.cpu().double().sum(-1)
    assert ((ss - (y * y).sum(-1)).abs() / (y * y).sum(-1)).max() < 1e-5
    rows = yb.float().cpu().double()
    assert ((rows[:, :hid] + rows[:, hid:]) - y * nw.double()).abs().max() / (y * nw.double()).abs().max() < 2e-5, "split rows of y * w"

    z = torch.full((m, n2), float("nan"), device=cuda)
    b = L.GemmArgs()
    b.w, b.M, b.prologue, b.eps = o2, m, L.PRO_RAW, 1e-6
    b.x_bf16, b.x_rowss, b.x_rowss_parts, b.xb = yb.data_ptr(), rowss.data_ptr(), hid // 128, xb.data_ptr()
    b.y, b.y_stride = z.data_ptr(), n2
    if splitk:
        b.splitk_ws, b.splitk_ws_floats, b.splitk_counters = ws.data_ptr(), ws.numel(), cnt.data_ptr()
    L.check(lib.q3t_w8_gemm(C.byref(b), L.stream_ptr()), "gemm consumer")
    torch.cuda.synchronize()
    assert (z.cpu().double() - z_ref).abs().max() / z_ref.abs().max() < 1e-4

    sw = torch.full((m, half), float("nan"), device=cuda)
    b.w, b.swiglu_out, b.y, b.y_stride = o3, 1, sw.data_ptr(), half
    L.check(lib.q3t_w8_gemm(C.byref(b), L.stream_ptr()), "gemm consumer swiglu")
    torch.cuda.synchronize()
    assert (sw.cpu().double() - sw_ref).abs().max() / sw_ref.abs().max() < 1e-4
    # the same pair through the prologue launch (act_prep_kernel) agrees to the rounding of the operands
    z2 = torch.empty(m, n2, device=cuda)
    c = L.GemmArgs()
    c.w, c.M, c.prologue = o2, m, L.PRO_RMSNORM
    c.x, c.x_stride, c.norm_w, c.eps, c.xb = rd.data_ptr(), hid, nwd.data_ptr(), 1e-6, xb.data_ptr()
    c.y, c.y_stride = z2.data_ptr(), n2
    L.check(lib.q3t_w8_gemm(C.byref(c), L.stream_ptr()), "gemm prologue")
    torch.cuda.synchronize()
    assert (z2 - z).abs().max() / z.abs().max() < 5e-5
    assert int(cnt.abs().sum()) == 0


def test_w8_gemv_prologues_epilogues(cuda):
    lib = L.load()
    n, k = 2048, 1024
    o, blob, wd = _w8(n, k, 3, cuda)
    g = torch.Generator().manual_seed(11)
    # RMSNorm prologue + bias + residual
    x = torch.randn(2, k, generator=g)
    nw = 1 + 0.1 * torch.randn(k, generator=g)
    bias = torch.randn(n, generator=g)
    resid = torch.randn(2, n, generator=g)
    ref = O.rms_norm(x, nw, 1e-6).double() @ wd.double().T + bias.double() + resid.double()
    xd, nwd, bd, rd = x.to(cuda), nw.to(cuda), bias.to(cuda), resid.to(cuda)
    o.lin_bias = bd.data_ptr()
    a = L.GemvArgs()
    a.w, a.M, a.prologue = o, 2, L.PRO_RMSNORM
    a.x, a.x_stride, a.norm_w, a.eps = xd.data_ptr(), k, nwd.data_ptr(), 1e-6
    a.resid, a.resid_stride, a.y, a.y_stride = rd.data_ptr(), n, rd.data_ptr(), n     # in place on the residual
    L.check(lib.q3t_w8_gemv(C.byref(a), L.stream_ptr()), "gemv")
    torch.cuda.synchronize()
    assert (rd.cpu().double() - ref).abs().max() / ref.abs().max() < 3e-6
    # SwiGLU prologue + SiLU epilogue
    o.lin_bias = 0
    gu = torch.randn(1, 2 * k, generator=g)
    ref = torch.nn.functional.silu(((torch.nn.functional.silu(gu[:, :k]) * gu[:, k:]).double() @ wd.double().T))
    # device layout of a fused gate/up output: blocks of 8 gate values followed by the matching 8 up values
    gud = torch.cat([gu[:, :k].reshape(1, -1, 8), gu[:, k:].reshape(1, -1, 8)], 2).reshape(1, 2 * k).contiguous().to(cuda)
    y = torch.empty(1, n, device=cuda)
    a = L.GemvArgs()
    a.w, a.M, a.prologue, a.act = o, 1, L.PRO_SWIGLU, L.ACT_SILU
    a.x, a.x_stride, a.y, a.y_stride = gud.data_ptr(), 2 * k, y.data_ptr(), n
    L.check(lib.q3t_w8_gemv(C.byref(a), L.stream_ptr()), "gemv")
    torch.cuda.synchronize()
    assert (y.cpu().double() - ref).abs().max() / ref.abs().max() < 3e-6
    # gather prologue (embedding row picked on device) with a zero row
    table = torch.randn(50, k, generator=g)
    table[7] = 0
    idx = torch.tensor([33, 7], dtype=torch.int32)
    ref = table[idx.long()].double() @ wd.double().T
    td, idd = table.to(cuda), idx.to(cuda)
    y = torch.empty(2, n, device=cuda)
    a = L.GemvArgs()
    a.w, a.M, a.prologue = o, 2, L.PRO_RAW
    a.x, a.x_stride, a.gather_idx, a.gather_idx_stride, a.gather_row_stride = td.data_ptr(), 0, idd.data_ptr(), 1, k
    a.y, a.y_stride = y.data_ptr(), n
    L.check(lib.q3t_w8_gemv(C.byref(a), L.stream_ptr()), "gemv")
    torch.cuda.synchronize()
    assert (y.cpu().double() - ref).abs().max() / ref.abs().max() < 3e-6
    assert (y[1] == 0).all()


def test_w8_gemv_rejects_bad_shapes(cuda):
    lib = L.load()
    a = L.GemvArgs()
    a.w.N, a.w.K, a.M = 24, 256, 1
    assert lib.q3t_w8_gemv(C.byref(a), L.stream_ptr()) != 0
    assert b"N%16" in lib.q3t_last_error()
    a.w.N, a.M = 32, 3
    assert lib.q3t_w8_gemv(C.byref(a), L.stream_ptr()) != 0


def _attn_ref(qkv, qn, kn, eps, theta, H, Hkv, D, kcache, vcache, pos):
    """fp32 reference for one new token; kcache/vcache [pos, Hkv, D] already rounded to bf16."""
    q = qkv[: H * D].view(H, D)
    k = qkv[H * D:(H + Hkv) * D].view(Hkv, D)
    v = qkv[(H + Hkv) * D:].view(Hkv, D)
    cos, sin = O.rope_cos_sin(torch.tensor([pos]), D, theta)
    q = O.apply_rope(O.rms_norm(q, qn, eps)[None], cos, sin)[0]
    k = O.apply_rope(O.rms_norm(k, kn, eps)[None], cos, sin)[0]
    k, v = k.bfloat16().float(), v.bfloat16().float()
    K = torch.cat([kcache, k[None]], 0).repeat_interleave(H // Hkv, 1)
    V = torch.cat([vcache, v[None]], 0).repeat_interleave(H // Hkv, 1)
    s = torch.einsum("hd,shd->hs", q, K) * D ** -0.5
    return torch.einsum("hs,shd->hd", torch.softmax(s, -1), V).reshape(-1), k, v


@pytest.mark.parametrize("pos,nsplit", [(0, 1), (0, 16), (5, 4), (16, 16), (17, 1), (250, 16), (1000, 16), (1000, 3), (300, 8), (40, 2), (299, 4)])
def test_attn_decode(cuda, pos, nsplit):
    lib = L.load()
    H, Hkv, D, B, theta, eps = 4, 2, 128, 2, 1e6, 1e-6
    g = torch.Generator().manual_seed(pos * 31 + nsplit)
    max_pages = (pos + 1 + 15) // 16 + 1
    n_pages = B * max_pages
    pool = torch.zeros(n_pages, 2, Hkv, 16, D, dtype=torch.bfloat16)
    perm = torch.randperm(n_pages, generator=g).int().reshape(B, max_pages)      # scattered pages
    kc = torch.randn(B, pos, Hkv, D, generator=g).bfloat16()
    vc = torch.randn(B, pos, Hkv, D, generator=g).bfloat16()
    for b in range(B):
        for t in range(pos):
            pool[perm[b, t // 16], 0, :, t % 16] = kc[b, t]
            pool[perm[b, t // 16], 1, :, t % 16] = vc[b, t]
    qkv = torch.randn(B, (H + 2 * Hkv) * D, generator=g)
    qn = 1 + 0.1 * torch.randn(D, generator=g)
    kn = 1 + 0.1 * torch.randn(D, generator=g)
    inv = 1.0 / (theta ** (torch.arange(0, D, 2, dtype=torch.float32) / D))
    d = lambda t: t.to(cuda).contiguous()
    pool_d, perm_d, qkv_d, qn_d, kn_d, inv_d = map(d, (pool, perm, qkv, qn, kn, inv))
    pos_d = torch.full((B,), pos, dtype=torch.int32, device=cuda)
    out = torch.full((B, H * D), float("nan"), device=cuda)
    work = torch.zeros(B * Hkv * nsplit * (H // Hkv) * (D + 2), device=cuda)
    cnt = torch.zeros(B * Hkv, dtype=torch.int32, device=cuda)
    a = L.AttnArgs()
    a.qkv, a.q_norm_w, a.k_norm_w, a.eps, a.inv_freq = qkv_d.data_ptr(), qn_d.data_ptr(), kn_d.data_ptr(), eps, inv_d.data_ptr()
    a.kv_pool, a.block_tbl, a.max_pages, a.pos = pool_d.data_ptr(), perm_d.data_ptr(), max_pages, pos_d.data_ptr()
    a.out, a.work, a.counters = out.data_ptr(), work.data_ptr(), cnt.data_ptr()
    a.B, a.H, a.Hkv, a.D, a.nsplit = B, H, Hkv, D, nsplit
    for _ in range(2):   # second launch checks the counters re-arm themselves
        L.check(lib.q3t_attn_decode(C.byref(a), L.stream_ptr()), "attn")
    torch.cuda.synchronize()
    assert int(cnt.abs().sum()) == 0
    pool_h = pool_d.cpu()
    for b in range(B):
        ref, k_new, v_new = _attn_ref(qkv[b], qn, kn, eps, theta, H, Hkv, D, kc[b].float(), vc[b].float(), pos)
        err = (out[b].cpu() - ref).abs().max() / ref.abs().max()
        assert err < 2e-5, f"b={b} rel err {err}"
        pg = perm[b, pos // 16]
        # the new K/V row landed in its page (bf16): allow one bf16 ulp for rounding-boundary cases
        assert (pool_h[pg, 0, :, pos % 16].float() - k_new).abs().max() <= 2 ** -7 * k_new.abs().max()
        assert torch.equal(pool_h[pg, 1, :, pos % 16].float(), v_new)


@pytest.mark.parametrize("lens", [(1,), (33, 7, 64), (300, 17), (129, 32, 31, 65)])
def test_attn_prefill_ragged_causal(cuda, lens):
    """Tensor-core prompt attention (csrc/attn_prefill.cu): ragged rows of several sequences, scattered pages, blocks of <= 32
    rows.  Reference: fp32 softmax attention over the bf16 K/V that are in the cache; q is rounded to bf16 by the kernel, so
    the tolerance is the bf16 one of BASELINE.json (1e-2 relative), measured per row."""
    lib = L.load()
    H, Hkv, D, theta, eps = 4, 2, 128, 1e6, 1e-6
    g = torch.Generator().manual_seed(sum(lens))
    nseq, M = len(lens), sum(lens)
    max_pages = (max(lens) + 15) // 16 + 1
    n_pages = nseq * max_pages
    perm = torch.randperm(n_pages, generator=g).int().reshape(nseq, max_pages)
    qkv = torch.randn(M, (H + 2 * Hkv) * D, generator=g)
    qn = 1 + 0.1 * torch.randn(D, generator=g)
    kn = 1 + 0.1 * torch.randn(D, generator=g)
    inv = 1.0 / (theta ** (torch.arange(0, D, 2, dtype=torch.float32) / D))
    pos = torch.cat([torch.arange(l, dtype=torch.int32) for l in lens])
    seq = torch.cat([torch.full((l,), i, dtype=torch.int32) for i, l in enumerate(lens)])
    blocks, r0 = [], 0
    for l in lens:
        blocks += [(r0 + o, min(32, l - o)) for o in range(0, l, 32)]
        r0 += l
    d = lambda t: t.to(cuda).contiguous()
    pool_d = torch.zeros(n_pages, 2, Hkv, 16, D, dtype=torch.bfloat16, device=cuda)
    perm_d, qkv_d, qn_d, kn_d, inv_d, pos_d, seq_d = map(d, (perm, qkv, qn, kn, inv, pos, seq))
    blk_d = torch.tensor(blocks, dtype=torch.int32, device=cuda)
    # pass 1 with the decode kernel: K/V (normed, rotated, bf16) of every row into the cache
    a = L.AttnArgs()
    a.qkv, a.q_norm_w, a.k_norm_w, a.eps, a.inv_freq = qkv_d.data_ptr(), qn_d.data_ptr(), kn_d.data_ptr(), eps, inv_d.data_ptr()
    a.kv_pool, a.block_tbl, a.max_pages, a.pos = pool_d.data_ptr(), perm_d.data_ptr(), max_pages, pos_d.data_ptr()
    out1 = torch.zeros(M, H * D, device=cuda)
    work = torch.zeros(M * Hkv * (H // Hkv) * (D + 2), device=cuda)
    cnt = torch.zeros(M * Hkv, dtype=torch.int32, device=cuda)
    a.out, a.work, a.counters = out1.data_ptr(), work.data_ptr(), cnt.data_ptr()
    a.B, a.H, a.Hkv, a.D, a.nsplit, a.mode, a.seq_of_row = M, H, Hkv, D, 1, 1, seq_d.data_ptr()
    L.check(lib.q3t_attn_decode(C.byref(a), L.stream_ptr()), "attn pass 1")
    # reference: the per-row decode kernel in attention-only mode (fp32 q) ...
    a.mode = 2
    L.check(lib.q3t_attn_decode(C.byref(a), L.stream_ptr()), "attn pass 2 (per row)")
    # ... against the tensor-core kernel, fp32 and bf16 outputs
    u = L.AttnPrefillArgs()
    u.qkv, u.q_norm_w, u.eps, u.inv_freq = qkv_d.data_ptr(), qn_d.data_ptr(), eps, inv_d.data_ptr()
    u.kv_pool, u.block_tbl, u.max_pages = pool_d.data_ptr(), perm_d.data_ptr(), max_pages
    u.pos, u.seq_of_row, u.blocks, u.n_blocks = pos_d.data_ptr(), seq_d.data_ptr(), blk_d.data_ptr(), len(blocks)
    out2 = torch.full((M, H * D), float("nan"), device=cuda)
    u.out, u.H, u.Hkv, u.D = out2.data_ptr(), H, Hkv, D
    L.check(lib.q3t_attn_prefill(C.byref(u), L.stream_ptr()), "attn_prefill")
    out3 = torch.zeros(M, 2 * H * D, device=cuda, dtype=torch.bfloat16)      # split rows [hi | lo]
    u.out, u.out_bf16 = 0, out3.data_ptr()
    L.check(lib.q3t_attn_prefill(C.byref(u), L.stream_ptr()), "attn_prefill bf16")
    torch.cuda.synchronize()
    assert bool(torch.isfinite(out2).all())
    ref = out1.cpu().double()
    err = (out2.cpu().double() - ref).abs().amax(1) / ref.abs().amax(1)
    assert float(err.max()) < 1e-2, f"worst row {int(err.argmax())}: rel err {float(err.max()):.3e}"
    # ... and against the ORACLE's attention arithmetic (oracle/qwen3_tts_oracle.py DecoderStack.forward: q/k RMSNorm, RoPE
    # in the rotate_half convention, K/V rounded to bf16 where the cache stores them, fp32 softmax over the causal prefix),
    # evaluated in fp64 per sequence - not against another CUDA kernel
    r0 = 0
    for l in lens:
        x = qkv[r0:r0 + l].double()
        q = x[:, :H * D].view(l, H, D); k = x[:, H * D:(H + Hkv) * D].view(l, Hkv, D); v = x[:, (H + Hkv) * D:].view(l, Hkv, D)
        cos, sin = O.rope_cos_sin(torch.arange(l), D, theta)
        rms = lambda t, w: w.double() * (t * torch.rsqrt(t.pow(2).mean(-1, keepdim=True) + eps))
        q = O.apply_rope(rms(q, qn), cos.double(), sin.double())
        k = O.apply_rope(rms(k, kn), cos.double(), sin.double()).float().bfloat16().double()
        v = v.float().bfloat16().double()
        K, V = k.repeat_interleave(H // Hkv, 1), v.repeat_interleave(H // Hkv, 1)
        sc = torch.einsum("thd,shd->hts", q, K) * D ** -0.5
        sc = sc.masked_fill(~(torch.arange(l)[None, :] <= torch.arange(l)[:, None])[None], float("-inf"))
        want = torch.einsum("hts,shd->thd", torch.softmax(sc, -1), V).reshape(l, H * D)
        for name, got in (("attn_prefill", out2), ("attn_decode", out1)):
            e_ = (got[r0:r0 + l].cpu().double() - want).abs().amax(1) / want.abs().amax(1)
            assert float(e_.max()) < (1e-2 if name == "attn_prefill" else 1e-3), f"{name} vs oracle attention: {float(e_.max()):.3e}"
        r0 += l
    hi = out2.bfloat16()
    assert torch.equal(out3[:, :H * D], hi) and torch.equal(out3[:, H * D:], (out2 - hi.float()).bfloat16())
    # pass 1 folded into the same call (k_norm_w given): the cache must come out bit-identical to the decode kernel's mode 1
    pool2 = torch.zeros_like(pool_d)
    out4 = torch.full((M, H * D), float("nan"), device=cuda)
    u.kv_pool, u.out, u.out_bf16, u.k_norm_w, u.M = pool2.data_ptr(), out4.data_ptr(), 0, kn_d.data_ptr(), M
    L.check(lib.q3t_attn_prefill(C.byref(u), L.stream_ptr()), "attn_prefill with K/V write")
    torch.cuda.synchronize()
    assert torch.equal(pool2, pool_d) and torch.equal(out4, out2)


@pytest.mark.parametrize("B,T,C,taps", [(1, 1, 96, 7), (2, 255, 96, 7), (3, 1000, 96, 7), (1, 513, 32, 3)])
def test_conv_out_clamp_matches_causal_conv1d(cuda, B, T, C, taps):
    """Output convolution of the vocoder (C -> 1, causal) + clamp against torch conv1d on the left-padded signal."""
    lib = L.load()
    g = torch.Generator().manual_seed(B * 1000 + T)
    act = torch.randn(B, T, C, generator=g)
    w = torch.randn(1, C, taps, generator=g) * 0.2           # nn.Conv1d layout [Cout, Cin, k]
    bias = torch.randn(1, generator=g) * 0.1
    ref = torch.nn.functional.conv1d(torch.nn.functional.pad(act.transpose(1, 2).double(), (taps - 1, 0)), w.double(), bias.double())[:, 0]
    ref = ref.clamp(-1, 1)
    act_d, w_d, b_d = act.to(cuda), w[0].t().contiguous().to(cuda), bias.to(cuda)
    out = torch.full((B, T), float("nan"), device=cuda)
    L.check(lib.q3t_conv_out_clamp(act_d.data_ptr(), B, T, C, w_d.data_ptr(), b_d.data_ptr(), taps, out.data_ptr(), L.stream_ptr()), "conv_out")
    torch.cuda.synchronize()
    assert float((out.cpu().double() - ref).abs().max()) < 2e-5
    assert float(out.max()) <= 1.0 and float(out.min()) >= -1.0 and bool((ref.abs() == 1).any()) == bool((out.abs() == 1).any())


def _sample(lib, cuda, logits, sp, seen=None, step=0, uniforms=None):
    B, V = logits.shape
    lg = logits.to(cuda).contiguous()
    out = torch.full((B,), -1, dtype=torch.int32, device=cuda)
    done = torch.zeros(B, dtype=torch.int32, device=cuda)
    st = torch.tensor([step], dtype=torch.int32, device=cuda)
    a = L.SampleArgs()
    a.logits, a.B, a.V, a.logits_stride, a.sp = lg.data_ptr(), B, V, V, sp
    a.seen = 0 if seen is None else seen.data_ptr()
    a.step = st.data_ptr()
    un = None if uniforms is None else uniforms.to(cuda)
    a.uniforms = 0 if un is None else un.data_ptr()
    a.out, a.out_stride, a.done = out.data_ptr(), 1, done.data_ptr()
    L.check(lib.q3t_sample(C.byref(a), L.stream_ptr()), "sample")
    torch.cuda.synchronize()
    return out.cpu(), done.cpu()


def test_sampler_greedy_masks_and_ties(cuda):
    lib = L.load()
    V = 3072
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(4, V, generator=g)
    logits[0, 2900] = 50.0                 # inside the suppressed range -> must be ignored
    logits[1, 2150] = 60.0                 # eos is exempt from suppression
    logits[2, 100] = logits[2, 200] = 40.0  # tie -> lowest index
    logits[3, 2150] = 60.0                 # eos but min_new_tokens forbids it at step 0
    sp = L.Sampling()
    sp.do_sample, sp.repetition_penalty, sp.suppress_lo, sp.suppress_hi, sp.eos_id = 0, 1.0, V - 1024, V, 2150
    out, done = _sample(lib, cuda, logits, sp)
    osp = O.SamplingParams(suppress_lo=V - 1024, suppress_hi=V, eos_id=2150)
    ref = [O.draw(O.process_logits(logits[b], osp, (), 0), osp) for b in range(4)]
    assert out.tolist() == ref and out[1] == 2150 and out[2] == 100 and done.tolist() == [0, 1, 0, 1]
    sp.min_new_tokens = 2
    out, done = _sample(lib, cuda, logits[3:], sp, step=0)
    osp.min_new_tokens = 2
    assert out.tolist() == [O.draw(O.process_logits(logits[3], osp, (), 0), osp)] and done.tolist() == [0]


def test_sampler_repetition_penalty_set_semantics(cuda):
    lib = L.load()
    V = 2048
    g = torch.Generator().manual_seed(1)
    logits = torch.randn(2, V, generator=g)
    logits[0, 5], logits[0, 9] = 10.0, 9.9       # 5 was generated before: 10/1.05 < 9.9 -> 9 wins
    logits[1, 5], logits[1, 9] = -0.001, -0.00101  # negative scores are multiplied
    hist = [5, 5, 77]
    seen = torch.zeros(2, V // 32, dtype=torch.int32)
    for h in hist:
        seen[:, h // 32] |= 1 << (h % 32)
    sp = L.Sampling()
    sp.do_sample, sp.repetition_penalty, sp.suppress_lo, sp.suppress_hi, sp.eos_id = 0, 1.05, -1, -1, -1
    seen_d = seen.to(cuda)
    out, _ = _sample(lib, cuda, logits, sp, seen=seen_d)
    osp = O.SamplingParams(repetition_penalty=1.05)
    ref = [O.draw(O.process_logits(logits[b], osp, hist, 3), osp) for b in range(2)]
    assert out.tolist() == ref
    s2 = seen_d.cpu()
    for b in range(2):   # the chosen id was added to the set
        assert (int(s2[b, out[b] // 32]) >> (int(out[b]) % 32)) & 1


@pytest.mark.parametrize("top_k,top_p,temp", [(50, 1.0, 0.9), (50, 0.8, 0.9), (0, 0.9, 1.0), (1, 1.0, 1.0), (5, 0.5, 2.0)])
def test_sampler_stochastic_with_shared_uniforms(cuda, top_k, top_p, temp):
    lib = L.load()
    V, B = 2048, 64
    g = torch.Generator().manual_seed(top_k + 3)
    logits = torch.randn(B, V, generator=g) * 3
    logits[:, 10] = logits[:, 20]                # a tie inside the candidate set
    u = torch.rand(B, generator=g)
    sp = L.Sampling()
    sp.do_sample, sp.temperature, sp.top_k, sp.top_p, sp.repetition_penalty = 1, temp, top_k, top_p, 1.0
    sp.suppress_lo, sp.suppress_hi, sp.eos_id = -1, -1, -1
    out, _ = _sample(lib, cuda, logits, sp, uniforms=u)
    osp = O.SamplingParams(do_sample=True, temperature=temp, top_k=top_k, top_p=top_p)
    mism = 0
    for b in range(B):
        s = O.process_logits(logits[b], osp, (), 0)
        ref = O.draw(s, osp, float(u[b]))
        assert torch.isfinite(s[out[b]]), "sampled a filtered id"
        mism += int(ref != int(out[b]))
    assert mism <= 1, f"{mism} draws differ (only CDF-boundary rounding may differ)"
