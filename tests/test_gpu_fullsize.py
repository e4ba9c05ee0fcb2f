"""Parity at BASELINE.json's FULL sizes (Qwen3-TTS-12Hz-1.7B shapes, random-init seed 0).

The CPU oracle is affordable for a short prompt and two frames (a talker step costs ~0.1 s on the host cores), so the
1.7B path is checked directly against it teacher-forced; everything longer is checked through size-independent properties:
persistent kernel == per-contraction kernels, launch-to-launch bit-reproducibility, batch rows independent of their
neighbours, chunked codec decode == one-shot decode, RVQ gather/sum bit-exact, output lengths of the 30 s clip."""
import pytest
import torch

from oracle import qwen3_tts_oracle as O
from qwen3_tts_b200 import config as Cfg
from qwen3_tts_b200.weights import make_weights

pytestmark = pytest.mark.gpu

LOGIT_RTOL = 1e-2          # BASELINE.json: logits within 1e-2 relative


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


@pytest.fixture(scope="module")
def full(cuda):
    cfg = Cfg.full("custom_voice")
    ws = make_weights(cfg, seed=0, device="cuda", keep_fp=True, keep_q=True)
    return cfg, ws


def _ids(cfg, n, seed):
    g = torch.Generator().manual_seed(seed)
    body = torch.randint(0, cfg.talker.text_vocab_size - 16, (n,), generator=g).tolist()
    return [cfg.im_start_id, cfg.assistant_id, 10] + body + [cfg.im_end_id, 10, cfg.im_start_id, cfg.assistant_id, 10]


def test_full_size_teacher_forced_against_the_cpu_oracle(full):
    """1.7B talker + code predictor, persistent kernel, exact-integer contractions: every logit vector of two frames
    within 1e-2 relative of the oracle's and the same argmax wherever the oracle's own top-2 gap is not a near-tie."""
    from qwen3_tts_b200.engine import TalkerEngine
    cfg, ws = full
    talker_cp = {k: v.cpu() for k, v in ws.fp.items() if not k.startswith("codec.")}
    oracle = O.OracleModel(cfg, talker_cp, kv_dtype=torch.bfloat16)
    with torch.no_grad():
        pre, tr = oracle.build_prefill(_ids(cfg, 6, 1), instruct_ids=[7, 8, 9], speaker="ryan", language="english")
        n = 2
        codes_o, rec = oracle.generate(pre, tr, n, record=True)
    e = TalkerEngine(cfg, ws, "cuda", batch=1, max_frames=8, max_ctx=64, keep_cp_logits=True, prefill="decode")
    e.set_sampling(do_sample=False)
    e.set_forced(codes_o[None])
    e.use_graphs = False
    e.prefill(pre[None], None, tr[None])
    for f in range(n):
        lt = e.logits.clone().cpu()[0]
        assert _rel(lt, rec["talker_logits"][f]) < LOGIT_RTOL, f"talker logits, frame {f}"
        e._frame()
        torch.cuda.synchronize()
        cpl = e.cp_logits.clone().cpu()[:, 0]
        assert _rel(cpl, rec["cp_logits"][f]) < LOGIT_RTOL, f"code-predictor logits, frame {f}"
        own = e.own_codes[0, f].cpu().long()
        for g in range(cfg.cp.num_code_groups):
            lg = rec["talker_logits"][f] if g == 0 else rec["cp_logits"][f][g - 1]
            want = int(rec["own_codes"][f][g])
            if int(own[g]) != want:
                gap = float(lg[want] - lg[int(own[g])])
                assert 0 <= gap <= 3e-3 * float(lg.abs().max()), f"frame {f} group {g}: not a near-tie (gap {gap:.3e})"


def test_full_size_persistent_kernel_equals_per_contraction_path_and_is_reproducible(full):
    from qwen3_tts_b200.engine import TalkerEngine
    cfg, ws = full
    e = TalkerEngine(cfg, ws, "cuda", batch=1, max_frames=8, max_ctx=1024)
    e.use_graphs = False
    torch.manual_seed(5)
    x0 = torch.randn_like(e.x) * 0.02
    outs = {}
    for mega in (False, True):
        e.set_mega(mega); e.use_graphs = False
        runs = []
        for _ in range(3):
            e.pos.fill_(700); e.x.copy_(x0); e._talker_step(True); torch.cuda.synchronize()
            runs.append((e.logits.clone(), e.hidden.clone()))
        assert all(torch.equal(r[0], runs[0][0]) and torch.equal(r[1], runs[0][1]) for r in runs), f"mega={mega}: not reproducible"
        outs[mega] = runs[0]
    assert _rel(outs[True][0].cpu(), outs[False][0].cpu()) < 1e-3
    assert _rel(outs[True][1].cpu(), outs[False][1].cpu()) < 1e-3


def test_full_size_batch_rows_do_not_see_their_neighbours(full):
    """Batch 8 on the tcgen05 path with one prompt repeated in every row: every row must produce the same codes and the
    same logits, bit for bit (rows share GEMM tiles, split-K work spaces and attention launches)."""
    from qwen3_tts_b200.engine import TalkerEngine
    cfg, ws = full
    B, Lp = 8, 40
    e = TalkerEngine(cfg, ws, "cuda", batch=B, max_frames=8, max_ctx=128, attn_nsplit=4)
    e.set_sampling(do_sample=False)
    torch.manual_seed(2)
    emb = (torch.randn(1, Lp, cfg.talker.hidden_size) * 0.02).expand(B, Lp, -1).contiguous()
    e.prefill(emb, None, None)
    codes = e.generate(3)
    assert all(torch.equal(codes[b], codes[0]) for b in range(B))
    assert all(torch.equal(e.logits[b], e.logits[0]) for b in range(B))
    assert int(codes.min()) >= 0 and int(codes[:, :, 1:].max()) < cfg.cp.vocab_size


def test_full_size_codec_30s_clip_properties(full):
    """BASELINE config 2: 375 frames (30 s).  Chunked decode (300 + 25 frames of context, the cousin's chunked_decode)
    against a one-shot decode, sample counts, RVQ gather/sum bit-exact, rows of a batch independent."""
    from qwen3_tts_b200.codec import CodecDecoder
    cfg, ws = full
    k = cfg.codec
    dec = CodecDecoder(cfg, ws, "cuda")
    g = torch.Generator().manual_seed(2)
    codes = torch.randint(0, k.codebook_size, (4, k.num_quantizers, 375), generator=g, dtype=torch.int32).cuda()
    wav = dec.decode(codes)                                    # chunked: 300 frames, then 75 with 25 frames of left context
    n_chunked = k.out_len(300) + k.out_len(100) - 25 * k.hop
    assert wav.shape == (4, n_chunked) and bool(torch.isfinite(wav).all()) and float(wav.abs().max()) <= 1.0
    one = dec.forward(codes)                                   # one vocoder call over all 375 frames
    assert one.shape[-1] == k.out_len(375)                     # (the cousin config gives 718 890; pinned in test_oracle_vs_cousins)
    # causal system: inside the first chunk (away from its right trim) both calls compute the same samples
    n1 = k.out_len(300) - 4 * k.hop
    err = (wav[:, :n1] - one[:, :n1]).double()
    snr = 10 * torch.log10(one[:, :n1].double().pow(2).sum() / err.pow(2).sum().clamp_min(1e-30))
    assert float(snr) > 60.0, f"first chunk vs one-shot: {float(snr):.1f} dB"
    # rows of a batch are independent, bit for bit
    single = dec.decode(codes[2:3])
    assert torch.equal(single[0], wav[2])
    # RVQ gather/sum at the clip's size: bit-exact against the fp32 table sums in the oracle's order
    w = {kk: v.cpu() for kk, v in ws.fp.items() if kk.startswith("codec.rvq.")}
    _, sums = O.rvq_decode(w, cfg, codes[:2].cpu().long(), split=True)
    sem, ac = dec.rvq_sums(codes[:2])
    assert torch.equal(sem.cpu(), sums[0]) and torch.equal(ac.cpu(), sums[1])
