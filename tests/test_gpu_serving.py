"""Continuous batching over a paged K/V pool (qwen3_tts_b200/serving.py; SURVEY 7 step 8, BASELINE config 4 "paged KV cache"):
requests enter and leave the slots of the batched engine at different frames, pages come from a free list and go back at EOS /
end of budget.  Every request must produce exactly the codes it produces alone."""
import pytest
import torch

from oracle import qwen3_tts_oracle as O
from qwen3_tts_b200 import config as Cfg
from qwen3_tts_b200.weights import make_weights

pytestmark = pytest.mark.gpu


def _ids(cfg, n, seed):
    g = torch.Generator().manual_seed(seed)
    body = torch.randint(0, cfg.talker.text_vocab_size - 16, (n,), generator=g).tolist()
    return [cfg.im_start_id, cfg.assistant_id, 10] + body + [cfg.im_end_id, 10, cfg.im_start_id, cfg.assistant_id, 10]


def test_page_pool_free_list():
    from qwen3_tts_b200.serving import PagePool
    p = PagePool(10)
    a, b = p.alloc(4), p.alloc(5)
    assert len(set(a) | set(b)) == 9 and 0 not in a + b and p.free == 1 and p.alloc(2) is None
    p.release(a)
    c = p.alloc(5)
    assert set(c) & set(b) == set() and p.free == 0 and p.peak_used == 10
    with pytest.raises(AssertionError):
        p.release([0])


def test_continuous_batching_matches_every_request_alone(cuda):
    from qwen3_tts_b200.engine import TalkerEngine
    from qwen3_tts_b200.serving import ContinuousBatcher, Request
    cfg = Cfg.small("voice_design")
    ws = make_weights(cfg, seed=5, head_std=0.2)
    oracle = O.OracleModel(cfg, ws.fp, kv_dtype=torch.bfloat16)
    specs = [(6, 5), (11, 21), (9, 9), (14, 30), (7, 13), (20, 6), (5, 17)]          # (text tokens, frame budget)
    prompts = [oracle.build_prefill(_ids(cfg, n, 40 + i), instruct_ids=[1 + i, 2, 3], streaming=(i % 2 == 1)) for i, (n, _) in enumerate(specs)]
    B = 4
    # a pool that cannot hold four of the larger requests at once: admission has to wait for pages as well as for slots
    eng = TalkerEngine(cfg, ws, "cuda", batch=B, max_frames=40, max_ctx=128, max_trailing=32, kv_pages=14)
    eng.set_sampling(do_sample=False)
    cb = ContinuousBatcher(eng, pool_pages=14, sync_every=4)
    reqs = [Request(i, p, t, specs[i][1]) for i, (p, t) in enumerate(prompts)]
    done = cb.run(reqs)
    assert sorted(done) == list(range(len(specs)))
    assert cb.pool.free == 14 and cb.pool.peak_used <= 14 and all(s is None for s in cb.slots)
    assert int(cb.active.sum()) == 0 and int(eng.talker_tbl.abs().sum()) == 0
    assert max(r.admitted_at for r in reqs) > 0, "nothing was admitted mid-flight"
    assert cb.stats["admitted"] == cb.stats["retired"] == len(specs) and cb.stats["prefill_calls"] >= 3
    # reference 1: every request alone through a lock-step engine of the same batch width (identity block table).  The decode frames
    # run the same kernels on the same batch width; the prompt GEMMs see a different row count (split-K choice), so a difference is
    # only accepted where the CPU oracle's own top-2 logits are a near-tie.  Reference 2: the oracle itself, same rule.
    ref = TalkerEngine(cfg, ws, "cuda", batch=B, max_frames=40, max_ctx=128, max_trailing=32)
    ref.set_sampling(do_sample=False)
    n_exact = 0
    for i, (p, t) in enumerate(prompts):
        T = specs[i][1]
        ref.prefill(p[None].expand(B, -1, -1).contiguous(), None, t[None].expand(B, -1, -1).contiguous())
        want = ref.generate(T, check_every=0)[0].cpu()
        got = done[i].codes
        assert got.shape == (T, cfg.cp.num_code_groups), (i, got.shape)
        n_exact += int(torch.equal(got, want))
        co, rec = oracle.generate(p, t, T, record=True)
        for name, cand in (("batcher", got.long()), ("lock-step", want.long())):
            diff = cand != co
            if diff.any():
                f = int(diff.any(1).nonzero()[0]); g = int(diff[f].nonzero()[0])
                lg = rec["talker_logits"][f] if g == 0 else rec["cp_logits"][f][g - 1]
                gap = float(lg[int(co[f, g])] - lg[int(cand[f, g])])
                assert 0 <= gap <= 1e-2 * float(lg.abs().max()), f"{name}, request {i} frame {f} group {g}: gap {gap:.3e}"
    assert n_exact >= len(specs) - 1, f"only {n_exact} of {len(specs)} requests equal their stand-alone run bit for bit"


def test_eos_returns_pages_early(cuda):
    """A slot that samples EOS is retired at the next look at the flags: its pages return to the pool before its frame budget ends.
    (Random-init models practically never emit EOS, so the codec head's EOS row is made dominant for one prompt.)"""
    from qwen3_tts_b200.engine import TalkerEngine
    from qwen3_tts_b200.serving import ContinuousBatcher, Request
    cfg = Cfg.small("voice_design")
    ws = make_weights(cfg, seed=6, head_std=0.2)
    oracle = O.OracleModel(cfg, ws.fp, kv_dtype=torch.bfloat16)
    prompts = [oracle.build_prefill(_ids(cfg, 8, 70 + i), instruct_ids=[2, 3]) for i in range(3)]
    eng = TalkerEngine(cfg, ws, "cuda", batch=3, max_frames=40, max_ctx=128, max_trailing=4, kv_pages=24)
    eng.set_sampling(do_sample=False)
    cb = ContinuousBatcher(eng, pool_pages=24, sync_every=2)
    reqs = [Request(i, p, t, 30) for i, (p, t) in enumerate(prompts)]
    # find out what request 0 generates, then force EOS as its 4th code-0 through teacher forcing of that slot only
    forced = torch.zeros(3, 40, cfg.cp.num_code_groups, dtype=torch.int32)
    out = cb.run([Request(9, prompts[0][0], prompts[0][1], 6)])
    base = out[9].codes
    assert base.shape[0] == 6
    f0 = base.clone().int()
    f0[3, 0] = cfg.talker.codec_eos_id
    # teacher forcing is per slot row: slot 0 follows f0 (EOS at frame 3), the others follow their own choices
    own = {}
    for i in (1, 2):
        own[i] = cb.run([Request(20 + i, prompts[i][0], prompts[i][1], 30)])[20 + i].codes
    forced[0, :6] = f0
    forced[1, :30] = own[1].int()
    forced[2, :30] = own[2].int()
    eng.set_forced(forced)
    eng._ensure_graphs()
    done = cb.run(reqs)
    assert done[0].codes.shape[0] == 3 and done[0].finished_at < done[1].finished_at
    assert done[1].codes.shape[0] == 30 and torch.equal(done[1].codes, own[1]) and torch.equal(done[2].codes, own[2])
    assert cb.pool.free == 24
