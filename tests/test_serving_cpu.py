"""Host logic of the continuous batcher that needs no GPU: the K/V page free list (qwen3_tts_b200/serving.py)."""
import random

import pytest

from qwen3_tts_b200.serving import PagePool


def test_page_pool_never_hands_out_page_zero_or_a_page_twice():
    p = PagePool(37)
    rnd = random.Random(3)
    held = []
    for _ in range(400):
        if held and (rnd.random() < 0.45 or p.free == 0):
            p.release(held.pop(rnd.randrange(len(held))))
        else:
            got = p.alloc(rnd.randint(1, 9))
            if got is not None:
                held.append(got)
        flat = [x for h in held for x in h]
        assert len(flat) == len(set(flat)) and all(1 <= x <= 37 for x in flat)
        assert p.free == 37 - len(flat) and p.peak_used <= 37
    for h in held:
        p.release(h)
    assert p.free == 37


def test_page_pool_refuses_what_it_cannot_serve_and_detects_double_free():
    p = PagePool(8)
    a = p.alloc(5)
    assert p.alloc(4) is None and p.free == 3          # no partial allocation
    b = p.alloc(3)
    assert p.free == 0 and p.alloc(1) is None
    p.release(a)
    with pytest.raises(AssertionError):
        p.release(a)                                   # double free
    with pytest.raises(AssertionError):
        p.release([0])                                 # the scratch page is never in the pool
    with pytest.raises(AssertionError):
        p.release([9])                                 # foreign page
    p.release(b)
    assert p.free == 8 and p.peak_used == 8


def _req(rid, rows, frames):
    import torch
    from qwen3_tts_b200.serving import Request
    return Request(rid, torch.zeros(rows, 1), torch.zeros(1, 1), frames)


def test_admission_policy_is_first_come_first_served_over_slots_and_pages():
    """A whole serving session on the host: the admission policy (serving.plan_admissions) + retirement, with frame counts standing in
    for the GPU.  Requests enter in arrival order, the head of the line blocks the queue when the pool cannot hold it, every page comes
    back, nobody starves."""
    import collections
    from qwen3_tts_b200.serving import plan_admissions
    rnd = random.Random(11)
    reqs = [_req(i, rnd.randint(20, 90), rnd.randint(8, 120)) for i in range(40)]
    pool, slots, sync = PagePool(30), [None] * 6, 8
    pending = collections.deque(reqs)
    left, order, t, peak_busy = {}, [], 0, 0
    while pending or any(s is not None for s in slots):
        for b, r in plan_admissions(pending, slots, pool, pages_per_seq=16, engine_max_frames=128, sync_every=sync):
            assert slots[b] is None and len(r.pages) == -(-(r.prefill.shape[0] + r.max_frames + sync) // 16)
            slots[b] = r
            left[r.rid] = r.max_frames
            order.append(r.rid)
        busy = [b for b, s in enumerate(slots) if s is not None]
        assert busy, "deadlock: requests pending, no slot busy"
        peak_busy = max(peak_busy, len(busy))
        held = [x for s in slots if s is not None for x in s.pages]
        assert len(held) == len(set(held)) == 30 - pool.free
        t += sync
        for b in busy:
            r = slots[b]
            left[r.rid] -= sync
            if left[r.rid] <= 0:
                pool.release(r.pages)
                r.pages = []
                slots[b] = None
    assert order == list(range(40)), "admission must follow arrival order"
    assert pool.free == 30 and pool.peak_used <= 30 and peak_busy >= 3


def test_admission_refuses_a_request_no_engine_could_hold():
    import collections
    from qwen3_tts_b200.serving import plan_admissions
    with pytest.raises(ValueError):
        plan_admissions(collections.deque([_req(0, 300, 100)]), [None], PagePool(100), pages_per_seq=16, engine_max_frames=128, sync_every=8)
    with pytest.raises(ValueError):
        plan_admissions(collections.deque([_req(0, 10, 500)]), [None], PagePool(100), pages_per_seq=64, engine_max_frames=128, sync_every=8)
    # a pool that is merely too small right now is not an error: the request waits
    pool = PagePool(4)
    q = collections.deque([_req(0, 60, 60)])
    assert plan_admissions(q, [None], pool, pages_per_seq=16, engine_max_frames=128, sync_every=8) == [] and len(q) == 1 and pool.free == 4
