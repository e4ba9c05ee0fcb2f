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
