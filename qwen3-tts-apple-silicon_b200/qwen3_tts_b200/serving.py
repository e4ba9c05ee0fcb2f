"""Continuous batching over a paged K/V pool (SURVEY.md 7 step 8, 8e "its own continuous-batching loop with its own paged KV
pool"; BASELINE config 4 "paged KV cache").  The reference runs one utterance at a time (sessions/custom.py:154-176); serving many
is an extension behind the same model: the rows of the batched engine become SLOTS that requests enter and leave at different
frames.

  * `PagePool`: free list over the talker's K/V pages (16 tokens each).  A request takes ceil((prompt + max_frames) / 16) pages on
    admission and gives them back the moment it samples EOS or reaches its frame budget; page 0 is a scratch page idle slots
    point at.  The pool may be smaller than slots x max_ctx: admission waits for pages, not for a whole-batch boundary.
  * admission: the prompt rows of all requests admitted at one boundary run as ONE ragged tcgen05 GEMM prefill (engine.prefill_rows)
    while the running slots wait; the tail (final norm + codec head) of just those rows is scattered into their slots.
  * the frame loop is the captured batched frame graph (csrc/engine.cu `frame`): every slot has its own frame counter
    (q3t_frame_args.step_per_row) and only active slots advance (q3t_frame_args.active).  The host looks at the `done` flags and
    the counters every `sync_every` frames - the only synchronisation - retires finished slots and admits waiting requests.
Everything arithmetic is a kernel of libq3tts_b200.so; this module only moves page ids and row indices.
"""
from __future__ import annotations

import collections
import ctypes as C
from dataclasses import dataclass, field
from typing import Deque, Dict, List, Optional, Sequence

import torch

from . import lib as L
from .engine import TalkerEngine


class PagePool:
    """Free-list allocator of K/V pages 1..n (page 0 is the scratch page of idle slots)."""

    def __init__(self, n_pages: int):
        self.n_pages = n_pages
        self._free: List[int] = list(range(n_pages, 0, -1))
        self.peak_used = 0

    @property
    def free(self) -> int:
        return len(self._free)

    def alloc(self, k: int) -> Optional[List[int]]:
        if k > len(self._free):
            return None
        pages = [self._free.pop() for _ in range(k)]
        self.peak_used = max(self.peak_used, self.n_pages - len(self._free))
        return pages

    def release(self, pages: Sequence[int]) -> None:
        assert all(0 < p <= self.n_pages for p in pages) and not (set(pages) & set(self._free)), "double free / foreign page"
        self._free.extend(reversed(list(pages)))


@dataclass
class Request:
    rid: int
    prefill: torch.Tensor                 # [L, H] prompt embeddings (Model.build_prefill)
    trailing: torch.Tensor                # [n, H] trailing text rows, last = tts_pad
    max_frames: int
    codes: Optional[torch.Tensor] = None  # [T, G] int32 on the host once finished (trimmed at EOS)
    admitted_at: int = -1                 # scheduler frame at admission / retirement (for the tests and the bench)
    finished_at: int = -1
    pages: List[int] = field(default_factory=list)


def plan_admissions(pending: Deque["Request"], slots: Sequence[Optional["Request"]], pool: PagePool, pages_per_seq: int,
                    engine_max_frames: int, sync_every: int) -> List:
    """Admission policy, first come first served: the head of the queue enters the lowest free slot as soon as the pool holds
    ceil((prompt + max_frames + sync_every) / 16) pages for it (a finished slot keeps stepping until the next look at the flags, so
    those frames are budgeted too); a head-of-line request that does not fit makes everybody behind it wait.  Takes the pages and
    pops the queue; returns [(slot, request)].  Pure host logic (tests/test_serving_cpu.py)."""
    pairs = []
    for b in range(len(slots)):
        if slots[b] is not None or not pending:
            continue
        r = pending[0]
        need = (int(r.prefill.shape[0]) + r.max_frames + sync_every + L.KV_PAGE - 1) // L.KV_PAGE
        if need > pages_per_seq or r.max_frames > engine_max_frames:
            raise ValueError(f"request {r.rid}: {need} pages / {r.max_frames} frames exceed the engine's max_ctx / max_frames")
        pages = pool.alloc(need)
        if pages is None:
            break                      # head-of-line request waits for pages
        r.pages = pages
        pending.popleft()
        pairs.append((b, r))
    return pairs


class ContinuousBatcher:
    def __init__(self, engine: TalkerEngine, pool_pages: int, sync_every: int = 8):
        e = self.e = engine
        assert e.B > 2 and not e.fa.use_mega, "continuous batching runs on the batched (tcgen05 GEMM) path: more than two slots"
        assert hasattr(e, "talker_tbl") and int(e.talker_tbl.abs().sum()) == 0, "build the engine with kv_pages=<pool size>"
        self.pool = PagePool(pool_pages)
        self.sync_every = sync_every
        dev = e.dev
        # one frame counter per slot + the active mask, then capture the frame graph with them baked in
        e.step = torch.zeros(e.B, device=dev, dtype=torch.int32)
        self.active = torch.zeros(e.B, device=dev, dtype=torch.int32)
        e.fa.step, e.fa.step_per_row, e.fa.active = e.step.data_ptr(), 1, self.active.data_ptr()
        e._graphs = {}
        e._ensure_cp_proj_rows()
        e._ensure_graphs()                # warm-up + capture on idle slots (pos 0 over the scratch page)
        e.reset()
        self.slots: List[Optional[Request]] = [None] * e.B
        self.frame = 0
        self.stats = dict(admitted=0, retired=0, prefill_calls=0, frames=0, slot_frames_active=0)

    # ---- admission ------------------------------------------------------------------------------------------------
    def _admit(self, pairs: List):
        """pairs = [(slot, request)]: pages, block-table rows, trailing text, ONE ragged prefill, tail scattered into the slots."""
        e = self.e
        dev = e.dev
        H = self.e.cfg.talker.hidden_size
        slots = [b for b, _ in pairs]
        lengths = [int(r.prefill.shape[0]) for _, r in pairs]
        tbl_rows = torch.zeros(len(pairs), e.talker_tbl.shape[1], dtype=torch.int32)
        for i, (b, r) in enumerate(pairs):
            tbl_rows[i, :len(r.pages)] = torch.tensor(r.pages, dtype=torch.int32)
            n_tr = r.trailing.shape[0]
            assert n_tr <= e.max_trailing, "raise max_trailing for streaming text"
            tr = r.trailing.to(dev, torch.float32)
            e.trailing[b, :n_tr] = tr
            e.trailing[b, n_tr:] = tr[-1:]
        sl = torch.tensor(slots, device=dev, dtype=torch.long)
        e.talker_tbl[sl] = tbl_rows.to(dev)
        rows = e.prefill_rows(torch.cat([r.prefill.to(dev, torch.float32) for _, r in pairs], 0), lengths, slots)
        last = torch.tensor([sum(lengths[:i + 1]) - 1 for i in range(len(pairs))], device=dev)
        # final norm + codec head of the admitted rows only: a scratch copy of the argument block with its own small buffers
        n = len(pairs)
        fa2 = L.FrameArgs.from_buffer_copy(e.fa)
        x2 = rows[last].contiguous()
        h2 = torch.empty(n, H, device=dev)
        l2 = torch.empty(n, e.cfg.talker.vocab_size, device=dev)
        fa2.B, fa2.x, fa2.hidden, fa2.logits = n, x2.data_ptr(), h2.data_ptr(), l2.data_ptr()
        L.check(e.lib.q3t_talker_tail(C.byref(fa2), L.stream_ptr()), "talker_tail")
        e.hidden[sl] = h2
        e.logits[sl] = l2
        e.pos[sl] = torch.tensor(lengths, device=dev, dtype=torch.int32)
        e.step[sl] = 0
        e.done[sl] = 0
        e.seen[sl] = 0
        self.active[sl] = 1
        torch.cuda.synchronize()
        for b, r in pairs:
            self.slots[b] = r
            r.admitted_at = self.frame
        self.stats["admitted"] += n
        self.stats["prefill_calls"] += 1

    def _retire(self, b: int, n_frames: int):
        e, r = self.e, self.slots[b]
        codes = e.codes[b, :n_frames].cpu()
        eos = (codes[:, 0] == e.cfg.talker.codec_eos_id).nonzero()
        r.codes = codes[: int(eos[0, 0])] if eos.numel() else codes
        r.finished_at = self.frame
        self.pool.release(r.pages)
        r.pages = []
        e.talker_tbl[b] = 0                # back to the scratch page
        e.pos[b] = 0
        self.active[b] = 0
        self.slots[b] = None
        self.stats["retired"] += 1

    # ---- the loop ---------------------------------------------------------------------------------------------------
    def run(self, requests: Sequence[Request]) -> Dict[int, Request]:
        e = self.e
        pending: Deque[Request] = collections.deque(requests)
        finished: Dict[int, Request] = {}
        per_seq = e.talker_tbl.shape[1]
        while pending or any(s is not None for s in self.slots):
            # admit: first come, first served, while a slot AND enough pages are free
            pairs = plan_admissions(pending, self.slots, self.pool, per_seq, e.max_frames, self.sync_every)
            if pairs:
                self._admit(pairs)
            n_active = sum(s is not None for s in self.slots)
            if n_active == 0:
                if pending:
                    raise RuntimeError("page pool too small for the next request")
                break
            for _ in range(self.sync_every):
                e._run("frame")
            self.frame += self.sync_every
            self.stats["frames"] += self.sync_every
            self.stats["slot_frames_active"] += self.sync_every * n_active
            state = torch.stack([e.done, e.step]).cpu()      # the one synchronisation per `sync_every` frames
            for b, r in enumerate(self.slots):
                if r is None:
                    continue
                n = int(state[1, b])
                if int(state[0, b]) or n >= r.max_frames:
                    self._retire(b, min(n, r.max_frames))
                    finished[r.rid] = r
        return finished
