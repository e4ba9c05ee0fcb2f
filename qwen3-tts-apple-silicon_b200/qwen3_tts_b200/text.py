"""Long-text segmentation (SURVEY 8f-4 / BASELINE config 5: "2-min texts chunked"): a text longer than one utterance
budget is cut at sentence boundaries into segments that are generated one after the other (the shim joins them into the
single audio_000.wav the reference reads, io.py:156).  Host-side string work only - no arithmetic of the hot path."""
from __future__ import annotations

import re
from typing import List

# sentence enders of the languages the reference lists (config.py:44-49 speakers: English, Chinese, Japanese, Korean ...),
# kept with the sentence; a newline always ends a sentence
_SENT = re.compile(r"[^.!?。！？；;\n]*(?:[.!?。！？；;]+[\"'”’)\]]*|\n|$)", re.S)
_SOFT = re.compile(r"(?<=[,，、:：])\s*|\s+")


def split_sentences(text: str) -> List[str]:
    out = [m.group(0).strip() for m in _SENT.finditer(text)]
    return [s for s in out if s]


def _hard_split(sentence: str, max_chars: int) -> List[str]:
    """A single sentence longer than the budget: cut at commas / spaces, at max_chars as the last resort."""
    parts, cur = [], ""
    for piece in [p for p in _SOFT.split(sentence) if p]:
        while len(piece) > max_chars:                     # no soft boundary at all (e.g. unspaced CJK run)
            if cur:
                parts.append(cur); cur = ""
            parts.append(piece[:max_chars]); piece = piece[max_chars:]
        cand = (cur + " " + piece) if cur else piece
        if len(cand) <= max_chars:
            cur = cand
        else:
            parts.append(cur); cur = piece
    if cur:
        parts.append(cur)
    return parts


def segment_text(text: str, max_chars: int = 400) -> List[str]:
    """Greedy packing of whole sentences into segments of at most max_chars characters (order preserved, nothing dropped
    but surrounding whitespace).  max_chars <= 0 disables segmentation."""
    text = text.strip()
    if max_chars <= 0 or len(text) <= max_chars:
        return [text] if text else []
    segs, cur = [], ""
    for sent in split_sentences(text):
        for piece in ([sent] if len(sent) <= max_chars else _hard_split(sent, max_chars)):
            cand = (cur + " " + piece) if cur else piece
            if len(cand) <= max_chars:
                cur = cand
            else:
                segs.append(cur); cur = piece
    if cur:
        segs.append(cur)
    return segs
