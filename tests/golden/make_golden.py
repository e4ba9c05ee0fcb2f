"""Generates tests/golden/small_golden.npz from the CPU oracle (seeded, deterministic).

Run from the repo root:  python tests/golden/make_golden.py
The reference holds no golden vectors for this path, so the fixtures are oracle outputs on the
`small` configuration; the oracle itself is pinned against the transformers cousins in
tests/test_oracle_vs_cousins.py.  GPU tests compare the device path with these arrays directly.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "qwen3-tts-apple-silicon_b200"))
from oracle import qwen3_tts_oracle as O  # noqa: E402
from qwen3_tts_b200 import config as Cfg  # noqa: E402
from qwen3_tts_b200.weights import make_weights  # noqa: E402

N_FRAMES = 12


def text_ids(cfg, n, seed):
    g = torch.Generator().manual_seed(seed)
    body = torch.randint(0, cfg.talker.text_vocab_size - 16, (n,), generator=g).tolist()
    return [cfg.im_start_id, cfg.assistant_id, 10] + body + [cfg.im_end_id, 10, cfg.im_start_id, cfg.assistant_id, 10]


def main():
    torch.manual_seed(0)
    cfg = Cfg.small("custom_voice")
    ws = make_weights(cfg, seed=0, head_std=0.2)
    m = O.OracleModel(cfg, ws.fp, kv_dtype=torch.bfloat16)
    ids = text_ids(cfg, 16, 123)
    pre, tr = m.build_prefill(ids, instruct_ids=[3, 1, 4, 1, 5], speaker="vivian", language="chinese")
    codes, rec = m.generate(pre, tr, N_FRAMES, record=True)
    g = torch.Generator().manual_seed(99)
    ccodes = torch.randint(0, cfg.codec.codebook_size, (1, 16, 9), generator=g)
    _, sums = O.rvq_decode(ws.fp, cfg, ccodes, split=True)
    wav = O.codec_forward(ws.fp, cfg, ccodes)[0, 0]
    out = dict(
        text_ids=np.array(ids, dtype=np.int64), prefill=pre.numpy(), trailing=tr.numpy(),
        codes=codes.numpy(), margins=np.array(rec["margins"], dtype=np.float32),
        talker_logits0=rec["talker_logits"][0].numpy(), cp_logits0=rec["cp_logits"][0].numpy(),
        codec_codes=ccodes.numpy(), rvq_sem=sums[0].numpy(), rvq_ac=sums[1].numpy(), wav=wav.numpy())
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "small_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()}, "min margin", float(min(rec["margins"])))


if __name__ == "__main__":
    main()
