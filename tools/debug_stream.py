import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "qwen3-tts-apple-silicon_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import qwen3_tts_oracle as O
from qwen3_tts_b200 import config as Cfg
from qwen3_tts_b200.model import Model
from qwen3_tts_b200.weights import make_weights
from test_gpu_pipeline import _text_ids
cfg = Cfg.small("custom_voice"); ws = make_weights(cfg, seed=0, head_std=0.2)
model = Model(cfg, ws, "cuda", max_frames=64, max_ctx=256, max_trailing=64)
oracle = O.OracleModel(cfg, ws.fp, kv_dtype=torch.bfloat16)
ids = _text_ids(cfg, 9, 4)
vec = torch.randn(cfg.talker.hidden_size, generator=torch.Generator().manual_seed(77)) * 0.02
pre, tr = oracle.build_prefill(ids, streaming=True, speaker_vec=vec)
n = 14
codes_o, rec = oracle.generate(pre, tr, n, record=True)
print("talker margins", [round(m, 4) for m in rec["margins"]])
cpm = [float((torch.topk(c, 2, -1).values[:, 0] - torch.topk(c, 2, -1).values[:, 1]).min()) for c in rec["cp_logits"]]
print("cp min margins", [round(m, 4) for m in cpm])
e = model.engine
e.set_sampling(do_sample=False)
for mega in (True, False):
    e.set_mega(mega)
    c = model.generate_codes(pre.cuda(), tr.cuda(), n).cpu().long()
    diff = (c != codes_o).any(1)
    print("mega", mega, "first differing frame", int(diff.nonzero()[0]) if diff.any() else None)
    if diff.any():
        f = int(diff.nonzero()[0]); print(" dev", c[f].tolist()); print(" ora", codes_o[f].tolist())
# teacher forced
for mega in (True, False):
    e2 = type(e)(cfg, ws, "cuda", batch=1, max_frames=64, max_ctx=256, keep_cp_logits=True, max_trailing=64, use_mega=mega)
    e2.set_sampling(do_sample=False); e2.set_forced(codes_o[None]); e2.use_graphs = False
    e2.prefill(pre[None], None, tr[None])
    worst = 0
    for f in range(n):
        tl = e2.logits[0].clone().cpu()
        e2._run("frame")
        cl = e2.cp_logits[:, 0].clone().cpu()
        r1 = float((tl - rec["talker_logits"][f]).abs().max() / rec["talker_logits"][f].abs().max())
        r2 = float((cl - rec["cp_logits"][f]).abs().max() / rec["cp_logits"][f].abs().max())
        print(f"mega={mega} frame {f}: talker rel {r1:.2e} cp rel {r2:.2e}")
