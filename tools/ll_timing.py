"""Per-phase timeline of the persistent data-flow kernel (csrc/frame_ll.cu): id-coded clock64 stamps written by
thread 0 of every CTA ({id:20 | cycles:44}; Q3T_SM_MHZ converts to time, default 1965).  Q3T_LL_FINE=1 adds sub-phase marks.  Also cross-checks the talker logits
against the one-kernel-per-contraction path and times talker step / frame with CUDA events.

usage: python tools/ll_timing.py [full|small] [ctx]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# the stamps are only compiled into the profiling build (csrc/build.sh); event timings of the production library: Q3T_LL_PROD=1
if not os.environ.get("Q3T_LIB") and not os.environ.get("Q3T_LL_PROD"):
    os.environ["Q3T_LIB"] = os.path.join(ROOT, "qwen3-tts-apple-silicon_b200", "qwen3_tts_b200", "libq3tts_b200_prof.so")
sys.path.insert(0, os.path.join(ROOT, "qwen3-tts-apple-silicon_b200"))
from qwen3_tts_b200 import config as Cfg
from qwen3_tts_b200.engine import TalkerEngine
from qwen3_tts_b200.weights import make_weights

NAMES = {0: "start", 1: "qkv.pro", 2: "qkv.gemv", 3: "attn", 4: "o.pro", 5: "o.gemv", 6: "gu.pro", 7: "gu.gemv", 8: "down.pro",
         9: "down.gemv", 10: "end", 11: "sample", 12: "cp_pass", 32: "f.poll", 34: "f.tiles", 35: "f.gbar", 36: "f.att.A.preload",
         37: "f.att.B.qwords", 38: "f.att.C.normrope", 39: "f.att.D.bar", 40: "f.att.E.scores", 41: "f.att.F.halfmerge",
         42: "f.att.G.bar", 43: "f.merge", 44: "f.dig"}
MHZ = float(os.environ.get("Q3T_SM_MHZ", "1965"))


def decode(timing, G, NST):
    t = timing.view(G, NST).cpu()
    ids = (t >> 44).numpy()
    cyc = (t & ((1 << 44) - 1)).double()
    ns = (cyc * (1e3 / MHZ)).numpy()          # clock64 cycles of the CTA's own SM -> ns at the SM clock under load
    return ids, ns


def report(ids, ns, ctas, title, skip_first=9, quiet=None):
    """Duration of every interval, attributed to (coarse phase it ends in, id of the closing stamp)."""
    if quiet is None:
        print(f"--- {title}")
    import collections
    for cta in ctas:
        n = int((ns[cta] > 0).sum())
        acc = collections.OrderedDict()
        cur_phase_start = 0
        # walk the stamps; key = (closing coarse id that follows, this stamp id)
        seq = [(int(ids[cta, i]), int(ns[cta, i])) for i in range(n)]
        # find for every stamp the next coarse stamp id (its phase)
        phase_of = [None] * n
        nxt = None
        for i in range(n - 1, -1, -1):
            if seq[i][0] < 32:
                nxt = seq[i][0]
            phase_of[i] = nxt
        coarse_seen = 0
        for i in range(1, n):
            if seq[i][0] < 32:
                coarse_seen += 1
            if coarse_seen <= skip_first:
                continue
            key = (phase_of[i], seq[i][0])
            d = seq[i][1] - seq[i - 1][1]
            a = acc.setdefault(key, [0, 0])
            a[0] += d
            a[1] += 1
        n_layers = max(1, max((v[1] for k, v in acc.items() if k[1] == 2), default=1))   # QKV phase-end stamps = layers
        parts = []
        tot = 0.0
        for (ph, sid), (d, c) in acc.items():
            if ph in (10, 11, 12) and sid < 32 and ph != sid:
                continue
            per_layer = d / 1e3 / n_layers
            tot += per_layer
            nm = NAMES.get(ph, str(ph)) + ("" if sid == ph else ":" + NAMES.get(sid, str(sid)))
            parts.append(f"{nm}={per_layer:.2f}")
        if quiet is None:
            print(f"cta {cta:3d} ({n} stamps, /{n_layers} layers): " + "  ".join(parts) + f"  | sum {tot:.2f} us")
        else:
            quiet[cta] = dict(p.split("=") for p in parts)


def main():
    size = sys.argv[1] if len(sys.argv) > 1 else "full"
    cfg = getattr(Cfg, size)()
    ctx = int(sys.argv[2]) if len(sys.argv) > 2 else (300 if size == "full" else 40)
    ws = make_weights(cfg, seed=0, device="cuda", keep_fp=False, parts=("talker", "cp"))
    e = TalkerEngine(cfg, ws, "cuda", batch=1, max_frames=64, max_ctx=int(os.environ.get("Q3T_MAX_CTX", "1024")))
    del ws
    e._ensure_cp_proj_rows()                # projected embedding tables of the code-predictor passes (Q3T_CP_PROJ_TABLES=0: off)
    NST = 2048
    G = torch.cuda.get_device_properties(0).multi_processor_count
    timing = torch.zeros(G * NST, dtype=torch.int64, device="cuda")
    e.use_graphs = False
    torch.manual_seed(1234)                 # the path-to-path difference depends on the input: keep runs comparable
    e.x.normal_(0, 0.02)
    x0 = e.x.clone()
    # ---- correctness: persistent vs multi-kernel talker step
    e.set_mega(False); e.use_graphs = False
    e.pos.fill_(ctx); e.x.copy_(x0); e._talker_step(True); torch.cuda.synchronize()
    ref = e.logits.clone(); refh = e.hidden.clone()
    e.set_mega(True); e.use_graphs = False
    e.pos.fill_(ctx); e.x.copy_(x0); e._talker_step(True); torch.cuda.synchronize()
    print("state", e.ll_state.tolist())
    rel = float((e.logits - ref).abs().max() / ref.abs().max())
    relh = float((e.hidden - refh).abs().max() / refh.abs().max())
    print(f"talker step persistent vs multi-kernel: logits rel {rel:.2e} hidden rel {relh:.2e}")
    # both paths must be bit-reproducible run to run (no atomics, fixed summation order), the very first launch included
    first_mega = e.logits.clone()
    for mega in (False, True):
        e.set_mega(mega); e.use_graphs = False
        outs = [ref if not mega else first_mega]
        for _ in range(3):
            e.pos.fill_(ctx); e.x.copy_(x0); e._talker_step(True); torch.cuda.synchronize()
            outs.append(e.logits.clone())
        print(f"  mega={mega}: max |diff| of launches 2..4 vs launch 1: " + " ".join(f"{float((o - outs[0]).abs().max()):.3e}" for o in outs[1:]) +
              f"   rel vs ref: " + " ".join(f"{float((o - ref).abs().max() / ref.abs().max()):.2e}" for o in outs))
    # does a fresh exchange workspace (as on the first launch) change the result?
    e.set_mega(True); e.use_graphs = False
    e.ll_work.zero_(); e.ll_state.zero_()
    e.pos.fill_(ctx); e.x.copy_(x0); e._talker_step(True); torch.cuda.synchronize()
    print(f"  mega after zeroing ll_work/ll_state: rel vs ref {float((e.logits - ref).abs().max() / ref.abs().max()):.2e}  vs launch 2 {float((e.logits - outs[1]).abs().max()):.3e}")
    e.set_mega(True); e.use_graphs = False
    # ---- timeline of the talker step
    e.fa.ll_timing = timing.data_ptr()
    for it in range(3):
        timing.zero_(); e.pos.fill_(ctx); e.x.copy_(x0)
        e._talker_step(True); torch.cuda.synchronize()
    ids, ns = decode(timing, G, NST)
    n0 = int((ns[0] > 0).sum())
    print(f"ctx={ctx}: talker step, CTA0 first->last stamp {(ns[0, n0 - 1] - ns[0, 0]) / 1e3:.1f} us")
    report(ids, ns, (0, 1, 7, 39, 40, 73, G - 1), f"talker step ctx={ctx} (steady-state per layer)")
    if os.environ.get("Q3T_LL_ALL"):
        # one compact line per CTA: who waits (large poll) and who is waited for (small poll) in every exchange
        rows = {}
        report(ids, ns, range(G), "", quiet=rows)
        keys = ["qkv.pro:f.poll", "qkv.pro", "qkv.gemv:f.tiles", "qkv.gemv", "attn", "o.pro:f.poll", "o.pro", "o.gemv:f.tiles", "o.gemv",
                "gu.pro:f.poll", "gu.pro", "gu.gemv:f.tiles", "gu.gemv:f.gbar", "gu.gemv", "down.pro:f.poll", "down.pro", "down.gemv:f.tiles", "down.gemv"]
        print("cta " + " ".join(k.replace(":f.", ":")[-9:].rjust(9) for k in keys))
        for cta in range(G):
            print(f"{cta:3d} " + " ".join(rows[cta].get(k, "-").rjust(9) for k in keys))
    # CUDA-event time of back-to-back steps
    s, f = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e.fa.ll_timing = 0
    t_ = cfg.talker
    params = t_.num_layers * (t_.hidden_size * (t_.q_dim + 2 * t_.kv_dim) + t_.q_dim * t_.hidden_size + 3 * t_.hidden_size * t_.intermediate_size) + t_.vocab_size * t_.hidden_size
    for c in (ctx, 100, 600, 900):
        if c >= e.max_ctx:
            continue
        e.pos.fill_(c)
        reps = 50
        s.record()
        for _ in range(reps):
            e._talker_step(True)
        f.record(); torch.cuda.synchronize()
        us = s.elapsed_time(f) * 1e3 / reps
        print(f"talker step ctx={c} (eager, events): {us:.1f} us -> {params * 1.0625 / us / 1e3:.0f} GB/s weights-only")
    # ---- whole frame
    e.reset(); e.pos.fill_(ctx); e.x.copy_(x0); e._talker_step(True)
    e.fa.ll_timing = timing.data_ptr()
    for it in range(2):
        timing.zero_()
        e._frame(); torch.cuda.synchronize()
    ids, ns = decode(timing, G, NST)
    n_used = int((ns[0] > 0).sum())
    print(f"frame: {n_used} stamps, CTA0 first->last {(ns[0, n_used - 1] - ns[0, 0]) / 1e3:.1f} us; state {e.ll_state.tolist()}; codes {e.cur_codes.tolist()}")
    # code-predictor passes: time between consecutive cp_pass stamps on CTA 0
    cp = [int(ns[0, i]) for i in range(n_used) if ids[0, i] == 12]
    if len(cp) > 2:
        d = [(cp[i + 1] - cp[i]) / 1e3 for i in range(len(cp) - 1)]
        print("cp pass us:", " ".join(f"{v:.1f}" for v in d))
    if os.environ.get("Q3T_LL_FINE"):
        # code-predictor passes only: the stamps between the first and the last cp_pass mark of a CTA, averaged per pass
        import numpy as np
        ids2, ns2 = np.zeros_like(ids), np.zeros_like(ns)
        for c in (0, 1, 73, G - 1):
            n_c = int((ns[c] > 0).sum())
            marks = [i for i in range(n_c) if ids[c, i] == 12]
            if len(marks) >= 2:
                a, b = marks[0], marks[-1] + 1
                ids2[c, :b - a] = ids[c, a:b]; ns2[c, :b - a] = ns[c, a:b]
                rows = {}
                report(ids2, ns2, (c,), "", skip_first=0, quiet=rows)
                npass = len(marks) - 1
                print(f"cp-only cta {c:3d} ({npass} passes, us per layer): " + "  ".join(f"{k}={float(v):.2f}" for k, v in rows[c].items()))
    e.fa.ll_timing = 0
    s.record()
    for _ in range(20):
        e._frame()
    f.record(); torch.cuda.synchronize()
    print(f"frame (eager, events): {s.elapsed_time(f) * 1e3 / 20:.1f} us -> RTFx {80e3 / (s.elapsed_time(f) * 1e3 / 20):.1f}")


if __name__ == "__main__":
    main()
