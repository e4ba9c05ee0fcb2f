#!/usr/bin/env python
"""bench.py -- headline metric of BASELINE.json on B200: RTFx (audio s / wall s) of the Qwen3-TTS-12Hz-1.7B
generation hot path at bs=1, with the talker-decode HBM roofline and the CPU oracle timed beside it.

A "step" = one utterance of BASELINE config 1 (CustomVoice, random-init, greedy, 64 text tokens + 8-token
instruct + speaker/language ids): talker prefill -> 240 frames of [talker decode, sample, 15 code-predictor
passes, next-input sum] -> codec decode of the 240 frames to a 24 kHz waveform (19.2 s of audio).

  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun, one replica per GPU)
  python bench.py --impl reference ...                       (the CPU oracle on the host cores; rank 0 only)

`value`  : device-resident inputs (prefill embeddings already in HBM), CUDA-event time, max over ranks.
`e2e`    : same metric through the reference-facing call - `generate_audio(model=, text=, voice=, instruct=, output_path=)`
           (mlx_audio shim; reference sessions/custom.py:163-170): HOST text -> tokenizer -> H2D ids -> prefill -> frames ->
           codec -> D2H waveform -> audio_000.wav written.
`roofline`: talker decode step (one persistent data-flow launch, csrc/frame_ll.cu): algorithmic bytes per step
             / CUDA-event step time, against MEASURED_PEAKS.json hbm_gbs; `roofline_frame` the same for the whole
            frame launch (talker step + 16 code-predictor passes, code-predictor weights counted once: L2-resident).
Other BASELINE configs as extra keys of the same line: `cfg2` codec alone (30 s clip, B = 1 and 32), `cfg3` Base voice cloning
(38 reference frames + 200 text tokens, streaming, 750 frames), `bs64` = config 4 (64 utterances x 300-token prompts, 960
frames, paged KV), `cfg5` data-parallel long-form (utterances sharded over the ranks in batches of 64; bounded sample).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "qwen3-tts-apple-silicon_b200"))

import torch  # noqa: E402

FRAMES = 240
TEXT_TOKENS = 64
INSTRUCT_TOKENS = 8
METRIC = "rtfx_bs1_talker+code_predictor+codec"
UNIT = "audio_s/wall_s"
# 64 / 8 characters: the byte-level stand-in tokenizer of the random-init set-up maps one character to one token
TEXT = "The quick brown fox jumps over the lazy dog near the river bank."
INSTRUCT = "Cheerful"


def synth_ids(cfg):
    g = torch.Generator().manual_seed(1)
    body = torch.randint(0, 151643 if cfg.talker.text_vocab_size > 151643 else cfg.talker.text_vocab_size - 16,
                         (TEXT_TOKENS,), generator=g).tolist()
    ids = [cfg.im_start_id, cfg.assistant_id, 198] + body + [cfg.im_end_id, 198, cfg.im_start_id, cfg.assistant_id, 198]
    ins = torch.randint(0, 1000, (INSTRUCT_TOKENS,), generator=g).tolist()
    return ids, ins


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.samples, self.stop, self.th = index, [], False, None

    def _loop(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.samples.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def __enter__(self):
        self.th = threading.Thread(target=self._loop, daemon=True)
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.th.join(timeout=6)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if len(s) >= 6 and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) >= 6 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples if len(s) >= 6 for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def talker_step_bytes(cfg, ctx):
    t = cfg.talker
    params = t.num_layers * (t.hidden_size * (t.q_dim + 2 * t.kv_dim) + t.q_dim * t.hidden_size +
                             3 * t.hidden_size * t.intermediate_size) + t.vocab_size * t.hidden_size
    w = params * 1.0625                                                   # uint8 + bf16 scale + bf16 bias per 64
    kv = (ctx + 1) * t.num_layers * 2 * t.kv_dim * 2                      # bf16 K/V read + the new row
    return w + kv, w, kv


def cp_frame_bytes_l2_resident(cfg):
    """Code-predictor weights touched by one frame, counted ONCE (they stay L2-resident between the 16 passes)."""
    c, t = cfg.cp, cfg.talker
    layers = c.num_layers * (c.hidden_size * (c.q_dim + 2 * c.kv_dim) + c.q_dim * c.hidden_size + 3 * c.hidden_size * c.intermediate_size)
    return (layers + c.hidden_size * t.hidden_size + (c.num_code_groups - 1) * c.vocab_size * c.hidden_size) * 1.0625


# ------------------------------------------------------------------------------------------------------------------
# the reference arm / CPU baseline: the oracle (kind "port": neither mlx nor the QwenLM package installs offline)
# ------------------------------------------------------------------------------------------------------------------
class CpuOracle:
    def __init__(self, cfg_name, threads=None):
        from oracle import qwen3_tts_oracle as O
        from qwen3_tts_b200 import config as Cfg
        from qwen3_tts_b200.weights import make_weights
        if threads:
            torch.set_num_threads(threads)
        self.O, self.cfg = O, getattr(Cfg, cfg_name)("custom_voice")
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        ws = make_weights(self.cfg, seed=0, device=dev, keep_fp=True, keep_q=False)       # generate fast, then move to host
        self.w = {k: v.cpu() for k, v in ws.fp.items()}
        del ws
        self.m = O.OracleModel(self.cfg, self.w, kv_dtype=torch.bfloat16)
        self.ids, self.ins = synth_ids(self.cfg)

    def sample(self, frames_sample):
        """prefill + `frames_sample` REAL frames (context grows from the prompt on) + codec of those frames, projected to the
        240-frame utterance (prefill counted once; the per-frame cost of the oracle grows < 2 % over 240 frames: attention over
        <= 330 tokens is noise next to 1.5 G weight MACs per step)."""
        O, m = self.O, self.m
        with torch.no_grad():
            t0 = time.perf_counter()
            pre, tr = m.build_prefill(self.ids, instruct_ids=self.ins, speaker="ryan", language="english")
            m.talker.reset()
            m.talker_forward(pre)
            t_prefill = time.perf_counter() - t0
            t0 = time.perf_counter()
            codes = m.generate(pre, tr, frames_sample)          # re-runs the prefill; subtract it below
            t_gen = time.perf_counter() - t0 - t_prefill
            t0 = time.perf_counter()
            O.codec_forward(self.w, self.cfg, codes.t()[None])
            t_codec = time.perf_counter() - t0
        per_frame = (t_gen + t_codec) / frames_sample
        total = t_prefill + FRAMES * per_frame
        return (FRAMES * 0.08) / total, dict(prefill_s=t_prefill, gen_s_per_frame=t_gen / frames_sample,
                                             codec_s_per_frame=t_codec / frames_sample, frames_measured=frames_sample)


def cpu_sample_text(n):
    if n >= FRAMES:
        return (f"PyTorch-CPU fp32 oracle (oracle/qwen3_tts_oracle.py): the WHOLE utterance - prompt prefill, all {FRAMES} frames (talker step, sampler, "
                "15 code-predictor passes, next-input sum), codec of all frames; nothing projected")
    return (f"PyTorch-CPU fp32 oracle (oracle/qwen3_tts_oracle.py): real prompt prefill + {n} real frames (talker step, sampler, 15 "
            f"code-predictor passes, next-input sum; context growing from the prompt) + codec of those frames, projected to {FRAMES} frames")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    oracle = CpuOracle(args.size)
    # every step runs the whole utterance (--ref-frames 240) when W + K such steps fit ~4 minutes of host time; otherwise the largest
    # number of real frames per step that does (never fewer than 24), projected to the utterance - decided from a 24-frame probe
    t0 = time.perf_counter()
    oracle.sample(min(24, args.ref_frames))
    per_frame = (time.perf_counter() - t0) / min(24, args.ref_frames)
    n_steps = args.warmup + args.steps
    frames = max(min(24, args.ref_frames), min(args.ref_frames, int(240.0 / (n_steps * per_frame))))
    vals, detail = [], {}
    for i in range(n_steps):
        v, detail = oracle.sample(frames)
        if i >= args.warmup:
            vals.append(v)
    v = statistics.mean(vals)
    args.ref_frames = frames
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": FRAMES * 80.0 / v, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                             "sample": cpu_sample_text(args.ref_frames), **detail},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(args):
    return {"workload": "BASELINE config 1: Qwen3-TTS-12Hz-1.7B-CustomVoice random-init, greedy, 64 text tokens + "
                        f"{INSTRUCT_TOKENS}-token instruct, bs=1 per GPU -> {FRAMES} frames ({FRAMES*0.08:.1f} s) codes + 24 kHz wav",
            "size": args.size, "frames": FRAMES, "weights": "W8 affine g64 (1.0625 B/param)", "kv": "bf16 paged (16)",
            "l2": "inputs larger than L2: 1.5 GB of weights stream per token vs 126 MB L2", "parallelism": f"replica x{args.gpus}"}


# ------------------------------------------------------------------------------------------------------------------
# legs for the other BASELINE configs
# ------------------------------------------------------------------------------------------------------------------
def cfg2_leg(model, cfg, timed, bf16_peak):
    """BASELINE config 2: codec decoder alone, synthetic codes of a 30 s clip (375 frames), B = 1 and B = 32."""
    out = {}
    g = torch.Generator().manual_seed(2)
    for B in (1, 32):
        codes = torch.randint(0, cfg.codec.codebook_size, (B, cfg.codec.num_quantizers, 375), generator=g, dtype=torch.int32).cuda()
        model.codec.decode(codes)
        ms, wav = timed(lambda: model.codec.decode(codes), 3)
        flops = B * 375 * 4.96e9
        out[f"B{B}"] = {"ms_per_clip_batch": ms, "rtfx": B * 30.0 / (ms / 1e3), "tflops": flops / (ms / 1e3) / 1e12,
                        "frac_of_bf16_sustained": flops / (ms / 1e3) / 1e12 / bf16_peak, "samples": int(wav.shape[-1])}
    out["note"] = "chunked decode (300 + 75 frames, 25 frames of left context); every conv / linear on the tcgen05 tap-GEMM (fp16 operands between the vocoder layers, TF32 elsewhere)"
    return out


def cfg3_leg(model, cfg):
    """BASELINE config 3: Base voice cloning, synthetic prompt codes standing in for a 3 s reference clip (38 frames) +
    synthetic speaker vector + 200-token text (+ 8-token reference text), bs=1, streaming (text trails one token per frame),
    750 frames, codec per 25-frame interval with the chunked decode's left context, every piece copied to the host."""
    g = torch.Generator().manual_seed(3)
    n_ref, n_text, T = 38, 200, 750
    body = torch.randint(0, 151643, (n_text,), generator=g).tolist()
    ids = [cfg.im_start_id, cfg.assistant_id, 198] + body + [cfg.im_end_id, 198, cfg.im_start_id, cfg.assistant_id, 198]
    ref_ids = [cfg.im_start_id, cfg.assistant_id, 198] + torch.randint(0, 151643, (8,), generator=g).tolist() + [cfg.im_end_id, 198]
    ref_codes = torch.randint(0, cfg.codec.codebook_size, (n_ref, cfg.cp.num_code_groups), generator=g)
    vec = torch.randn(cfg.talker.hidden_size, generator=g) * 0.02
    pre, tr = model.build_prefill(ids, None, None, None, vec, True, ref_codes=ref_codes, ref_text_ids=ref_ids)
    res = {}
    for rep in range(2):                       # first pass warms the per-interval codec shapes
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        first_ms, n = None, 0
        for _c, w in model.stream_codes(pre, tr, T, 25, ref_codes=ref_codes):
            w.cpu()
            n += int(w.numel())
            if first_ms is None:
                first_ms = (time.perf_counter() - t0) * 1e3
        ms = (time.perf_counter() - t0) * 1e3
        res = {"frames": T, "ref_frames": n_ref, "text_tokens": n_text, "prompt_rows": int(pre.shape[0]), "trailing_rows": int(tr.shape[0]),
               "interval_frames": 25, "first_audio_ms": first_ms, "ms_total": ms, "rtfx": n / 24000.0 / (ms / 1e3), "audio_samples": n}
    res["note"] = "ICL prompt (SURVEY App. C), streaming; wall-clock incl. prefill, one codec call per interval, D2H of every piece"
    return res


def bs64_leg(cfg, args, world, dist, timed):
    """BASELINE config 4: 64 lock-step utterances per GPU, 300-token prompts (256 text + 32 instruct + control), `bs64_frames`
    frames each (960 by default: the context grows to ~1 260 tokens, K/V reads from 2.2 to 9.3 GB per step)."""
    from qwen3_tts_b200.codec import CodecDecoder
    from qwen3_tts_b200.engine import TalkerEngine
    from qwen3_tts_b200.weights import make_weights
    B, Lp, T = 64, 300, args.bs64_frames
    ws = make_weights(cfg, seed=0, device="cuda", keep_fp=False)
    eng = TalkerEngine(cfg, ws, "cuda", batch=B, max_frames=T + 8, max_ctx=((Lp + T + 24) // 16) * 16, attn_nsplit=4)
    codec = CodecDecoder(cfg, ws, "cuda")
    del ws
    eng.set_sampling(do_sample=False)
    emb = torch.randn(B, Lp, cfg.talker.hidden_size, device="cuda") * 0.02
    emb_host = emb.cpu().pin_memory()

    def step_dev():
        eng.prefill(emb, None, None)
        codes = eng.generate(T, check_every=0)
        return codec.decode(codes.transpose(1, 2).contiguous())

    def step_e2e():
        eng.prefill(emb_host.cuda(non_blocking=True), None, None)
        codes = eng.generate(T, check_every=0)
        return codec.decode(codes.transpose(1, 2).contiguous()).cpu()

    eng.prefill(emb, None, None)
    eng.generate(8, check_every=0)
    # the codec's chunk shapes (300 + 25 context frames) once before the timed run: the caching allocator grows by several GB
    codec.decode(torch.randint(0, cfg.codec.codebook_size, (B, cfg.codec.num_quantizers, min(T, 400)), device="cuda", dtype=torch.int32))
    ms, wav = timed(step_dev, 1)
    ms_e2e, wav_h = timed(step_e2e, 1)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    a, b, c, d, m1, m2 = ev(), ev(), ev(), ev(), ev(), ev()
    a.record(); eng.prefill(emb, None, None); b.record()
    eng.generate(min(32, T), check_every=0); m1.record()
    torch.cuda.synchronize()
    first32 = b.elapsed_time(m1) / min(32, T)
    audio_s = B * T * 0.08
    return {"value": world * audio_s / (ms / 1e3), "unit": UNIT, "batch_per_gpu": B, "prompt_tokens": Lp, "frames": T,
            "ctx_end": Lp + T, "e2e": {"value": world * audio_s / (ms_e2e / 1e3), "h2d_bytes_per_step": int(emb_host.numel() * 4),
                                        "d2h_bytes_per_step": int(wav_h.numel() * 4)},
            "ms_total": ms, "ms_prefill": a.elapsed_time(b), "ms_per_frame_first32": first32,
            "ms_per_frame_mean": (ms - a.elapsed_time(b)) / T, "launches_per_frame": eng.launches_per_frame,
            "note": "greedy, random-init; talker/CP contractions on the tcgen05 W8 GEMM (split-bf16 operands, split-K as thread-block clusters), codec "
                    "convolutions on the tcgen05 tap-GEMM (fp16 operands between the vocoder layers, TF32 elsewhere, fp32 accumulate); prefill and "
                    "codec counted inside the timed region (ms_per_frame_mean includes the codec)"}, eng, codec


def cfg5_leg(cfg, args, world, rank, dist, eng, codec):
    """BASELINE config 5: data-parallel long-form.  Independent 30 s utterances (375 frames) sharded over the ranks by
    dp.run_sharded (utterance-batch i -> rank i mod N), each rank running lock-step batches of 64 through its own engine.  The
    default run is a bounded sample (`cfg5_batches` batches of 64 per rank) of the 4096-utterance job; aggregate RTFx = audio
    seconds of all ranks / max-over-ranks wall time.  No data-path collective: results return through host memory."""
    from qwen3_tts_b200 import dp
    B, Lp, T = 64, 120, 375
    n_batches = args.cfg5_batches * world
    g = torch.Generator().manual_seed(5)
    batches = [torch.randn(B, Lp, cfg.talker.hidden_size, generator=g).mul_(0.02).pin_memory() for _ in range(min(n_batches, 2))]

    def run_batch(i):
        eng.prefill(batches[i % len(batches)].cuda(non_blocking=True), None, None)
        codes = eng.generate(T, check_every=0)
        wav = codec.decode(codes.transpose(1, 2).contiguous()).cpu()
        return int(wav.shape[0]), int(wav.shape[1])

    torch.cuda.synchronize()
    if dist is not None:
        dp.run_sharded(list(range(world)), lambda i: i, dist)      # first gather sets up the point-to-point channels: not part of the job
        dist.barrier()
    t0 = time.perf_counter()
    out = dp.run_sharded(list(range(n_batches)), run_batch, dist)
    torch.cuda.synchronize()
    sec = dp.max_over_ranks(time.perf_counter() - t0, dist, device="cuda")
    if rank != 0:
        return None
    n_utt = sum(o[0] for o in out)
    return {"utterances": n_utt, "of_job": 4096, "frames_each": T, "batches_per_rank": args.cfg5_batches, "seconds": sec,
            "rtfx_aggregate": n_utt * T * 0.08 / sec, "projected_seconds_for_4096": sec * 4096 / max(n_utt, 1),
            "note": "sharded by dp.run_sharded over the ranks, batches of 64 utterances x 120-token prompts x 375 frames, waveforms copied to the host"}


def serve_leg(cfg, args, world, dist, codec):
    """BASELINE config 4's "paged KV cache" made to mean something: `serve_requests` requests with ragged
    prompts (60..180 rows) and ragged frame budgets (100..500 frames; a random-init model never samples EOS, so the budget stands
    in for the utterance length) through the 64 SLOTS of qwen3_tts_b200.serving.ContinuousBatcher: K/V pages from a free list that
    cannot hold 64 worst-case requests, admission mid-flight (one ragged tcgen05 prefill per boundary), pages returned the moment
    a request ends; then the codec over the finished requests in length-sorted batches of 32.  Lock-step batches of 64 would run
    every batch to its longest member: `lockstep_slot_frames` is what that would have cost."""
    from qwen3_tts_b200.engine import TalkerEngine
    from qwen3_tts_b200.serving import ContinuousBatcher, Request
    from qwen3_tts_b200.weights import make_weights
    B, n_req, H = 64, args.serve_requests, cfg.talker.hidden_size
    g = torch.Generator().manual_seed(11)
    lens = torch.randint(60, 181, (n_req,), generator=g).tolist()
    budgets = torch.randint(100, 501, (n_req,), generator=g).tolist()
    per_seq = (180 + 500 + 8 + 15) // 16 + 1
    pool_pages = int(0.7 * B * per_seq)                        # 70 % of the worst case: admission also waits for pages
    ws = make_weights(cfg, seed=0, device="cuda", keep_fp=False)
    eng = TalkerEngine(cfg, ws, "cuda", batch=B, max_frames=512, max_ctx=per_seq * 16, kv_pages=pool_pages)
    del ws
    eng.set_sampling(do_sample=False)
    cb = ContinuousBatcher(eng, pool_pages=pool_pages, sync_every=8)
    tr = torch.zeros(1, H)
    reqs = [Request(i, torch.randn(lens[i], H, generator=g) * 0.02, tr, budgets[i]) for i in range(n_req)]
    cb.run([Request(10_000 + i, torch.randn(64, H, generator=g) * 0.02, tr, 16) for i in range(B)])      # warm-up: every slot once
    cb.stats = dict(admitted=0, retired=0, prefill_calls=0, frames=0, slot_frames_active=0)
    cb.frame = 0
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    done = cb.run(reqs)
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t0
    order = sorted(range(n_req), key=lambda i: budgets[i])
    n_samples = 0
    for o in range(0, n_req, 32):
        grp = order[o:o + 32]
        T = max(budgets[i] for i in grp)
        codes = torch.zeros(len(grp), cfg.codec.num_quantizers, T, dtype=torch.int32)
        for j, i in enumerate(grp):
            c = done[i].codes.clamp(0, cfg.codec.codebook_size - 1)
            codes[j, :, :c.shape[0]] = c.t()
        wav = codec.decode(codes.cuda()).cpu()
        n_samples += sum(int(done[i].codes.shape[0]) * cfg.codec.hop for i in grp)
    sec = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([sec], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t.item())
    audio_s = sum(budgets) * 0.08
    lock = sum(max(budgets[o:o + B]) * B for o in range(0, n_req, B))
    assert sorted(done) == list(range(n_req)) and cb.pool.free == pool_pages
    out = {"value": world * audio_s / sec, "unit": UNIT, "requests": n_req, "slots": B, "prompt_rows": [min(lens), max(lens)],
           "frame_budgets": [min(budgets), max(budgets)], "pool_pages": pool_pages, "worst_case_pages": B * per_seq,
           "peak_pages_used": cb.pool.peak_used, "seconds": sec, "seconds_generation": t_gen, "scheduler_frames": cb.stats["frames"],
           "prefill_calls": cb.stats["prefill_calls"], "slot_utilisation": cb.stats["slot_frames_active"] / max(1, cb.stats["frames"] * B),
           "slot_frames": sum(budgets), "lockstep_slot_frames": lock, "audio_samples": n_samples,
           "note": "continuous batching: per-slot frame counters, K/V pages from a free list (16 tokens each), ragged tcgen05 prefill at admission, "
                   "one host look at the done flags / counters every 8 frames; codec of the finished requests inside the timed region"}
    del cb, eng
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", default="full", choices=["full", "small"])
    ap.add_argument("--ref-frames", type=int, default=FRAMES, help="frames the reference arm really runs per step (240 = the whole utterance, no projection)")
    ap.add_argument("--cpu-frames", type=int, default=24)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-bs64", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the cfg2 / cfg3 / cfg5 legs")
    ap.add_argument("--bs64-frames", type=int, default=960)
    ap.add_argument("--serve-requests", type=int, default=160, help="requests of the continuous-batching leg")
    ap.add_argument("--cfg5-batches", type=int, default=1)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from mlx_audio.tts.generate import generate_audio
    from qwen3_tts_b200 import config as Cfg, lib as L
    from qwen3_tts_b200.model import Model
    from qwen3_tts_b200.weights import make_weights
    lib = L.load()
    cfg = getattr(Cfg, args.size)("custom_voice")
    ws = make_weights(cfg, seed=0, device="cuda", keep_fp=False)
    model = Model(cfg, ws, "cuda", max_frames=768, max_ctx=1024, max_trailing=256)
    del ws
    torch.cuda.empty_cache()
    e = model.engine
    e.set_sampling(do_sample=False)
    ids, ins = synth_ids(cfg)
    prefill, trailing = model.build_prefill(ids, ins, "ryan", "english")
    L0 = prefill.shape[0]
    audio_s = FRAMES * 0.08
    out_dir = tempfile.mkdtemp(prefix="q3t_bench_")

    def step_device():
        codes = model.generate_codes(prefill, trailing, FRAMES)
        return model.decode(codes)

    def step_e2e():
        # what a reference session does per utterance (custom.py:163-170): host strings in, audio_000.wav out
        return generate_audio(model=model, text=TEXT, voice="ryan", instruct=INSTRUCT, speed=1.0, lang_code="english",
                              output_path=out_dir, greedy=True, max_tokens=FRAMES)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(k):
            out = fn()
        t.record()
        barrier()
        ms = s.elapsed_time(t)
        if dist is not None:
            tt = torch.tensor([ms], device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms / k, out

    for _ in range(max(args.warmup, 1)):
        wav = step_device()
    n_codec0 = lib.q3t_launch_count()
    model.decode(e.codes[0, :FRAMES])
    codec_launches = int(lib.q3t_launch_count() - n_codec0)
    n_pre0 = lib.q3t_launch_count()                   # prompt rows on the tcgen05 GEMM (eager launches, counted by the library)
    e.prefill(prefill[None], None, trailing[None])
    prefill_launches = int(lib.q3t_launch_count() - n_pre0)
    reps = 200
    with ClockSampler(local) as clk:
        ms_dev, wav = timed(step_device, args.steps)
        # talker decode step alone (the north-star roofline): graph of one token at ctx ~ L0 + FRAMES/2
        e.pos.fill_(L0 + FRAMES // 2)

        def talker_steps():
            for _ in range(reps):
                e._graphs["step_logits"].replay()
            e.pos.fill_(L0 + FRAMES // 2)
        ms_tok, _ = timed(talker_steps, 1)
        ms_tok /= reps
        # the whole frame launch at the same context (16 code-predictor passes + talker step), greedy
        e.reset(); e.pos.fill_(L0 + FRAMES // 2)

        def frames():
            for _ in range(reps):
                e._graphs["frame"].replay()
            e.reset(); e.pos.fill_(L0 + FRAMES // 2)
        ms_frame, _ = timed(frames, 1)
        ms_frame /= reps
    for _ in range(max(args.warmup, 1)):
        step_e2e()
    ms_e2e, wav_path = timed(step_e2e, args.steps)
    text_ids_bytes = 8 * (len(model.chat_ids(TEXT)) + len(model.instruct_ids(INSTRUCT)))

    value = world * audio_s / (ms_dev / 1e3)
    e2e = world * audio_s / (ms_e2e / 1e3)
    ctx = L0 + FRAMES // 2
    step_b, w_b, kv_b = talker_step_bytes(cfg, ctx)
    peak, bf16_peak, peak_src = peaks()
    achieved = step_b / (ms_tok / 1e3) / 1e9
    frame_b = step_b + cp_frame_bytes_l2_resident(cfg)
    traffic, traffic_src = None, None    # dram read+write bytes of the same launch from the committed ncu --set full capture
    for name in ("r02_ncu_full_frame_ll_talker_step.json", "r01_ncu_full_frame_ll_talker_step.json"):
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", name)))["traffic_bytes"]
            traffic_src = f"profiles/{name} (ncu --set full)"
            break
        except Exception:
            pass
    launches_step = e.launches_per_frame
    # launches inside one timed step: prefill + FRAMES frames (counted at graph capture) + the codec
    if prefill_launches == 0:                          # token-by-token prefill replays the captured step graph
        prefill_launches = L0 * e.launches.get("step", 0)
    gpu_launches = args.steps * (prefill_launches + FRAMES * (launches_step or 0) + codec_launches)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "w8a32(f32 accumulate)",
            "data": "synthetic", "config": workload_config(args),
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(text_ids_bytes), "d2h_bytes_per_step": int(wav.numel() * 4 + FRAMES * 16 * 4),
                    "ms_per_step": ms_e2e, "api": "mlx_audio.tts.generate.generate_audio(model=, text=, voice=, instruct=, output_path=) -> audio_000.wav"},
            "gpu_launches": int(gpu_launches), "launches_per_frame": launches_step,
            "roofline": {"bound": "hbm", "kernel": "frame_ll_kernel, stack mode (talker decode step = ONE persistent launch: 28 layers + final norm + codec head)",
                         "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                         "frac_of_8TBps": achieved / 8000.0, "traffic": traffic, "traffic_source": traffic_src, "bytes_per_step": step_b,
                         "weight_bytes": w_b, "kv_bytes": kv_b, "us_per_talker_step": ms_tok * 1e3, "ctx": ctx},
            "roofline_frame": {"bound": "hbm", "kernel": "frame_ll_kernel, frame mode (sample + 16 code-predictor passes + next-input sum + talker step)",
                               "achieved": frame_b / (ms_frame / 1e3) / 1e9, "peak": peak, "unit": "GB/s", "frac": frame_b / (ms_frame / 1e3) / 1e9 / peak,
                               "bytes_per_frame": frame_b, "us_per_frame": ms_frame * 1e3, "us_code_predictor": (ms_frame - ms_tok) * 1e3,
                               "note": "code-predictor weights (0.119 GB W8) counted once per frame: L2-resident across the 16 passes"},
            "clocks": clk.summary(), "frames_per_s": world * FRAMES / (ms_dev / 1e3), "audio_samples": int(wav.numel())}
    if not args.no_extra:
        line["cfg2"] = cfg2_leg(model, cfg, timed, bf16_peak)
        line["cfg3"] = cfg3_leg(model, cfg)
    if not args.no_bs64:
        # ---- batch-64 serving leg (BASELINE config 4): tcgen05 GEMM prefill + batched frame graph + batched codec
        del model, e
        torch.cuda.empty_cache()
        line["bs64"], eng, codec = bs64_leg(cfg, args, world, dist, timed)
        if not args.no_extra:
            c5 = cfg5_leg(cfg, args, world, rank, dist, eng, codec)
            if c5 is not None:
                line["cfg5"] = c5
        del eng
        torch.cuda.empty_cache()
        if not args.no_extra:
            sv = serve_leg(cfg, args, world, dist, codec)
            if rank == 0:
                line["serve"] = sv
        del codec
        torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, detail = CpuOracle(args.size).sample(args.cpu_frames)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                "sample": cpu_sample_text(args.cpu_frames), **detail}
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
