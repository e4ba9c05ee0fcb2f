#!/usr/bin/env python
"""bench.py -- headline metric of BASELINE.json on B200: RTFx (audio s / wall s) of the Qwen3-TTS-12Hz-1.7B
generation hot path at bs=1, with the talker-decode HBM roofline and the CPU oracle timed beside it.

A "step" = one utterance of BASELINE config 1 (CustomVoice, random-init, greedy, 64 text tokens + 8-token
instruct + speaker/language ids): talker prefill -> 240 frames of [talker decode, sample, 15 code-predictor
passes, next-input sum] -> codec decode of the 240 frames to a 24 kHz waveform (19.2 s of audio).

  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun, one replica per GPU)
  python bench.py --impl reference ...                       (the CPU oracle on the host cores; rank 0 only)

`value`  : device-resident inputs (prefill embeddings already in HBM), CUDA-event time, max over ranks.
`e2e`    : same metric through the public API with HOST inputs (pinned token ids -> H2D -> ... -> waveform D2H).
`roofline`: talker decode step (one persistent data-flow launch, csrc/frame_ll.cu): algorithmic bytes per step
             / CUDA-event step time, against MEASURED_PEAKS.json hbm_gbs.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
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
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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


# ------------------------------------------------------------------------------------------------------------------
def cpu_oracle_rtfx(cfg_name, frames_sample, threads=None):
    """The CPU oracle (kind 'port') on the host cores: prefill + `frames_sample` frames + codec of those frames,
    projected to the 240-frame utterance (prefill counted once)."""
    from oracle import qwen3_tts_oracle as O
    from qwen3_tts_b200 import config as Cfg
    from qwen3_tts_b200.weights import dequantize_w8, make_weights
    if threads:
        torch.set_num_threads(threads)
    cfg = getattr(Cfg, cfg_name)("custom_voice")
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    ws = make_weights(cfg, seed=0, device=dev, keep_fp=True, keep_q=False)       # generate fast, then move to host
    w = {k: v.cpu() for k, v in ws.fp.items()}
    del ws
    m = O.OracleModel(cfg, w, kv_dtype=torch.bfloat16)
    ids, ins = synth_ids(cfg)
    with torch.no_grad():
        t0 = time.perf_counter()
        pre, tr = m.build_prefill(ids, instruct_ids=ins, speaker="ryan", language="english")
        m.talker.reset()
        h, lg = m.talker_forward(pre)
        t_prefill = time.perf_counter() - t0
        t0 = time.perf_counter()
        codes = m.generate(pre, tr, frames_sample)          # re-runs the prefill; subtract it below
        t_gen = time.perf_counter() - t0 - t_prefill
        t0 = time.perf_counter()
        wav = O.codec_forward(w, cfg, codes.t()[None])
        t_codec = time.perf_counter() - t0
    per_frame = (t_gen + t_codec) / frames_sample
    total = t_prefill + FRAMES * per_frame
    return (FRAMES * 0.08) / total, dict(prefill_s=t_prefill, gen_s_per_frame=t_gen / frames_sample,
                                         codec_s_per_frame=t_codec / frames_sample)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    vals = []
    for i in range(args.warmup + args.steps):
        v, detail = cpu_oracle_rtfx(args.size, args.ref_frames)
        if i >= args.warmup:
            vals.append(v)
        if i == 0 and args.steps + args.warmup > 1:
            pass
    v = statistics.mean(vals)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": FRAMES * 80.0 / v, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"PyTorch-CPU oracle: prefill + {args.ref_frames} frames + codec of those frames, "
                                       f"projected to {FRAMES} frames", **detail},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(args):
    return {"workload": "BASELINE config 1: Qwen3-TTS-12Hz-1.7B-CustomVoice random-init, greedy, 64 text tokens + "
                        f"{INSTRUCT_TOKENS}-token instruct, bs=1 per GPU -> {FRAMES} frames ({FRAMES*0.08:.1f} s) codes + 24 kHz wav",
            "size": args.size, "frames": FRAMES, "weights": "W8 affine g64 (1.0625 B/param)", "kv": "bf16 paged (16)",
            "l2": "inputs larger than L2: 1.5 GB of weights stream per token vs 126 MB L2", "parallelism": f"replica x{args.gpus}"}


def bs64_leg(cfg, args, world, dist, timed):
    """64 lock-step utterances per GPU: 300-token prompts (256 text + 32 instruct + control), B64_FRAMES frames each."""
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
    state = {}

    def step_dev():
        eng.prefill(emb, None, None)
        codes = eng.generate(T, check_every=0)
        return codec.decode(codes.transpose(1, 2).contiguous())

    def step_e2e():
        eng.prefill(emb_host.cuda(non_blocking=True), None, None)
        codes = eng.generate(T, check_every=0)
        return codec.decode(codes.transpose(1, 2).contiguous()).cpu()

    step_dev()
    ms, wav = timed(step_dev, max(1, min(args.steps, 2)))
    ms_e2e, wav_h = timed(step_e2e, 1)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    a, b, c, d = ev(), ev(), ev(), ev()
    a.record(); eng.prefill(emb, None, None); b.record(); codes = eng.generate(T, check_every=0); c.record()
    codec.decode(codes.transpose(1, 2).contiguous()); d.record(); torch.cuda.synchronize()
    audio_s = B * T * 0.08
    return {"value": world * audio_s / (ms / 1e3), "unit": UNIT, "batch_per_gpu": B, "prompt_tokens": Lp, "frames": T,
            "e2e": {"value": world * audio_s / (ms_e2e / 1e3), "h2d_bytes_per_step": int(emb_host.numel() * 4),
                    "d2h_bytes_per_step": int(wav_h.numel() * 4)},
            "ms_prefill": a.elapsed_time(b), "ms_per_frame": b.elapsed_time(c) / T, "ms_codec": c.elapsed_time(d),
            "launches_per_frame": eng.launches_per_frame,
            "note": "greedy, random-init; talker/CP contractions on the tcgen05 W8 GEMM (bf16 operands), codec convolutions on the "
                    "tcgen05 TF32 tap-GEMM; prefill counted inside the timed region"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", default="full", choices=["full", "small"])
    ap.add_argument("--ref-frames", type=int, default=4)
    ap.add_argument("--cpu-frames", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-bs64", action="store_true")
    ap.add_argument("--bs64-frames", type=int, default=96)
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

    from qwen3_tts_b200 import config as Cfg, lib as L
    from qwen3_tts_b200.model import Model
    from qwen3_tts_b200.weights import make_weights
    lib = L.load()
    cfg = getattr(Cfg, args.size)("custom_voice")
    ws = make_weights(cfg, seed=0, device="cuda", keep_fp=False)
    model = Model(cfg, ws, "cuda", max_frames=FRAMES, max_ctx=512, max_trailing=1)
    del ws
    torch.cuda.empty_cache()
    e = model.engine
    e.set_sampling(do_sample=False)
    ids, ins = synth_ids(cfg)
    ids_host = torch.tensor(ids, dtype=torch.int64).pin_memory()
    ins_host = torch.tensor(ins, dtype=torch.int64).pin_memory()
    prefill, trailing = model.build_prefill(ids, ins, "ryan", "english")
    L0 = prefill.shape[0]
    audio_s = FRAMES * 0.08

    def step_device():
        codes = model.generate_codes(prefill, trailing, FRAMES)
        return model.decode(codes)

    def step_e2e():
        i_d = ids_host.cuda(non_blocking=True)
        n_d = ins_host.cuda(non_blocking=True)
        pre, tr = model.build_prefill(i_d.tolist(), n_d.tolist(), "ryan", "english")
        codes = model.generate_codes(pre, tr, FRAMES)
        wav = model.decode(codes)
        return wav.cpu()

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
    with ClockSampler(local) as clk:
        ms_dev, wav = timed(step_device, args.steps)
        # talker decode step alone (the north-star roofline): graph of one token at ctx ~ L0 + FRAMES/2
        e.pos.fill_(L0 + FRAMES // 2)
        reps = 200

        def talker_steps():
            for _ in range(reps):
                e._graphs["step_logits"].replay()
            e.pos.fill_(L0 + FRAMES // 2)
        ms_tok, _ = timed(talker_steps, 1)
        ms_tok /= reps
    for _ in range(max(args.warmup, 1)):
        step_e2e()
    ms_e2e, wav_host = timed(step_e2e, args.steps)

    # streaming (BASELINE config 3 mechanics on this workload): one codec call per 25 frames, every piece read back to the host
    import time as _time
    torch.cuda.synchronize()
    t0 = _time.perf_counter()
    first_ms, n_stream = None, 0
    for _c, _w in model.stream_codes(prefill, trailing, FRAMES, 25):
        _w.cpu()
        n_stream += int(_w.numel())
        if first_ms is None:
            first_ms = (_time.perf_counter() - t0) * 1e3
    stream_ms = (_time.perf_counter() - t0) * 1e3
    streaming = {"interval_frames": 25, "first_audio_ms": first_ms, "rtfx": n_stream / 24000.0 / (stream_ms / 1e3),
                 "note": "prefill + frames + one codec call per interval with 25 frames of left context, each piece copied to the host"}

    value = world * audio_s / (ms_dev / 1e3)
    e2e = world * audio_s / (ms_e2e / 1e3)
    step_b, w_b, kv_b = talker_step_bytes(cfg, L0 + FRAMES // 2)
    peak, peak_src = peaks()
    achieved = step_b / (ms_tok / 1e3) / 1e9
    traffic = None                       # dram read+write bytes of the same launch from the committed ncu --set full capture
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_full_frame_ll_talker_step.json")))["traffic_bytes"]
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
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(ids_host.numel() * 8 + ins_host.numel() * 8),
                    "d2h_bytes_per_step": int(wav_host.numel() * 4), "ms_per_step": ms_e2e},
            "gpu_launches": int(gpu_launches), "launches_per_frame": launches_step,
            "roofline": {"bound": "hbm", "kernel": "frame_ll_kernel, stack mode (talker decode step = ONE persistent launch: 28 layers + final norm + codec head)",
                         "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                         "frac_of_8TBps": achieved / 8000.0, "traffic": traffic,
                         "traffic_source": "profiles/r01_ncu_full_frame_ll_talker_step.json (ncu --set full, ctx=300)", "bytes_per_step": step_b, "weight_bytes": w_b,
                         "kv_bytes": kv_b, "us_per_talker_step": ms_tok * 1e3, "ctx": L0 + FRAMES // 2},
            "clocks": clk.summary(), "frames_per_s": world * FRAMES / (ms_dev / 1e3), "audio_samples": int(wav.numel()),
            "streaming": streaming}
    if not args.no_bs64:
        # ---- batch-64 serving leg (BASELINE config 4 shapes): tcgen05 GEMM prefill + batched frame graph + batched codec
        del model, e
        torch.cuda.empty_cache()
        line["bs64"] = bs64_leg(cfg, args, world, dist, timed)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, detail = cpu_oracle_rtfx(args.size, args.cpu_frames)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"PyTorch-CPU oracle: prefill + {args.cpu_frames} frames + codec of those frames, "
                                          f"projected to {FRAMES} frames", **detail}
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
