"""`Model` = the object `load_model(path)` returns and `generate_audio(model=...)` drives.

Keeps the reference's surface: `load_model(path).generate(text, voice / instruct / ref_audio, ...)` yields
result objects carrying `audio`, `sample_rate`, ... (SURVEY.md 8b "Output contract"); the three session
kinds of the reference map to `tts_model_type`:
    custom_voice  sessions/custom.py:163-170   generate(text, voice=<speaker>, instruct=<emotion>, speed=)
    voice_design  sessions/design.py:76-81     generate(text, instruct=<voice description>)
    base          sessions/clone.py:218-224    generate(text, ref_audio=<wav path>, ref_text=)
"""
from __future__ import annotations

import json
import os
import time
import wave
from dataclasses import dataclass
from typing import Iterator, List, Optional, Sequence

import numpy as np
import torch

from . import config as cfgmod
from .codec import CodecDecoder
from .config import ModelConfig
from .engine import TalkerEngine
from .text import segment_text
from .weights import WeightStore, make_weights


@dataclass
class GenerationResult:
    audio: np.ndarray               # float32 mono in [-1, 1]
    sample_rate: int
    samples: int
    segment_idx: int
    token_count: int                # frames generated (12.5 Hz)
    audio_duration: float
    processing_time_seconds: float
    real_time_factor: float         # audio seconds / wall seconds (RTFx)
    codes: Optional[np.ndarray] = None


class ByteTokenizer:
    """Stand-in text front end for RANDOM-INIT runs only (bench, parity tests: the Qwen2 BPE vocabulary files are not
    available offline): text is mapped byte-wise into the text vocabulary.  A folder that holds real weights
    never gets this tokenizer - `_load_tokenizer` raises instead (a hashed id stream through trained weights is noise)."""

    def __init__(self, cfg: ModelConfig):
        self.cfg = cfg

    def encode(self, text: str) -> List[int]:
        v = self.cfg.talker.text_vocab_size
        return [(b * 2654435761 + 17) % max(v - 16, 1) for b in text.encode("utf-8")]


# the Qwen2 pre-tokenizer split (tokenization_qwen2.PRETOKENIZE_REGEX in the transformers cousin)
_QWEN2_SPLIT = r"""(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+"""


class _HFTokenizer:
    def __init__(self, tok):
        self.tok = tok

    def encode(self, text: str) -> List[int]:
        return self.tok.encode(text, add_special_tokens=False).ids


def _load_tokenizer(path: Optional[str], cfg: ModelConfig, require: bool = False):
    """`tokenizer.json` (fast-tokenizer file) or `vocab.json` + `merges.txt` (the files Qwen checkpoints ship: byte-level
    BPE with the Qwen2 pre-tokenizer split) from the model folder.  `require`: the folder holds real weights, so a missing
    tokenizer is an error (ValueError -> the reference prints "Failed to load model", io.py:115-117)."""
    if path:
        tj, vj, mt = (os.path.join(path, n) for n in ("tokenizer.json", "vocab.json", "merges.txt"))
        if os.path.exists(tj):
            from tokenizers import Tokenizer
            return _HFTokenizer(Tokenizer.from_file(tj))
        if os.path.exists(vj) and os.path.exists(mt):
            from tokenizers import Regex, Tokenizer, decoders, models, pre_tokenizers
            tok = Tokenizer(models.BPE.from_file(vj, mt))
            tok.pre_tokenizer = pre_tokenizers.Sequence([
                pre_tokenizers.Split(Regex(_QWEN2_SPLIT), behavior="isolated", invert=False),
                pre_tokenizers.ByteLevel(add_prefix_space=False, use_regex=False)])
            tok.decoder = decoders.ByteLevel()
            return _HFTokenizer(tok)
    if require:
        raise ValueError(f"{path}: neither tokenizer.json nor vocab.json + merges.txt found next to the weights")
    return ByteTokenizer(cfg)


class Model:
    def __init__(self, cfg: ModelConfig, ws: WeightStore, device: str = "cuda", model_path: Optional[str] = None,
                 max_frames: int = 2048, max_ctx: int = 4096, batch: int = 1, max_trailing: int = 1024,
                 prefill: str = "auto", require_tokenizer: bool = False):
        if not torch.cuda.is_available():
            raise RuntimeError("qwen3_tts_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.cfg = cfg
        self.sample_rate = cfg.codec.sample_rate
        self.device = device
        self.engine = TalkerEngine(cfg, ws, device, batch=batch, max_frames=max_frames, max_ctx=max_ctx,
                                   max_trailing=max_trailing, prefill=prefill)
        self.codec = CodecDecoder(cfg, ws, device)
        self.tokenizer = _load_tokenizer(model_path, cfg, require=require_tokenizer)
        # reference-clip encoders (voice cloning): built on first use from the enc.* / spk.* tensors of the store
        self._ref_weights = {n: t for n, t in ws.fp.items() if n.startswith("enc.") or n.startswith("spk.")}
        self._speech_encoder = self._speaker_encoder = None
        self._ref_cache = {}

    # ---- prompt assembly (mirrors oracle.OracleModel.build_prefill) -------------------------
    def chat_ids(self, text: str) -> List[int]:
        """<|im_start|>assistant\\n{text}<|im_end|>\\n<|im_start|>assistant\\n  ->  ids[:3] prefix, ids[-5:] tail."""
        c = self.cfg
        nl = self.tokenizer.encode("\n")[:1] or [10]
        return [c.im_start_id, c.assistant_id, nl[0]] + list(self.tokenizer.encode(text)) + \
               [c.im_end_id, nl[0], c.im_start_id, c.assistant_id, nl[0]]

    def instruct_ids(self, instruct: str) -> List[int]:
        c = self.cfg
        nl = self.tokenizer.encode("\n")[:1] or [10]
        user = self.tokenizer.encode("user")[:1] or [11]
        return [c.im_start_id, user[0], nl[0]] + list(self.tokenizer.encode(instruct)) + [c.im_end_id, nl[0]]

    def ref_text_chat_ids(self, ref_text: str) -> List[int]:
        """<|im_start|>assistant\n{ref_text}<|im_end|>\n  ->  ids[:3] prefix, ids[3:-2] body."""
        c = self.cfg
        nl = self.tokenizer.encode("\n")[:1] or [10]
        return [c.im_start_id, c.assistant_id, nl[0]] + list(self.tokenizer.encode(ref_text)) + [c.im_end_id, nl[0]]

    def ref_code_embeds(self, ref_codes: torch.Tensor) -> torch.Tensor:
        """[T_ref, G] -> [T_ref, H]: the 16-way embedding sum of every reference frame, added in order g = 0..15 (a8)."""
        e = self.engine
        rc = ref_codes.to(e.dev).long()
        acc = e.codec_embedding[rc[:, 0]]
        for g in range(1, rc.shape[1]):
            acc = acc + e.cp_embeddings[g - 1][rc[:, g]]
        return acc

    def build_prefill(self, text_ids: Sequence[int], instruct_ids: Optional[Sequence[int]] = None,
                      speaker: Optional[str] = None, language: Optional[str] = None,
                      speaker_vec: Optional[torch.Tensor] = None, streaming: bool = False,
                      ref_codes: Optional[torch.Tensor] = None, ref_text_ids: Optional[Sequence[int]] = None):
        """Prompt embeddings; with `ref_codes` the in-context (voice cloning) layout: the text side
        P(ref text ++ text) ++ eos paired with E(codec_bos) ++ per-frame code embedding sums of the reference clip
        (reference call site sessions/clone.py:218-224; mirrors oracle.OracleModel.build_prefill)."""
        e, t, cfg = self.engine, self.cfg.talker, self.cfg
        dev = e.dev
        ids = torch.as_tensor(list(text_ids), dtype=torch.long, device=dev)
        special = e.text_embed(torch.tensor([cfg.tts_pad_token_id, cfg.tts_bos_token_id, cfg.tts_eos_token_id]))
        pad, bos, eos = special[0], special[1], special[2]
        cemb = lambda lst: e.codec_embedding[torch.as_tensor(lst, dtype=torch.long, device=dev)]
        if language is not None and language.lower() in t.codec_language_id:
            prefix = [t.codec_think_id, t.codec_think_bos_id, t.codec_language_id[language.lower()], t.codec_think_eos_id]
        else:
            prefix = [t.codec_nothink_id, t.codec_think_bos_id, t.codec_think_eos_id]
        parts = [cemb(prefix)]
        if speaker_vec is not None:
            parts.append(speaker_vec.to(dev, torch.float32).view(1, -1))
        elif speaker is not None:
            parts.append(cemb([t.spk_id[speaker.lower()]]))
        parts.append(cemb([t.codec_pad_id, t.codec_bos_id]))
        codec_seq = torch.cat(parts, 0)
        n = codec_seq.shape[0]
        # one batched projection of every text-side id used below
        body_ids = ids[3:-5]
        segs = []
        if instruct_ids is not None and len(instruct_ids):
            segs.append(e.text_embed(torch.as_tensor(list(instruct_ids), dtype=torch.long)))
        head = e.text_embed(ids[:3])
        mid = torch.cat([pad.expand(n - 2, -1), bos[None]], 0) + codec_seq[:-1]
        segs += [head, mid]
        if ref_codes is not None:
            ref_body = torch.as_tensor(list(ref_text_ids), dtype=torch.long, device=dev)[3:-2] if ref_text_ids is not None \
                else torch.zeros(0, dtype=torch.long, device=dev)
            text_all = torch.cat([e.text_embed(torch.cat([ref_body, body_ids])), eos[None]], 0)
            codec_all = torch.cat([cemb([t.codec_bos_id]), self.ref_code_embeds(ref_codes)], 0)
            t1, t2 = text_all.shape[0], codec_all.shape[0]
            if streaming:
                if t1 > t2:
                    segs.append(text_all[:t2] + codec_all)
                    trailing = torch.cat([text_all[t2:], pad[None]], 0)
                else:
                    segs.append(torch.cat([text_all, pad[None].expand(t2 - t1, -1)], 0) + codec_all)
                    trailing = pad[None]
            else:
                segs.append(text_all + cemb([t.codec_pad_id] * t1))
                segs.append(codec_all + pad[None])
                trailing = pad[None]
            return torch.cat(segs, 0), trailing
        body_e = e.text_embed(body_ids) if len(body_ids) else torch.zeros(0, t.hidden_size, device=dev)
        if not streaming:
            body = torch.cat([body_e, eos[None]], 0) + cemb([t.codec_pad_id] * (len(body_ids) + 1))
            tail = pad[None] + cemb([t.codec_bos_id])
            segs += [body, tail]
            trailing = pad[None]
        else:
            segs.append(body_e[:1] + codec_seq[-1:])
            trailing = torch.cat([body_e[1:], eos[None], pad[None]], 0)
        return torch.cat(segs, 0), trailing

    # ---- low-level: ids -> codes -> wav -----------------------------------------------------------------------
    def generate_codes(self, prefill: torch.Tensor, trailing: torch.Tensor, max_frames: int) -> torch.Tensor:
        """prefill [L, H], trailing [n, H] -> codes [T, 16] (int32, trimmed at EOS)."""
        e = self.engine
        e.prefill(prefill[None], None, trailing[None])
        codes = e.generate(max_frames)[0]
        eos = (codes[:, 0] == self.cfg.talker.codec_eos_id).nonzero()
        if eos.numel():
            codes = codes[: int(eos[0, 0])]
        return codes

    def stream_codes(self, prefill: torch.Tensor, trailing: torch.Tensor, max_frames: int, interval: int = 25,
                     ref_codes: Optional[torch.Tensor] = None):
        """Streaming generation (BASELINE config 3): frames are produced one persistent launch at a time and every
        `interval` frames the codec decodes the new ones with its left context.  Yields (codes [n, 16], wav [n * hop])
        per interval; the concatenated pieces are bit-identical to generate_codes() followed by
        CodecDecoder.decode(chunk_size=interval).  One host synchronisation per interval (the EOS check).
        `ref_codes` [T_ref, 16] (voice cloning): the reference frames precede the generated ones in the codec's left context, as in
        the offline path that decodes ref ++ new and cuts the reference span off."""
        e = self.engine
        n_ref = 0 if ref_codes is None else int(ref_codes.shape[0])
        ref_t = None if ref_codes is None else ref_codes.to(e.dev, torch.int32).t()[None].contiguous()
        assert max_frames <= e.max_frames and interval >= 1
        e.prefill(prefill[None], None, trailing[None])
        eos_id = self.cfg.talker.codec_eos_id
        start = 0
        while start < max_frames:
            end = min(start + interval, max_frames)
            for _ in range(end - start):
                e._run("frame")
            new = e.codes[0, start:end]
            eos = (new[:, 0] == eos_id).nonzero()
            stop = eos.numel() > 0
            if stop:
                end = start + int(eos[0, 0])
            if end > start:
                cur = e.codes[0, :end].t()[None].contiguous()
                if ref_t is not None:
                    cur = torch.cat([ref_t, cur], -1)
                wav = self.codec.decode_interval(cur, n_ref + start, n_ref + end)[0]
                yield e.codes[0, start:end], wav
            if stop:
                return
            start = end

    def decode(self, codes: torch.Tensor) -> torch.Tensor:
        """codes [T, 16] -> wav [n] on device."""
        if codes.shape[0] == 0:
            return torch.zeros(0, device=self.engine.dev)
        return self.codec.decode(codes.t()[None].contiguous())[0]

    # ---- the reference-facing surface ------------------------------------------------------------------------------
    def generate(self, text: str, voice: Optional[str] = None, instruct: Optional[str] = None, speed: float = 1.0,
                 lang_code: str = "auto", ref_audio: Optional[str] = None, ref_text: Optional[str] = None,
                 temperature: Optional[float] = None, top_k: int = 50, top_p: float = 1.0,
                 repetition_penalty: float = 1.05, max_tokens: int = 1200, seed: Optional[int] = None, verbose: bool = False,
                 stream: bool = False, streaming_interval: float = 2.0, max_segment_chars: int = 600,
                 **kwargs) -> Iterator[GenerationResult]:
        """One utterance -> one result (the reference consumes only audio_000.wav, io.py:156); with `stream=True` one
        result per `streaming_interval` seconds of audio as it is generated (segment_idx counts the pieces).  A text
        longer than `max_segment_chars` is cut at sentence boundaries (text.segment_text) and generated segment by
        segment, one result each (the shim joins them); 0 disables the segmentation.
        `speed` is accepted and ignored exactly like an unknown library kwarg; `temperature=0`
        or `None` with greedy=True selects the greedy parity path."""
        t0 = time.perf_counter()
        mode = self.cfg.tts_model_type
        greedy = kwargs.pop("greedy", False) or (temperature is not None and temperature <= 0)
        if seed is None:
            # like the reference stack's global RNG, two sampled calls with the same text differ; pass `seed` to reproduce one
            self._auto_seed = (getattr(self, "_auto_seed", None) or int.from_bytes(os.urandom(4), "little")) + 1
            seed = 0 if greedy else self._auto_seed
        self.engine.set_sampling(do_sample=not greedy, temperature=temperature or 0.9, top_k=top_k, top_p=top_p,
                                 repetition_penalty=repetition_penalty, seed=seed)
        language = None if lang_code in (None, "auto") else lang_code
        speaker, speaker_vec, streaming, ref_codes, ref_ids = None, None, False, None, None
        ins = self.instruct_ids(instruct) if instruct else None
        if mode == "custom_voice":
            if voice is not None and voice.lower() not in self.cfg.talker.spk_id:
                raise ValueError(f"unknown speaker '{voice}'")
            speaker = voice
        elif mode == "base" and ref_audio is not None:
            # in-context cloning: codes of the clip + its transcript in the prompt, speaker vector in the speaker slot
            ref_codes, speaker_vec = self._reference_prompt(ref_audio)
            ref_ids = self.ref_text_chat_ids(ref_text if ref_text else ".")          # the reference passes "." without a transcript (clone.py:148-150)
            streaming = True
        segments = segment_text(text, max_segment_chars)
        if len(segments) > 1:
            kw = dict(voice=voice, instruct=instruct, speed=speed, lang_code=lang_code, ref_audio=ref_audio, ref_text=ref_text,
                      temperature=temperature, top_k=top_k, top_p=top_p, repetition_penalty=repetition_penalty,
                      max_tokens=max_tokens, seed=seed, verbose=verbose, stream=stream, streaming_interval=streaming_interval,
                      max_segment_chars=0, greedy=greedy, **kwargs)
            idx = 0
            for seg in segments:
                for r in self.generate(seg, **kw):
                    r.segment_idx = idx
                    idx += 1
                    yield r
            return
        ids = self.chat_ids(text)
        prefill, trailing = self.build_prefill(ids, ins, speaker, language, speaker_vec, streaming, ref_codes=ref_codes, ref_text_ids=ref_ids)
        if trailing.shape[0] > self.engine.max_trailing:
            raise ValueError(f"text of {trailing.shape[0]} trailing tokens exceeds max_trailing={self.engine.max_trailing}")
        max_frames = min(max_tokens, self.engine.max_frames)
        if stream:
            interval = max(1, int(round(streaming_interval * self.sample_rate / self.cfg.codec.hop)))
            for i, (c, w) in enumerate(self.stream_codes(prefill, trailing, max_frames, interval, ref_codes=ref_codes)):
                audio = w.float().cpu().numpy()
                dt = time.perf_counter() - t0
                dur = audio.shape[0] / self.sample_rate
                yield GenerationResult(audio=audio, sample_rate=self.sample_rate, samples=int(audio.shape[0]), segment_idx=i,
                                       token_count=int(c.shape[0]), audio_duration=dur, processing_time_seconds=dt,
                                       real_time_factor=(dur / dt if dt > 0 else 0.0), codes=c.cpu().numpy())
                t0 = time.perf_counter()
            return
        codes = self.generate_codes(prefill, trailing, max_frames)
        if ref_codes is not None and codes.shape[0] > 0:
            # decode ref ++ new and cut the reference span off the front (proportional cut, SURVEY App. C last line)
            n_ref = int(ref_codes.shape[0])
            wav = self.decode(torch.cat([ref_codes.to(codes.device, codes.dtype), codes], 0))
            wav = wav[int(n_ref / (n_ref + codes.shape[0]) * wav.shape[0]):]
        else:
            wav = self.decode(codes)
        audio = wav.float().cpu().numpy()
        dt = time.perf_counter() - t0
        dur = audio.shape[0] / self.sample_rate
        yield GenerationResult(audio=audio, sample_rate=self.sample_rate, samples=int(audio.shape[0]), segment_idx=0,
                               token_count=int(codes.shape[0]), audio_duration=dur, processing_time_seconds=dt,
                               real_time_factor=(dur / dt if dt > 0 else 0.0), codes=codes.cpu().numpy())

    def _reference_prompt(self, ref_audio: str):
        """Voice-cloning prompt of the clip at `ref_audio` (24 kHz mono PCM16, what the reference hands over: clone.py:140,158;
        io.py:243-247): (codes [T_ref, 16] from the speech-tokenizer encoder, speaker vector [H] from the ECAPA speaker encoder).
        Both run on the GPU (qwen3_tts_b200/encoders.py); the result is cached per file so the segments of a long text share it."""
        st = os.stat(ref_audio)
        key = (os.path.abspath(ref_audio), st.st_mtime_ns, st.st_size)
        if key in self._ref_cache:
            return self._ref_cache[key]
        if self._speech_encoder is None:
            from .encoders import SpeakerEncoder, SpeechEncoder
            if not any(n.startswith("enc.") for n in self._ref_weights) or not any(n.startswith("spk.") for n in self._ref_weights):
                raise ValueError("this checkpoint holds no speech-tokenizer encoder / speaker encoder: ref_audio needs the Base model")
            store = WeightStore(self.cfg)
            store.fp = self._ref_weights
            self._speech_encoder = SpeechEncoder(self.cfg, store, self.device)
            self._speaker_encoder = SpeakerEncoder(self.cfg, store, self.device)
        from .encoders import read_wav
        wav = read_wav(ref_audio, self.sample_rate).to(self.engine.dev)[None]
        if wav.shape[1] < self.cfg.spk.n_fft:
            raise ValueError(f"{ref_audio}: reference clip too short ({wav.shape[1]} samples)")
        codes = self._speech_encoder.encode(wav)[0].t().contiguous()            # [T_ref, 16]
        vec = self._speaker_encoder.embed(wav)[0]
        self._ref_cache = {key: (codes, vec)}
        return codes, vec


def write_wav(path: str, audio: np.ndarray, sample_rate: int) -> None:
    """Mono PCM16 WAV (what `save_audio_file` moves out of the temp dir, reference io.py:156-160)."""
    pcm = np.clip(np.asarray(audio, dtype=np.float32), -1.0, 1.0)
    pcm = np.round(pcm * 32767.0).astype("<i2")
    with wave.open(path, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(sample_rate)
        w.writeframes(pcm.tobytes())


_MODE_BY_FOLDER = {"customvoice": "custom_voice", "voicedesign": "voice_design", "base": "base"}


def load_model(model_path: str, device: str = "cuda", **kw) -> Model:
    """Drop-in for `mlx_audio.tts.utils.load_model(model_path)` (reference io.py:111-112).

    Reads `<model_path>/config.json` (+ `speech_tokenizer/config.json`): the checkpoint's own `talker_config` /
    `code_predictor_config` / `spk_id` / `codec_language_id` / token ids / `quantization` drive every architecture constant
    (config.from_hf_config).  `model.safetensors` (the `mlx-community/*-8bit` layout the reference downloads, io.py:42-52) is
    read by `mlx_loader.load_mlx_checkpoint` - the affine 8-bit codes are taken as they are.  Anything that would make the
    reference's "loaded" message a lie raises instead: a folder without weights (incomplete download) -> OSError, a
    malformed checkpoint or a missing tokenizer next to real weights -> ValueError; the reference reports both as "Failed
    to load model" (io.py:115-117).  Seeded random-init weights (the offline parity / benchmark set-up) are an explicit
    opt-in: `random_init_seed` or `b200_size` in config.json, or Q3T_ALLOW_RANDOM_INIT=1."""
    if not os.path.isdir(model_path):
        raise OSError(f"model directory not found: {model_path}")
    has_ckpt = os.path.exists(os.path.join(model_path, "model.safetensors")) or \
        os.path.exists(os.path.join(model_path, "model.safetensors.index.json"))

    def read_json(*parts):
        fn = os.path.join(model_path, *parts)
        if not os.path.exists(fn):
            return {}
        with open(fn) as f:
            return json.load(f)

    meta, speech_meta = read_json("config.json"), read_json("speech_tokenizer", "config.json")
    folder = os.path.basename(os.path.normpath(model_path)).lower().replace("-", "")
    mode = meta.get("tts_model_type") or next((v for k, v in _MODE_BY_FOLDER.items() if k in folder), "custom_voice")
    if "b200_config" in meta:                          # folders written by mlx_loader.export_mlx_checkpoint (fixtures)
        cfg = ModelConfig.from_dict(meta["b200_config"])
    elif "talker_config" in meta:                      # the checkpoint's own HF-style config tree
        cfg = cfgmod.from_hf_config(meta, speech_meta)
    else:
        cfg = getattr(cfgmod, meta.get("b200_size", "full"))(mode)
    cfg.tts_model_type = mode
    if has_ckpt:
        from .mlx_loader import load_mlx_checkpoint
        ws = load_mlx_checkpoint(model_path, cfg, device=device)
    else:
        if not ("random_init_seed" in meta or "b200_size" in meta or os.environ.get("Q3T_ALLOW_RANDOM_INIT") == "1"):
            raise OSError(f"{model_path} holds no model.safetensors (incomplete download?); random-init weights need an explicit "
                          "opt-in (random_init_seed / b200_size in config.json or Q3T_ALLOW_RANDOM_INIT=1)")
        parts = ("talker", "cp", "codec") + (("enc", "spk") if mode == "base" else ())
        ws = make_weights(cfg, seed=int(meta.get("random_init_seed", 0)), device=device, keep_fp=False, parts=parts)
    explicit_stub = bool(meta.get("b200_byte_tokenizer"))
    return Model(cfg, ws, device, model_path=model_path, require_tokenizer=has_ckpt and not explicit_stub, **kw)
