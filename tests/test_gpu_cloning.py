"""Voice cloning path (reference call site sessions/clone.py:218-224: generate_audio(model, text, ref_audio=<wav>, ref_text=)):
speech-tokenizer ENCODER, ECAPA speaker encoder (8f-3) and the in-context prompt layout (8a a3, App. C) on the
GPU against the CPU oracle (oracle/qwen3_tts_encoders_oracle.py, pinned on the transformers cousins)."""
import wave

import numpy as np
import pytest
import torch

from oracle import qwen3_tts_encoders_oracle as E
from oracle import qwen3_tts_oracle as O
from qwen3_tts_b200 import config as Cfg
from qwen3_tts_b200.weights import make_weights

pytestmark = pytest.mark.gpu
ALL = ("talker", "cp", "codec", "enc", "spk")


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


def _clip(n, seed, b=1):
    """Band-limited noise + a few tones, peak ~0.5: something with structure in every mel band."""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(n) / 24000.0
    x = 0.15 * torch.randn(b, n, generator=g)
    for f in (180.0, 440.0, 1250.0, 3100.0):
        x = x + 0.1 * torch.sin(2 * torch.pi * f * t + float(torch.rand((), generator=g)) * 6.28)[None]
    return x.clamp(-1, 1)


def _codes_match(codes_d, codes_o, rec, tol):
    """Integer parity of the RVQ encode: every index equal, except where the oracle's own nearest / second-nearest codebook
    distances differ by less than `tol` (relative to the distance) - a different but equally valid pick; after such a level the
    residuals of that frame differ, so later levels of the same frame are not compared."""
    B, Q, T = codes_o.shape
    n_tie = 0
    for b in range(B):
        for t in range(T):
            for q in range(Q):
                if int(codes_d[b, q, t]) == int(codes_o[b, q, t]):
                    continue
                gap, d0 = float(rec["gap"][q][b, t]), float(rec["dist0"][q][b, t])
                assert gap <= tol * max(d0, 1e-6), f"b={b} frame {t} level {q}: {int(codes_d[b, q, t])} vs {int(codes_o[b, q, t])}, gap {gap:.3e} of {d0:.3e}"
                n_tie += 1
                break
    return n_tie


@pytest.mark.parametrize("size,n", [("small", 31111), ("full", 36000)])
def test_speech_encoder_against_the_cpu_oracle(cuda, size, n):
    from qwen3_tts_b200.encoders import SpeechEncoder
    cfg = getattr(Cfg, size)("base")
    ws = make_weights(cfg, seed=7, parts=("enc",))
    enc = SpeechEncoder(cfg, ws, "cuda")
    wav = _clip(n, 3, b=2)
    rec, sd = {}, {}
    with torch.no_grad():
        codes_o = E.speech_encode(ws.fp, cfg.enc, wav, rec)
    emb_d = enc.embeddings(wav.cuda(), sd)
    for name in ("stage0", "stage3", "seanet", "transformer"):
        assert _rel(sd[name].cpu().transpose(1, 2), rec[name]) < 2e-5, name
    assert _rel(emb_d.cpu().transpose(1, 2), rec["embeddings"]) < 2e-5
    codes_d = enc.quantize(emb_d).cpu().long()
    assert codes_d.shape == codes_o.shape == (2, cfg.enc.valid_quantizers, -(-n // cfg.enc.hop))
    n_tie = _codes_match(codes_d, codes_o, rec, 1e-4)
    # the quantiser alone on the ORACLE's embeddings: integer work on identical inputs
    codes_t = enc.quantize(rec["embeddings"].transpose(1, 2).contiguous().cuda()).cpu().long()
    n_tie_t = _codes_match(codes_t, codes_o, rec, 2e-6)
    frames = codes_o.shape[0] * codes_o.shape[2]
    assert n_tie <= max(1, frames // 10) and n_tie_t <= 1, (n_tie, n_tie_t, frames)


@pytest.mark.parametrize("size", ["small", "full"])
def test_speaker_encoder_against_the_cpu_oracle(cuda, size):
    from qwen3_tts_b200.encoders import SpeakerEncoder, _mel_filter_bank
    cfg = getattr(Cfg, size)("base")
    sc = cfg.spk
    ws = make_weights(cfg, seed=8, parts=("spk",))
    spk = SpeakerEncoder(cfg, ws, "cuda")
    assert _rel(_mel_filter_bank(sc.n_fft, sc.n_mels, sc.sample_rate, sc.fmin, sc.fmax), E.mel_filter_bank(sc.n_fft, sc.n_mels, sc.sample_rate, sc.fmin, sc.fmax)) < 1e-6
    wav = _clip(24000 * 2 + 123, 5, b=2)
    so, sd = {}, {}
    with torch.no_grad():
        mel_o = E.log_mel(wav, sc)
        vec_o = E.ecapa_forward(ws.fp, sc, mel_o, so)
    mel_d = spk.log_mel(wav.cuda())
    assert mel_d.shape == mel_o.shape
    assert float((mel_d.cpu() - mel_o).abs().max()) < 2e-3          # log domain; fp32 DFT as a GEMM vs the FFT
    vec_d = spk.ecapa(mel_o.cuda(), sd)                              # the network alone on identical mel input
    for name in so:
        assert _rel(sd[name].cpu().transpose(1, 2), so[name]) < 5e-5, name
    assert vec_d.shape == vec_o.shape == (2, sc.enc_dim) and _rel(vec_d.cpu(), vec_o) < 5e-5
    assert _rel(spk.embed(wav.cuda()).cpu(), vec_o) < 2e-3          # end to end, through the device mel


@pytest.fixture(scope="module")
def base_setup(cuda):
    from qwen3_tts_b200.model import Model
    cfg = Cfg.small("base")
    ws = make_weights(cfg, seed=0, head_std=0.2, parts=ALL)
    model = Model(cfg, ws, "cuda", max_frames=64, max_ctx=512, max_trailing=128, prefill="decode")
    oracle = O.OracleModel(cfg, ws.fp, kv_dtype=torch.bfloat16)
    return cfg, ws, model, oracle


def _ids(cfg, n, seed):
    g = torch.Generator().manual_seed(seed)
    body = torch.randint(0, cfg.talker.text_vocab_size - 16, (n,), generator=g).tolist()
    return [cfg.im_start_id, cfg.assistant_id, 10] + body + [cfg.im_end_id, 10, cfg.im_start_id, cfg.assistant_id, 10]


def test_icl_prefill_layout_and_frames_against_the_oracle(base_setup):
    """BASELINE config 3 mechanics: reference codes + reference text in the prompt.  Embeddings of all four layouts (streaming /
    non-streaming x text longer / shorter than the clip) equal the oracle's; then teacher-forced frames from the streaming ICL
    prompt: logits within 1e-2 relative, same argmax (trailing text rows feed the first frames)."""
    cfg, ws, model, oracle = base_setup
    g = torch.Generator().manual_seed(11)
    vec = torch.randn(cfg.talker.hidden_size, generator=g) * 0.02
    ref_ids = [cfg.im_start_id, cfg.assistant_id, 10] + [31, 32, 33, 34] + [cfg.im_end_id, 10]
    for n_ref, n_text in ((9, 20), (30, 6)):
        rc = torch.randint(0, cfg.codec.codebook_size, (n_ref, 16), generator=g)
        for streaming in (True, False):
            kw = dict(speaker_vec=vec, streaming=streaming, ref_codes=rc, ref_text_ids=ref_ids, language="english")
            pre_o, tr_o = oracle.build_prefill(_ids(cfg, n_text, n_ref), **kw)
            pre_d, tr_d = model.build_prefill(_ids(cfg, n_text, n_ref), **kw)
            assert pre_d.shape == pre_o.shape and tr_d.shape == tr_o.shape, (n_ref, n_text, streaming)
            assert _rel(pre_d.cpu(), pre_o) < 1e-5 and _rel(tr_d.cpu(), tr_o) < 1e-5
    rc = torch.randint(0, cfg.codec.codebook_size, (9, 16), generator=g)
    pre, tr = oracle.build_prefill(_ids(cfg, 20, 1), speaker_vec=vec, streaming=True, ref_codes=rc, ref_text_ids=ref_ids)
    assert tr.shape[0] > 5
    n = 10
    codes_o, rec = oracle.generate(pre, tr, n, record=True)
    e2 = type(model.engine)(cfg, ws, "cuda", batch=1, max_frames=64, max_ctx=256, keep_cp_logits=True, prefill="decode", max_trailing=64)
    e2.set_sampling(do_sample=False)
    e2.set_forced(codes_o[None])
    e2.prefill(pre[None], None, tr[None])
    for f in range(n):
        assert _rel(e2.logits[0].cpu(), rec["talker_logits"][f]) < 1e-2, f
        e2._run("frame")
        assert _rel(e2.cp_logits[:, 0].cpu(), rec["cp_logits"][f]) < 1e-2, f
    assert torch.equal(e2.own_codes[0, :n].cpu().long(), torch.tensor(rec["own_codes"]))


def test_generate_with_ref_audio_follows_the_oracle_pipeline(base_setup, tmp_path):
    """The reference-facing call of the clone session: a 24 kHz mono PCM16 wav on disk + its transcript -> encoder codes,
    speaker vector, ICL prompt, frames, codec over ref ++ new codes with the reference span cut off.  Every stage is compared
    with the oracle run on the same file."""
    cfg, ws, model, oracle = base_setup
    pcm = (np.round(_clip(24000 * 2, 21)[0].numpy() * 32767.0)).astype("<i2")
    path = str(tmp_path / "Boss.wav")
    with wave.open(path, "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(24000); w.writeframes(pcm.tobytes())
    wav = torch.from_numpy(pcm.astype(np.float32) / 32768.0)[None]
    rec = {}
    with torch.no_grad():
        codes_ref_o = E.speech_encode(ws.fp, cfg.enc, wav, rec)[0].t()               # [T_ref, 16]
        vec_o = E.speaker_embed(ws.fp, cfg.spk, wav)[0]
    codes_ref_d, vec_d = model._reference_prompt(path)
    assert codes_ref_d.shape == codes_ref_o.shape == (25, 16)
    _codes_match(codes_ref_d.t()[None].cpu().long(), codes_ref_o.t()[None], rec, 1e-4)
    assert _rel(vec_d.cpu(), vec_o) < 2e-3
    n = 6
    res = list(model.generate("clone me please", ref_audio=path, ref_text="This is what the boss says.", greedy=True, max_tokens=n))
    assert len(res) == 1 and res[0].token_count == n
    # the oracle on the DEVICE's own reference codes / vector (integer ties aside, they are the oracle's): same prompt -> same frames
    text_ids = model.chat_ids("clone me please")
    ref_ids = model.ref_text_chat_ids("This is what the boss says.")
    pre, tr = oracle.build_prefill(text_ids, speaker_vec=vec_d.cpu(), streaming=True, ref_codes=codes_ref_d.cpu(), ref_text_ids=ref_ids)
    codes_o, orec = oracle.generate(pre, tr, n, record=True)
    got = torch.from_numpy(res[0].codes).long()
    diff = got != codes_o
    if diff.any():
        f = int(diff.any(1).nonzero()[0]); gq = int(diff[f].nonzero()[0])
        lg = orec["talker_logits"][f] if gq == 0 else orec["cp_logits"][f][gq - 1]
        gap = float(lg[int(codes_o[f, gq])] - lg[int(got[f, gq])])
        assert 0 <= gap <= 3e-3 * float(lg.abs().max()), f"frame {f} group {gq}: gap {gap:.3e}"
        got = codes_o
    # audio = codec(ref ++ new) with the reference span cut proportionally
    allc = torch.cat([codes_ref_d.cpu().long(), got], 0)
    wav_o = O.codec_chunked_decode(ws.fp, cfg, allc.t()[None])[0, 0]
    wav_o = wav_o[int(25 / (25 + n) * wav_o.shape[0]):]
    if not diff.any():
        a = torch.from_numpy(res[0].audio)
        assert a.shape == wav_o.shape
        snr = 10 * torch.log10(wav_o.double().pow(2).sum() / (a.double() - wav_o.double()).pow(2).sum().clamp_min(1e-30))
        assert float(snr) >= 40.0
    # streaming pieces concatenate to the same number of samples as frames were generated, and a second call hits the cache
    pieces = list(model.generate("clone me please", ref_audio=path, ref_text=".", greedy=True, max_tokens=n, stream=True, streaming_interval=0.24))
    assert sum(p.token_count for p in pieces) == n and all(0 < p.samples <= p.token_count * cfg.codec.hop for p in pieces)
    assert len(model._ref_cache) == 1
    # a model without the encoders refuses instead of inventing a voice
    from qwen3_tts_b200.model import Model
    m2 = Model(cfg, make_weights(cfg, seed=0, parts=("talker", "cp", "codec")), "cuda", max_frames=8, max_ctx=128)
    with pytest.raises(ValueError, match="encoder"):
        list(m2.generate("x", ref_audio=path, greedy=True, max_tokens=2))
