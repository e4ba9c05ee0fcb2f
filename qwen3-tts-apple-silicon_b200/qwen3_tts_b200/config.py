"""Architecture constants of the Qwen3-TTS-12Hz generation hot path.

Every value mirrors the checkpoint's ``config.json``.  The
reference never hard-codes these: it reads them through the un-vendored
``mlx_audio.tts.utils.load_model`` (reference call site: src/qwen3_tts/io.py:111-112),
so nothing here is a literal inside a kernel -- kernels receive them as arguments.

Three sizes are defined:
  * ``full()``  -- the 1.7B shape named by BASELINE.json (bench + full-size parity)
  * ``small()`` -- same head_dim / group structure, fewer and narrower layers (GPU parity)
  * ``tiny()``  -- seconds on CPU (oracle-vs-cousin tests, golden fixtures)
"""
from __future__ import annotations

from dataclasses import dataclass, field, asdict
from typing import Dict, Tuple

QUANT_GROUP = 64  # config.json "quantization": {"group_size": 64, "bits": 8}
QUANT_BITS = 8


@dataclass
class TalkerConfig:
    hidden_size: int = 2048
    num_layers: int = 28
    num_heads: int = 16
    num_kv_heads: int = 8
    head_dim: int = 128
    intermediate_size: int = 6144
    vocab_size: int = 3072          # codec vocabulary of code group 0 (+ control ids)
    text_vocab_size: int = 151936
    text_hidden_size: int = 2048
    rms_norm_eps: float = 1e-6
    rope_theta: float = 1e6
    # control ids inside the codec vocabulary
    codec_pad_id: int = 2148
    codec_bos_id: int = 2149
    codec_eos_id: int = 2150
    codec_think_id: int = 2154
    codec_nothink_id: int = 2155
    codec_think_bos_id: int = 2156
    codec_think_eos_id: int = 2157
    codec_language_id: Dict[str, int] = field(default_factory=lambda: {
        "chinese": 2055, "english": 2050, "german": 2053, "italian": 2070, "portuguese": 2071,
        "spanish": 2054, "japanese": 2058, "korean": 2064, "french": 2061, "russian": 2069})
    # the nine preset speakers of reference config.py:44-49 (lower-cased at custom.py:166)
    spk_id: Dict[str, int] = field(default_factory=lambda: {
        "serena": 3066, "vivian": 3065, "uncle_fu": 3010, "ryan": 3061, "aiden": 2861,
        "ono_anna": 2873, "sohee": 2864, "eric": 2875, "dylan": 2878})

    @property
    def q_dim(self) -> int:
        return self.num_heads * self.head_dim

    @property
    def kv_dim(self) -> int:
        return self.num_kv_heads * self.head_dim


@dataclass
class CodePredictorConfig:
    hidden_size: int = 1024
    num_layers: int = 5
    num_heads: int = 16
    num_kv_heads: int = 8
    head_dim: int = 128
    intermediate_size: int = 3072
    vocab_size: int = 2048
    num_code_groups: int = 16       # -> 15 embeddings + 15 heads
    embed_dim: int = 2048           # codec_embedding width == talker hidden
    rms_norm_eps: float = 1e-6
    rope_theta: float = 1e6

    @property
    def q_dim(self) -> int:
        return self.num_heads * self.head_dim

    @property
    def kv_dim(self) -> int:
        return self.num_kv_heads * self.head_dim


@dataclass
class CodecConfig:
    """Speech-tokenizer decoder (12.5 Hz -> 24 kHz)."""
    num_quantizers: int = 16
    num_semantic: int = 1
    codebook_size: int = 2048
    codebook_dim: int = 256         # vector width of every codebook
    rvq_out_dim: int = 512          # 1x1 output projections 256 -> 512
    latent_dim: int = 1024          # pre_conv output / ConvNeXt width
    tf_hidden: int = 512
    tf_intermediate: int = 1024
    tf_heads: int = 16
    tf_head_dim: int = 64
    tf_layers: int = 8
    sliding_window: int = 72
    tf_rope_theta: float = 1e4
    tf_rms_eps: float = 1e-5
    layer_scale: float = 0.01
    upsampling_ratios: Tuple[int, ...] = (2, 2)
    upsample_rates: Tuple[int, ...] = (8, 5, 4, 3)
    decoder_dim: int = 1536
    chunk_size: int = 300
    left_context: int = 25
    # the on-disk cousin trims k-s on BOTH sides of every transposed conv
    # (qwen3_omni_moe:3319-3331); "right" is the strictly causal alternative.
    transconv_trim: str = "both"
    sample_rate: int = 24000

    @property
    def hop(self) -> int:
        h = 1
        for r in self.upsampling_ratios + self.upsample_rates:
            h *= r
        return h

    def out_len(self, frames: int) -> int:
        """Samples produced by ONE vocoder call on `frames` code frames."""
        n = frames
        for r in self.upsampling_ratios:
            n *= r
        for r in self.upsample_rates:
            n = (n - 1) * r if self.transconv_trim == "both" else n * r
        return n


@dataclass
class EncoderConfig:
    """Speech-tokenizer ENCODER (24 kHz wav -> 12.5 Hz RVQ codes; voice cloning only).  Qwen3-TTS-Tokenizer-12Hz
    encodes with a Mimi model; defaults = transformers MimiConfig."""
    hidden_size: int = 512
    num_filters: int = 64
    ratios: Tuple[int, ...] = (8, 6, 5, 4)          # applied in REVERSE order by the encoder (mimi:466)
    kernel_size: int = 7
    last_kernel_size: int = 3
    residual_kernel_size: int = 3
    compress: int = 2
    tf_layers: int = 8
    tf_heads: int = 8
    tf_head_dim: int = 64
    tf_intermediate: int = 2048
    sliding_window: int = 250
    rope_theta: float = 1e4
    norm_eps: float = 1e-5
    layer_scale: float = 0.01
    codebook_size: int = 2048
    codebook_dim: int = 256
    num_quantizers: int = 32                        # codebooks in the checkpoint
    num_semantic: int = 1
    valid_quantizers: int = 16                      # how many the TTS model consumes (encoder_valid_num_quantizers)

    @property
    def hop(self) -> int:
        h = 2                                       # stride-2 downsample after the transformer (25 Hz -> 12.5 Hz)
        for r in self.ratios:
            h *= r
        return h


@dataclass
class SpeakerEncoderConfig:
    """Speaker encoder of the Base model: log-mel front end + ECAPA-TDNN -> one vector of the talker width."""
    sample_rate: int = 24000
    n_fft: int = 1024
    hop: int = 256
    win: int = 1024
    n_mels: int = 128
    fmin: float = 0.0
    fmax: float = 12000.0
    channels: Tuple[int, ...] = (512, 512, 512, 512, 1536)
    kernel_sizes: Tuple[int, ...] = (5, 3, 3, 3, 1)
    dilations: Tuple[int, ...] = (1, 2, 3, 4, 1)
    attention_channels: int = 128
    res2net_scale: int = 8
    se_channels: int = 128
    enc_dim: int = 2048


@dataclass
class ModelConfig:
    tts_model_type: str = "custom_voice"      # custom_voice | voice_design | base
    talker: TalkerConfig = field(default_factory=TalkerConfig)
    cp: CodePredictorConfig = field(default_factory=CodePredictorConfig)
    codec: CodecConfig = field(default_factory=CodecConfig)
    enc: EncoderConfig = field(default_factory=EncoderConfig)
    spk: SpeakerEncoderConfig = field(default_factory=SpeakerEncoderConfig)
    tts_pad_token_id: int = 151671
    tts_bos_token_id: int = 151672
    tts_eos_token_id: int = 151673
    im_start_id: int = 151644
    im_end_id: int = 151645
    assistant_id: int = 77091
    quant_group: int = QUANT_GROUP
    quant_bits: int = QUANT_BITS

    def to_dict(self) -> dict:
        return asdict(self)

    @classmethod
    def from_dict(cls, d: dict) -> "ModelConfig":
        d = dict(d)
        t = TalkerConfig(**d.pop("talker", {}))
        c = CodePredictorConfig(**d.pop("cp", {}))
        k = dict(d.pop("codec", {}))
        for key in ("upsampling_ratios", "upsample_rates"):
            if key in k:
                k[key] = tuple(k[key])
        en = {kk: (tuple(v) if isinstance(v, list) else v) for kk, v in d.pop("enc", {}).items()}
        sp = {kk: (tuple(v) if isinstance(v, list) else v) for kk, v in d.pop("spk", {}).items()}
        return cls(talker=t, cp=c, codec=CodecConfig(**k), enc=EncoderConfig(**en), spk=SpeakerEncoderConfig(**sp), **d)


# ---------------------------------------------------------------------------------------------------------------------
# the checkpoint's own config.json (what mlx_audio.tts.utils.load_model reads; reference io.py:111-112, folders of
# config.py:17,26,35).  Layout per SURVEY Appendix A: top-level token ids + `talker_config` (with a nested
# `code_predictor_config`, `spk_id`, `codec_language_id`, codec control ids) + `quantization`, and
# `speech_tokenizer/config.json` with a `decoder_config`.  Every [U] constant of the survey is read from the file; the
# dataclass defaults above only fill keys a checkpoint does not carry.  Several spellings are accepted per key because the
# exact upstream names cannot be checked offline.
# ---------------------------------------------------------------------------------------------------------------------
def _pick(d: dict, names, default):
    for n in names:
        if n in d and d[n] is not None:
            return d[n]
    return default


def _rope_theta(d: dict, default: float) -> float:
    rp = d.get("rope_parameters") or {}
    return float(_pick(d, ("rope_theta",), rp.get("rope_theta", default)))


def from_hf_config(meta: dict, speech_meta: dict = None) -> ModelConfig:
    """`config.json` (+ `speech_tokenizer/config.json`) of a Qwen3-TTS folder -> ModelConfig."""
    cfg = ModelConfig()
    cfg.tts_model_type = meta.get("tts_model_type", cfg.tts_model_type)
    for ours, theirs in (("tts_pad_token_id", ("tts_pad_token_id",)), ("tts_bos_token_id", ("tts_bos_token_id",)),
                         ("tts_eos_token_id", ("tts_eos_token_id",)), ("im_start_id", ("im_start_token_id", "im_start_id")),
                         ("im_end_id", ("im_end_token_id", "im_end_id")), ("assistant_id", ("assistant_token_id", "assistant_id"))):
        setattr(cfg, ours, int(_pick(meta, theirs, getattr(cfg, ours))))
    q = meta.get("quantization") or meta.get("quantization_config") or {}
    cfg.quant_group, cfg.quant_bits = int(q.get("group_size", cfg.quant_group)), int(q.get("bits", cfg.quant_bits))
    if cfg.quant_bits != 8 or cfg.quant_group != 64:
        raise ValueError(f"unsupported quantization {q}: this build streams affine 8-bit, group 64 (the *-8bit checkpoints)")
    tc = meta.get("talker_config") or {}
    t = cfg.talker
    for ours, theirs in (("hidden_size", ("hidden_size",)), ("num_layers", ("num_hidden_layers",)), ("num_heads", ("num_attention_heads",)),
                         ("num_kv_heads", ("num_key_value_heads",)), ("intermediate_size", ("intermediate_size",)),
                         ("vocab_size", ("vocab_size",)), ("text_vocab_size", ("text_vocab_size",)), ("text_hidden_size", ("text_hidden_size",)),
                         ("codec_pad_id", ("codec_pad_id", "codec_pad_token_id")), ("codec_bos_id", ("codec_bos_id", "codec_bos_token_id")),
                         ("codec_eos_id", ("codec_eos_token_id", "codec_eos_id")), ("codec_think_id", ("codec_think_id",)),
                         ("codec_nothink_id", ("codec_nothink_id",)), ("codec_think_bos_id", ("codec_think_bos_id",)),
                         ("codec_think_eos_id", ("codec_think_eos_id",))):
        setattr(t, ours, int(_pick(tc, theirs, getattr(t, ours))))
    t.head_dim = int(_pick(tc, ("head_dim",), t.hidden_size // t.num_heads if "hidden_size" in tc and "head_dim" not in tc else t.head_dim))
    t.rms_norm_eps = float(_pick(tc, ("rms_norm_eps",), t.rms_norm_eps))
    t.rope_theta = _rope_theta(tc, t.rope_theta)
    if tc.get("codec_language_id"):
        t.codec_language_id = {str(k).lower(): int(v) for k, v in tc["codec_language_id"].items()}
    if tc.get("spk_id"):
        # values may be plain ids or one-element lists (the speaker slot of the prefill takes one codec id)
        t.spk_id = {str(k).lower(): int(v[0] if isinstance(v, (list, tuple)) else v) for k, v in tc["spk_id"].items()}
    pc = tc.get("code_predictor_config") or meta.get("code_predictor_config") or {}
    c = cfg.cp
    for ours, theirs in (("hidden_size", ("hidden_size",)), ("num_layers", ("num_hidden_layers",)), ("num_heads", ("num_attention_heads",)),
                         ("num_kv_heads", ("num_key_value_heads",)), ("intermediate_size", ("intermediate_size",)),
                         ("vocab_size", ("vocab_size",)), ("num_code_groups", ("num_code_groups",))):
        setattr(c, ours, int(_pick(pc, theirs, getattr(c, ours))))
    c.num_code_groups = int(_pick(tc, ("num_code_groups",), c.num_code_groups)) if "num_code_groups" not in pc else c.num_code_groups
    c.head_dim = int(_pick(pc, ("head_dim",), c.head_dim))
    c.rms_norm_eps = float(_pick(pc, ("rms_norm_eps",), c.rms_norm_eps))
    c.rope_theta = _rope_theta(pc, c.rope_theta)
    c.embed_dim = int(_pick(pc, ("codec_embedding_dim", "embed_dim"), t.hidden_size))
    dc = (speech_meta or {}).get("decoder_config") or {}
    k = cfg.codec
    for ours, theirs in (("num_quantizers", ("num_quantizers",)), ("num_semantic", ("num_semantic_quantizers",)),
                         ("codebook_size", ("codebook_size",)), ("latent_dim", ("latent_dim",)), ("tf_hidden", ("hidden_size",)),
                         ("tf_intermediate", ("intermediate_size",)), ("tf_heads", ("num_attention_heads",)), ("tf_head_dim", ("head_dim",)),
                         ("tf_layers", ("num_hidden_layers",)), ("sliding_window", ("sliding_window",)), ("decoder_dim", ("decoder_dim",))):
        setattr(k, ours, int(_pick(dc, theirs, getattr(k, ours))))
    if "codebook_dim" in dc:                     # upstream names the concatenated width (512); each codebook vector is half of it
        k.rvq_out_dim = int(dc["codebook_dim"])
        k.codebook_dim = int(_pick(dc, ("vector_quantization_hidden_dimension",), k.rvq_out_dim // 2))
    k.tf_rope_theta = _rope_theta(dc, k.tf_rope_theta)
    k.tf_rms_eps = float(_pick(dc, ("rms_norm_eps",), k.tf_rms_eps))
    k.layer_scale = float(_pick(dc, ("layer_scale_initial_scale",), k.layer_scale))
    for key in ("upsampling_ratios", "upsample_rates"):
        if key in dc:
            setattr(k, key, tuple(int(v) for v in dc[key]))
    k.sample_rate = int(_pick(speech_meta or {}, ("output_sample_rate", "sample_rate"), k.sample_rate))
    ec = (speech_meta or {}).get("encoder_config") or {}                   # a MimiConfig dict
    en = cfg.enc
    for ours, theirs in (("hidden_size", ("hidden_size",)), ("num_filters", ("num_filters",)), ("kernel_size", ("kernel_size",)),
                         ("last_kernel_size", ("last_kernel_size",)), ("residual_kernel_size", ("residual_kernel_size",)),
                         ("compress", ("compress",)), ("tf_layers", ("num_hidden_layers",)), ("tf_heads", ("num_attention_heads",)),
                         ("tf_head_dim", ("head_dim",)), ("tf_intermediate", ("intermediate_size",)), ("sliding_window", ("sliding_window",)),
                         ("codebook_size", ("codebook_size",)), ("codebook_dim", ("vector_quantization_hidden_dimension", "codebook_dim")),
                         ("num_quantizers", ("num_quantizers",)), ("num_semantic", ("num_semantic_quantizers",))):
        setattr(en, ours, int(_pick(ec, theirs, getattr(en, ours))))
    if "upsampling_ratios" in ec:
        en.ratios = tuple(int(v) for v in ec["upsampling_ratios"])
    en.rope_theta = _rope_theta(ec, en.rope_theta)
    en.norm_eps = float(_pick(ec, ("norm_eps",), en.norm_eps))
    en.layer_scale = float(_pick(ec, ("layer_scale_initial_scale",), en.layer_scale))
    en.valid_quantizers = int(_pick(speech_meta or {}, ("encoder_valid_num_quantizers",), min(en.valid_quantizers, en.num_quantizers)))
    sc = meta.get("speaker_encoder_config") or {}
    sp = cfg.spk
    for ours, theirs in (("sample_rate", ("sample_rate",)), ("n_fft", ("n_fft",)), ("hop", ("hop_size", "hop_length")), ("win", ("win_size", "win_length")),
                         ("n_mels", ("mel_dim", "num_mels")), ("attention_channels", ("enc_attention_channels",)),
                         ("res2net_scale", ("enc_res2net_scale",)), ("se_channels", ("enc_se_channels",))):
        setattr(sp, ours, int(_pick(sc, theirs, getattr(sp, ours))))
    for ours, theirs in (("channels", "enc_channels"), ("kernel_sizes", "enc_kernel_sizes"), ("dilations", "enc_dilations")):
        if theirs in sc:
            setattr(sp, ours, tuple(int(v) for v in sc[theirs]))
    sp.fmin, sp.fmax = float(_pick(sc, ("fmin",), sp.fmin)), float(_pick(sc, ("fmax",), sp.fmax))
    sp.enc_dim = int(_pick(sc, ("enc_dim",), t.hidden_size))
    return cfg


def to_hf_config(cfg: ModelConfig):
    """Inverse of `from_hf_config` (fixtures: a folder in the checkpoint's own layout).  Returns (config.json dict,
    speech_tokenizer/config.json dict)."""
    t, c, k = cfg.talker, cfg.cp, cfg.codec
    meta = {
        "model_type": "qwen3_tts", "tts_model_type": cfg.tts_model_type, "tokenizer_type": "qwen3_tts_tokenizer_12hz",
        "tts_pad_token_id": cfg.tts_pad_token_id, "tts_bos_token_id": cfg.tts_bos_token_id, "tts_eos_token_id": cfg.tts_eos_token_id,
        "im_start_token_id": cfg.im_start_id, "im_end_token_id": cfg.im_end_id, "assistant_token_id": cfg.assistant_id,
        "quantization": {"group_size": cfg.quant_group, "bits": cfg.quant_bits},
        "talker_config": {
            "hidden_size": t.hidden_size, "num_hidden_layers": t.num_layers, "num_attention_heads": t.num_heads,
            "num_key_value_heads": t.num_kv_heads, "head_dim": t.head_dim, "intermediate_size": t.intermediate_size,
            "vocab_size": t.vocab_size, "text_vocab_size": t.text_vocab_size, "text_hidden_size": t.text_hidden_size,
            "rms_norm_eps": t.rms_norm_eps, "rope_theta": t.rope_theta, "num_code_groups": c.num_code_groups,
            "codec_pad_id": t.codec_pad_id, "codec_bos_id": t.codec_bos_id, "codec_eos_token_id": t.codec_eos_id,
            "codec_think_id": t.codec_think_id, "codec_nothink_id": t.codec_nothink_id, "codec_think_bos_id": t.codec_think_bos_id,
            "codec_think_eos_id": t.codec_think_eos_id, "codec_language_id": dict(t.codec_language_id), "spk_id": dict(t.spk_id),
            "code_predictor_config": {
                "hidden_size": c.hidden_size, "num_hidden_layers": c.num_layers, "num_attention_heads": c.num_heads,
                "num_key_value_heads": c.num_kv_heads, "head_dim": c.head_dim, "intermediate_size": c.intermediate_size,
                "vocab_size": c.vocab_size, "num_code_groups": c.num_code_groups, "rms_norm_eps": c.rms_norm_eps,
                "rope_theta": c.rope_theta}}}
    en, sp = cfg.enc, cfg.spk
    meta["speaker_encoder_config"] = {
        "sample_rate": sp.sample_rate, "n_fft": sp.n_fft, "hop_size": sp.hop, "win_size": sp.win, "mel_dim": sp.n_mels, "fmin": sp.fmin,
        "fmax": sp.fmax, "enc_channels": list(sp.channels), "enc_kernel_sizes": list(sp.kernel_sizes), "enc_dilations": list(sp.dilations),
        "enc_attention_channels": sp.attention_channels, "enc_res2net_scale": sp.res2net_scale, "enc_se_channels": sp.se_channels,
        "enc_dim": sp.enc_dim}
    speech = {"output_sample_rate": k.sample_rate, "encoder_valid_num_quantizers": en.valid_quantizers, "encoder_config": {
        "hidden_size": en.hidden_size, "num_filters": en.num_filters, "upsampling_ratios": list(en.ratios), "kernel_size": en.kernel_size,
        "last_kernel_size": en.last_kernel_size, "residual_kernel_size": en.residual_kernel_size, "compress": en.compress,
        "num_hidden_layers": en.tf_layers, "num_attention_heads": en.tf_heads, "head_dim": en.tf_head_dim,
        "intermediate_size": en.tf_intermediate, "sliding_window": en.sliding_window, "rope_theta": en.rope_theta,
        "norm_eps": en.norm_eps, "layer_scale_initial_scale": en.layer_scale, "codebook_size": en.codebook_size,
        "vector_quantization_hidden_dimension": en.codebook_dim, "num_quantizers": en.num_quantizers,
        "num_semantic_quantizers": en.num_semantic}, "decoder_config": {
        "num_quantizers": k.num_quantizers, "num_semantic_quantizers": k.num_semantic, "codebook_size": k.codebook_size,
        "codebook_dim": k.rvq_out_dim, "vector_quantization_hidden_dimension": k.codebook_dim, "latent_dim": k.latent_dim,
        "hidden_size": k.tf_hidden, "intermediate_size": k.tf_intermediate, "num_attention_heads": k.tf_heads, "head_dim": k.tf_head_dim,
        "num_hidden_layers": k.tf_layers, "sliding_window": k.sliding_window, "rope_theta": k.tf_rope_theta,
        "rms_norm_eps": k.tf_rms_eps, "layer_scale_initial_scale": k.layer_scale, "upsampling_ratios": list(k.upsampling_ratios),
        "upsample_rates": list(k.upsample_rates), "decoder_dim": k.decoder_dim}}
    return meta, speech


def full(tts_model_type: str = "custom_voice") -> ModelConfig:
    return ModelConfig(tts_model_type=tts_model_type)


def _with_small_text_vocab(cfg: ModelConfig) -> ModelConfig:
    """Reduced text vocabularies keep the special text ids at the top of the table."""
    v = cfg.talker.text_vocab_size
    cfg.tts_pad_token_id, cfg.tts_bos_token_id, cfg.tts_eos_token_id = v - 3, v - 2, v - 1
    cfg.im_start_id, cfg.im_end_id, cfg.assistant_id = v - 6, v - 5, v - 4
    return cfg


def small(tts_model_type: str = "custom_voice") -> ModelConfig:
    """GPU-parity size: real head_dim/GQA/group structure, 3+2 layers, narrow."""
    t = TalkerConfig(hidden_size=512, num_layers=3, num_heads=4, num_kv_heads=2, head_dim=128,
                     intermediate_size=1024, text_vocab_size=4096, text_hidden_size=512)
    c = CodePredictorConfig(hidden_size=256, num_layers=2, num_heads=4, num_kv_heads=2, head_dim=128,
                            intermediate_size=512, embed_dim=512)
    k = CodecConfig(codebook_dim=64, rvq_out_dim=128, latent_dim=256, tf_hidden=128, tf_intermediate=256,
                    tf_heads=4, tf_head_dim=32, tf_layers=2, decoder_dim=192)
    en = EncoderConfig(hidden_size=128, num_filters=16, tf_layers=2, tf_heads=4, tf_head_dim=32, tf_intermediate=256,
                       sliding_window=20, codebook_dim=64, num_quantizers=18)
    sp = SpeakerEncoderConfig(n_mels=32, channels=(64, 64, 64, 128), kernel_sizes=(5, 3, 3, 1), dilations=(1, 2, 3, 1),
                              attention_channels=32, res2net_scale=4, se_channels=32, enc_dim=t.hidden_size)
    return _with_small_text_vocab(ModelConfig(tts_model_type=tts_model_type, talker=t, cp=c, codec=k, enc=en, spk=sp))


def tiny(tts_model_type: str = "custom_voice") -> ModelConfig:
    """CPU-seconds size used by the golden fixtures."""
    t = TalkerConfig(hidden_size=128, num_layers=2, num_heads=4, num_kv_heads=2, head_dim=32,
                     intermediate_size=256, text_vocab_size=1024, text_hidden_size=128)
    c = CodePredictorConfig(hidden_size=64, num_layers=2, num_heads=4, num_kv_heads=2, head_dim=32,
                            intermediate_size=128, embed_dim=128)
    k = CodecConfig(codebook_dim=32, rvq_out_dim=64, latent_dim=64, tf_hidden=64, tf_intermediate=128,
                    tf_heads=4, tf_head_dim=16, tf_layers=2, decoder_dim=96)
    en = EncoderConfig(hidden_size=64, num_filters=8, tf_layers=2, tf_heads=4, tf_head_dim=16, tf_intermediate=128,
                       sliding_window=12, codebook_dim=32, num_quantizers=17)
    sp = SpeakerEncoderConfig(n_fft=256, hop=64, win=256, n_mels=16, channels=(32, 32, 32, 64), kernel_sizes=(5, 3, 3, 1), dilations=(1, 2, 3, 1),
                              attention_channels=16, res2net_scale=4, se_channels=16, enc_dim=t.hidden_size)
    return _with_small_text_vocab(ModelConfig(tts_model_type=tts_model_type, talker=t, cp=c, codec=k, enc=en, spk=sp))
