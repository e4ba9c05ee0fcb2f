"""MLX-quantised safetensors -> WeightStore (the first "next" row after the hot path).

The reference resolves `models/<folder>/` (reference `src/qwen3_tts/io.py:42-52`) and hands the path to
`mlx_audio.tts.utils.load_model` (`io.py:111-112`); the folders it names are `mlx-community/*-8bit` conversions
(`config.py:17,26,35`): `config.json`, `model.safetensors` (talker + code predictor, affine 8-bit, group 64) and
`speech_tokenizer/model.safetensors` (codec, float).

What is certain and what is not (no checkpoint is reachable offline):
  * the QUANTISATION format is MLX's published affine scheme: `<name>.weight` uint32 [out, in/4] with four
    codes per word, little-endian along the input axis, `<name>.scales` / `<name>.biases` [out, in/64].  It is bit-identical
    to this repo's W8 triple, so codes, scales and biases are taken as they are - nothing is re-quantised.
  * the KEY NAMES follow the HF module tree of the structurally identical classes on disk (transformers
    `qwen3_omni_moe` talker / code predictor / Code2Wav, `mimi` split RVQ) plus the Qwen3-TTS specific names of SURVEY
    App. A.  They are rules, not facts: every rule is listed in KEY_RULES, `b200_key_map.json` in the model folder overrides
    or extends them, and the loader refuses to run with unmapped or missing tensors (it lists them) instead of guessing.
"""
from __future__ import annotations

import json
import os
import re
from typing import Dict, Iterable, List, Optional, Tuple

import torch

from .config import ModelConfig
from .weights import W8Triple, WeightStore, dequantize_w8

try:                                    # safetensors ships in the image; the loader says so if it ever does not
    from safetensors import safe_open
    from safetensors.torch import save_file
except Exception:                       # pragma: no cover
    safe_open = None
    save_file = None


# ------------------------------------------------------------------------------------------------------------------
# MLX affine 8-bit format
# ------------------------------------------------------------------------------------------------------------------
def unpack_mlx_affine(weight_u32: torch.Tensor, scales: torch.Tensor, biases: torch.Tensor, group: int = 64) -> W8Triple:
    """uint32 [N, K/4] (+ scales/biases [N, K/group]) -> (codes uint8 [N, K], scale bf16, bias bf16).  Bit-exact."""
    assert weight_u32.dtype in (torch.uint32, torch.int32), weight_u32.dtype
    n, k4 = weight_u32.shape
    k = k4 * 4
    assert scales.shape == (n, k // group) and biases.shape == (n, k // group), (weight_u32.shape, scales.shape, group)
    # element i of a word sits in bits [8i, 8i+8): on a little-endian host the byte view IS the code sequence
    q = weight_u32.contiguous().view(torch.uint8).reshape(n, k)
    return q.clone(), scales.to(torch.bfloat16).contiguous(), biases.to(torch.bfloat16).contiguous()


def pack_mlx_affine(q: torch.Tensor, scale: torch.Tensor, bias: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Inverse of `unpack_mlx_affine` (fixtures and tests)."""
    n, k = q.shape
    return q.contiguous().view(torch.int32).reshape(n, k // 4).view(torch.uint32), scale.contiguous(), bias.contiguous()


# ------------------------------------------------------------------------------------------------------------------
# checkpoint key  ->  name in the WeightStore
# ------------------------------------------------------------------------------------------------------------------
_LAYER = (
    (r"input_layernorm", "input_norm"), (r"post_attention_layernorm", "post_norm"),
    (r"self_attn\.q_proj", "q_proj"), (r"self_attn\.k_proj", "k_proj"), (r"self_attn\.v_proj", "v_proj"),
    (r"self_attn\.o_proj", "o_proj"), (r"self_attn\.q_norm", "q_norm"), (r"self_attn\.k_norm", "k_norm"),
    (r"mlp\.gate_proj", "gate_proj"), (r"mlp\.up_proj", "up_proj"), (r"mlp\.down_proj", "down_proj"),
)

# (regex on the key WITHOUT its trailing .weight/.bias/.scales/.biases/..., replacement).  First match wins.
KEY_RULES: List[Tuple[str, str]] = [
    # ---- talker (cousin: Qwen3OmniMoeTalkerModel / Qwen3Model layer names)
    (r"^talker\.model\.text_embedding$", "talker.text_embedding"),
    (r"^talker\.model\.codec_embedding$", "talker.codec_embedding"),
    (r"^talker\.text_projection\.(?:linear_)?fc1$", "talker.text_projection.fc1"),
    (r"^talker\.text_projection\.(?:linear_)?fc2$", "talker.text_projection.fc2"),
    (r"^talker\.model\.norm$", "talker.norm"),
    (r"^talker\.codec_head$", "talker.codec_head"),
] + [(rf"^talker\.model\.layers\.(\d+)\.{a}$", rf"talker.layers.\1.{b}") for a, b in _LAYER] + [
    # ---- code predictor (cousin: Qwen3OmniMoeTalkerCodePredictorModelForConditionalGeneration)
    (r"^talker\.code_predictor\.small_to_mtp_projection$", "cp.proj"),
    (r"^talker\.code_predictor\.model\.codec_embedding\.(\d+)$", r"cp.embeddings.\1"),
    (r"^talker\.code_predictor\.model\.norm$", "cp.norm"),
    (r"^talker\.code_predictor\.lm_head\.(\d+)$", r"cp.heads.\1"),
] + [(rf"^talker\.code_predictor\.model\.layers\.(\d+)\.{a}$", rf"cp.layers.\1.{b}") for a, b in _LAYER] + [
    # ---- speech tokenizer decoder (cousins: Qwen3OmniMoeCode2Wav*, MimiSplitResidualVectorQuantizer)
    (r"^decoder\.quantizer\.(?:rvq_first|semantic_residual_vector_quantizer)\.(?:vq\.)?layers\.(\d+)\.(?:_)?codebook\.(?:embedding_sum|embed_sum)$",
     r"codec.rvq.semantic.codebooks.\1.embed_sum"),
    (r"^decoder\.quantizer\.(?:rvq_first|semantic_residual_vector_quantizer)\.(?:vq\.)?layers\.(\d+)\.(?:_)?codebook\.cluster_usage$",
     r"codec.rvq.semantic.codebooks.\1.cluster_usage"),
    (r"^decoder\.quantizer\.(?:rvq_first|semantic_residual_vector_quantizer)\.output_proj$", "codec.rvq.semantic.out_proj"),
    (r"^decoder\.quantizer\.(?:rvq_rest|acoustic_residual_vector_quantizer)\.(?:vq\.)?layers\.(\d+)\.(?:_)?codebook\.(?:embedding_sum|embed_sum)$",
     r"codec.rvq.acoustic.codebooks.\1.embed_sum"),
    (r"^decoder\.quantizer\.(?:rvq_rest|acoustic_residual_vector_quantizer)\.(?:vq\.)?layers\.(\d+)\.(?:_)?codebook\.cluster_usage$",
     r"codec.rvq.acoustic.codebooks.\1.cluster_usage"),
    (r"^decoder\.quantizer\.(?:rvq_rest|acoustic_residual_vector_quantizer)\.output_proj$", "codec.rvq.acoustic.out_proj"),
    (r"^decoder\.pre_conv(?:\.conv)?$", "codec.pre_conv"),
    (r"^decoder\.pre_transformer\.input_proj$", "codec.tf.in_proj"),
    (r"^decoder\.pre_transformer\.output_proj$", "codec.tf.out_proj"),
    (r"^decoder\.pre_transformer\.norm$", "codec.tf.norm"),
    (r"^decoder\.pre_transformer\.layers\.(\d+)\.self_attn_layer_scale$", r"codec.tf.layers.\1.attn_scale"),
    (r"^decoder\.pre_transformer\.layers\.(\d+)\.mlp_layer_scale$", r"codec.tf.layers.\1.mlp_scale"),
] + [(rf"^decoder\.pre_transformer\.layers\.(\d+)\.{a}$", rf"codec.tf.layers.\1.{b}") for a, b in _LAYER] + [
    (r"^decoder\.upsample\.(\d+)\.0(?:\.conv)?$", r"codec.up.\1.tconv"),
    (r"^decoder\.upsample\.(\d+)\.1\.dwconv(?:\.conv)?$", r"codec.up.\1.cnx.dw"),
    (r"^decoder\.upsample\.(\d+)\.1\.norm$", r"codec.up.\1.cnx.ln"),
    (r"^decoder\.upsample\.(\d+)\.1\.pwconv1$", r"codec.up.\1.cnx.pw1"),
    (r"^decoder\.upsample\.(\d+)\.1\.pwconv2$", r"codec.up.\1.cnx.pw2"),
    (r"^decoder\.upsample\.(\d+)\.1$", r"codec.up.\1.cnx"),                       # .gamma
    (r"^decoder\.decoder\.0(?:\.conv)?$", "codec.dec.conv_in"),
    (r"^decoder\.decoder\.([1-4])\.block\.0$", lambda m: f"codec.dec.blocks.{int(m.group(1)) - 1}.snake"),
    (r"^decoder\.decoder\.([1-4])\.block\.1(?:\.conv)?$", lambda m: f"codec.dec.blocks.{int(m.group(1)) - 1}.tconv"),
    (r"^decoder\.decoder\.([1-4])\.block\.([2-4])\.act([12])$",
     lambda m: f"codec.dec.blocks.{int(m.group(1)) - 1}.units.{int(m.group(2)) - 2}.snake{m.group(3)}"),
    (r"^decoder\.decoder\.([1-4])\.block\.([2-4])\.conv([12])(?:\.conv)?$",
     lambda m: f"codec.dec.blocks.{int(m.group(1)) - 1}.units.{int(m.group(2)) - 2}.conv{m.group(3)}"),
    (r"^decoder\.decoder\.5$", "codec.dec.snake_out"),
    (r"^decoder\.decoder\.6(?:\.conv)?$", "codec.dec.conv_out"),
    # ---- speech tokenizer ENCODER (a Mimi model: cousin module tree transformers mimi/modeling_mimi.py:454-496, 1411-1452)
    (r"^encoder\.encoder\.layers\.0\.conv$", "enc.conv_in"),
    (r"^encoder\.encoder\.layers\.(\d+)\.block\.([13])\.conv$",
     lambda m: f"enc.stages.{(int(m.group(1)) - 1) // 3}.res.conv{1 if m.group(2) == '1' else 2}"),
    (r"^encoder\.encoder\.layers\.(\d+)\.conv$",
     lambda m: f"enc.stages.{int(m.group(1)) // 3 - 1}.down" if int(m.group(1)) % 3 == 0 else "enc.conv_out"),
    (r"^encoder\.encoder_transformer\.layers\.(\d+)\.self_attn_layer_scale$", r"enc.tf.layers.\1.attn_scale"),
    (r"^encoder\.encoder_transformer\.layers\.(\d+)\.mlp_layer_scale$", r"enc.tf.layers.\1.mlp_scale"),
    (r"^encoder\.encoder_transformer\.layers\.(\d+)\.mlp\.fc([12])$", r"enc.tf.layers.\1.fc\2"),
] + [(rf"^encoder\.encoder_transformer\.layers\.(\d+)\.{a}$", rf"enc.tf.layers.\1.{b}") for a, b in _LAYER[:6]] + [
    (r"^encoder\.downsample\.conv$", "enc.downsample"),
    (r"^encoder\.quantizer\.(semantic|acoustic)_residual_vector_quantizer\.input_proj$", r"enc.rvq.\1.in_proj"),
    (r"^encoder\.quantizer\.(semantic|acoustic)_residual_vector_quantizer\.layers\.(\d+)\.codebook\.(?:embedding_sum|embed_sum)$",
     r"enc.rvq.\1.codebooks.\2.embed_sum"),
    (r"^encoder\.quantizer\.(semantic|acoustic)_residual_vector_quantizer\.layers\.(\d+)\.codebook\.cluster_usage$",
     r"enc.rvq.\1.codebooks.\2.cluster_usage"),
    # ---- speaker encoder (ECAPA-TDNN: cousin module tree transformers qwen2_5_omni/modeling_qwen2_5_omni.py:2717-2790)
    (r"^speaker_encoder\.blocks\.0\.conv$", "spk.blocks.0.conv"),
    (r"^speaker_encoder\.blocks\.(\d+)\.tdnn([12])\.conv$", r"spk.blocks.\1.tdnn\2.conv"),
    (r"^speaker_encoder\.blocks\.(\d+)\.res2net_block\.blocks\.(\d+)\.conv$", r"spk.blocks.\1.res2net.\2.conv"),
    (r"^speaker_encoder\.blocks\.(\d+)\.se_block\.conv([12])$", r"spk.blocks.\1.se.conv\2"),
    (r"^speaker_encoder\.mfa\.conv$", "spk.mfa.conv"),
    (r"^speaker_encoder\.asp\.tdnn\.conv$", "spk.asp.tdnn.conv"),
    (r"^speaker_encoder\.asp\.conv$", "spk.asp.conv"),
    (r"^speaker_encoder\.fc$", "spk.fc"),
]

_SUFFIXES = (".weight", ".bias", ".scales", ".biases", ".alpha", ".beta", ".gamma", ".scale",
             ".embed_sum", ".embedding_sum", ".cluster_usage")
# tensors of the checkpoint that the generation hot path never reads (encoder side, speaker encoder, ...)
IGNORED = (r"^encoder\.(decoder|decoder_transformer|upsample)\.", r"^encoder\.quantizer\..*\.output_proj", r"\.rotary_emb\.",
           r"^decoder\.quantizer\..*\.input_proj", r"\.initialized$")


def split_key(key: str) -> Tuple[str, str]:
    for s in _SUFFIXES:
        if key.endswith(s):
            stem, suf = key[: -len(s)], s
            if suf in (".embed_sum", ".embedding_sum", ".cluster_usage"):
                return key, ""                    # the codebook rules match the full key
            return stem, suf
    return key, ""


def map_key(key: str, extra: Optional[Dict[str, str]] = None) -> Optional[str]:
    """Checkpoint key -> WeightStore name (with its suffix), or None if no rule applies."""
    if extra and key in extra:
        return extra[key]
    stem, suf = split_key(key)
    for pat, rep in KEY_RULES:
        m = re.match(pat, stem)
        if m:
            name = rep(m) if callable(rep) else m.expand(rep)
            if suf == ".scale":                   # LayerScale parameter of the codec transformer
                return name
            if suf == ".gamma":
                return name + ".gamma"
            return name + suf
    return None


def _iter_safetensors(path: str) -> Iterable[Tuple[str, torch.Tensor]]:
    if safe_open is None:
        raise RuntimeError("the safetensors package is required to load MLX checkpoints")
    files = [path]
    idx = path + ".index.json"
    if not os.path.exists(path) and os.path.exists(idx):
        with open(idx) as f:
            files = sorted({os.path.join(os.path.dirname(path), v) for v in json.load(f)["weight_map"].values()})
    for fn in files:
        with safe_open(fn, framework="pt", device="cpu") as f:
            for k in f.keys():
                yield k, f.get_tensor(k)


def _conv_layout_of(raw: Dict[str, torch.Tensor], expected: WeightStore) -> str:
    """'torch' ([Cout, Cin, k] / transposed [Cin, Cout, k], as weights.py) or 'mlx' (channel-last [Cout, k, Cin] for both).
    Square kernels cannot tell, so the decision is taken once from the tensors whose shape fits only one layout."""
    votes = {"torch": 0, "mlx": 0}
    for name, t in raw.items():
        exp = expected.fp.get(name)
        if exp is None or t.dim() != 3 or exp.dim() != 3:
            continue
        as_torch = tuple(t.shape) == tuple(exp.shape)
        perm = (2, 0, 1) if ".tconv." in name else (0, 2, 1)
        as_mlx = tuple(t.permute(*perm).shape) == tuple(exp.shape)
        if as_torch != as_mlx:
            votes["torch" if as_torch else "mlx"] += 1
    if votes["torch"] and votes["mlx"]:
        raise ValueError(f"conv weights mix layouts ({votes}); set b200_conv_layout in config.json")
    return "mlx" if votes["mlx"] else "torch"


def _conv_to_torch_layout(name: str, w: torch.Tensor, layout: str) -> torch.Tensor:
    """MLX keeps conv weights channel-last ([Cout, k, Cin], transposed convs too); the store uses the torch layouts of
    weights.py ([Cout, Cin, k] / [Cin, Cout, k])."""
    if w.dim() != 3 or layout == "torch":
        return w
    return w.permute(2, 0, 1).contiguous() if ".tconv." in name else w.permute(0, 2, 1).contiguous()


def load_mlx_checkpoint(model_path: str, cfg: ModelConfig, device: str = "cpu", expected: Optional[WeightStore] = None,
                        keep_fp: bool = False) -> WeightStore:
    """Reads `<model_path>/model.safetensors` (+ `speech_tokenizer/model.safetensors`) into a WeightStore.

    `expected` (a store of the same config; `weights.expected_shapes(cfg)` = meta tensors when omitted) is used for shape
    checks and conv layout detection."""
    extra = {}
    km = os.path.join(model_path, "b200_key_map.json")
    if os.path.exists(km):
        with open(km) as f:
            extra = json.load(f)
    if expected is None:
        from .weights import expected_shapes
        expected = expected_shapes(cfg)             # meta tensors: names + shapes only, nothing is materialised
        # (the reference-clip encoders are expected for the Base model; other folders may or may not carry them)
    ws = WeightStore(cfg)
    raw: Dict[str, torch.Tensor] = {}
    unmapped: List[str] = []
    srcs = [os.path.join(model_path, "model.safetensors"), os.path.join(model_path, "speech_tokenizer", "model.safetensors")]
    for si, src in enumerate(srcs):
        if not (os.path.exists(src) or os.path.exists(src + ".index.json")):
            if si == 0:
                raise OSError(f"{src} not found")
            continue
        for key, t in _iter_safetensors(src):
            if any(re.search(p, key) for p in IGNORED):
                continue
            name = map_key(key, extra)
            if name is None:
                unmapped.append(key)
                continue
            if name in raw:
                raise ValueError(f"MLX checkpoint keys collide on '{name}' (second key: {key}); fix b200_key_map.json")
            raw[name] = t
    if unmapped:
        raise ValueError("MLX checkpoint holds tensors no rule maps (add them to b200_key_map.json or IGNORED): "
                         + ", ".join(sorted(unmapped)[:12]) + (" ..." if len(unmapped) > 12 else ""))
    layout = None
    cj = os.path.join(model_path, "config.json")
    if os.path.exists(cj):
        with open(cj) as f:
            layout = json.load(f).get("b200_conv_layout")
    layout = layout or _conv_layout_of(raw, expected)
    # ---- assemble: quantised linears keep their codes; everything else is float
    done = set()
    for name in list(raw):
        if name.endswith(".scales"):
            stem = name[: -len(".scales")]
            q, s, b = unpack_mlx_affine(raw[stem + ".weight"], raw[name], raw[stem + ".biases"], cfg.quant_group)
            done.update({stem + ".weight", name, stem + ".biases"})
            if stem in expected.q:
                assert tuple(q.shape) == tuple(expected.q[stem][0].shape), (stem, q.shape, expected.q[stem][0].shape)
                ws.q[stem] = (q.to(device), s.to(device), b.to(device))
                if keep_fp:
                    ws.fp[stem + ".weight"] = dequantize_w8(q, s, b, cfg.quant_group).to(device)
            else:                                   # quantised embedding table: gathered rows are float here
                ws.fp[stem if stem + ".weight" not in expected.fp else stem + ".weight"] = dequantize_w8(q, s, b, cfg.quant_group).to(device)
    for name, t in raw.items():
        if name in done:
            continue
        tgt = name
        if tgt not in expected.fp and tgt.endswith(".weight") and tgt[: -len(".weight")] in expected.fp:
            tgt = tgt[: -len(".weight")]           # embedding tables are stored without a suffix
        stem = tgt[: -len(".weight")] if tgt.endswith(".weight") else None
        if stem is not None and stem in expected.q and tgt not in expected.fp:
            # a float linear where this build keeps W8: quantise with the same affine scheme
            from .weights import quantize_w8
            ws.q[stem] = tuple(x.to(device) for x in quantize_w8(t.float(), cfg.quant_group))
            continue
        exp = expected.fp.get(tgt)
        t = _conv_to_torch_layout(tgt, t.float(), layout)
        if exp is not None and tuple(t.shape) != tuple(exp.shape):
            if t.dim() == 3 and t.shape[-1] == 1 and tuple(t.shape[:2]) == tuple(exp.shape):
                t = t[..., 0]                      # 1x1 conv stored as a linear (RVQ output projections)
            else:
                raise ValueError(f"{tgt}: shape {tuple(t.shape)} != expected {tuple(exp.shape)}")
        ws.fp[tgt] = t.to(device).contiguous()
    missing = [n for n in expected.fp if n not in ws.fp and not (n.endswith(".weight") and n[: -len('.weight')] in ws.q)]
    missing += [n for n in expected.q if n not in ws.q]
    if missing:
        raise ValueError("MLX checkpoint is missing tensors the hot path needs: " + ", ".join(sorted(missing)[:12])
                         + (" ..." if len(missing) > 12 else ""))
    return ws


# ------------------------------------------------------------------------------------------------------------------
# writer (fixtures / round-trip tests): a WeightStore as an mlx-community style folder
# ------------------------------------------------------------------------------------------------------------------
_INV_LAYER = {b: a.replace("\\", "") for a, b in _LAYER}


def _export_name(name: str) -> Tuple[str, str]:
    """WeightStore name -> (file, checkpoint key stem) using the primary spelling of every rule."""
    m = re.match(r"^(talker|cp)\.layers\.(\d+)\.(\w+)$", name)
    if m:
        root = "talker.model" if m.group(1) == "talker" else "talker.code_predictor.model"
        return "model", f"{root}.layers.{m.group(2)}.{_INV_LAYER[m.group(3)]}"
    m = re.match(r"^codec\.tf\.layers\.(\d+)\.(\w+)$", name)
    if m:
        part = {"attn_scale": "self_attn_layer_scale", "mlp_scale": "mlp_layer_scale"}.get(m.group(2)) or _INV_LAYER[m.group(2)]
        return "speech", f"decoder.pre_transformer.layers.{m.group(1)}.{part}"
    table = {
        "talker.text_embedding": "talker.model.text_embedding", "talker.codec_embedding": "talker.model.codec_embedding",
        "talker.text_projection.fc1": "talker.text_projection.linear_fc1", "talker.text_projection.fc2": "talker.text_projection.linear_fc2",
        "talker.norm": "talker.model.norm", "talker.codec_head": "talker.codec_head",
        "cp.proj": "talker.code_predictor.small_to_mtp_projection", "cp.norm": "talker.code_predictor.model.norm",
    }
    if name in table:
        return "model", table[name]
    m = re.match(r"^cp\.embeddings\.(\d+)$", name)
    if m:
        return "model", f"talker.code_predictor.model.codec_embedding.{m.group(1)}"
    m = re.match(r"^cp\.heads\.(\d+)$", name)
    if m:
        return "model", f"talker.code_predictor.lm_head.{m.group(1)}"
    m = re.match(r"^codec\.rvq\.(semantic|acoustic)\.codebooks\.(\d+)\.(embed_sum|cluster_usage)$", name)
    if m:
        grp = "rvq_first" if m.group(1) == "semantic" else "rvq_rest"
        return "speech", f"decoder.quantizer.{grp}.vq.layers.{m.group(2)}._codebook.{ 'embedding_sum' if m.group(3) == 'embed_sum' else 'cluster_usage'}"
    m = re.match(r"^codec\.rvq\.(semantic|acoustic)\.out_proj$", name)
    if m:
        return "speech", f"decoder.quantizer.{'rvq_first' if m.group(1) == 'semantic' else 'rvq_rest'}.output_proj"
    simple = {"codec.pre_conv": "decoder.pre_conv.conv", "codec.tf.in_proj": "decoder.pre_transformer.input_proj",
              "codec.tf.out_proj": "decoder.pre_transformer.output_proj", "codec.tf.norm": "decoder.pre_transformer.norm",
              "codec.dec.conv_in": "decoder.decoder.0.conv", "codec.dec.snake_out": "decoder.decoder.5",
              "codec.dec.conv_out": "decoder.decoder.6.conv"}
    if name in simple:
        return "speech", simple[name]
    m = re.match(r"^codec\.up\.(\d+)\.(tconv|cnx\.dw|cnx\.ln|cnx\.pw1|cnx\.pw2|cnx)$", name)
    if m:
        part = {"tconv": "0.conv", "cnx.dw": "1.dwconv.conv", "cnx.ln": "1.norm", "cnx.pw1": "1.pwconv1", "cnx.pw2": "1.pwconv2",
                "cnx": "1"}[m.group(2)]
        return "speech", f"decoder.upsample.{m.group(1)}.{part}"
    m = re.match(r"^codec\.dec\.blocks\.(\d+)\.(snake|tconv)$", name)
    if m:
        return "speech", f"decoder.decoder.{int(m.group(1)) + 1}.block.{'0' if m.group(2) == 'snake' else '1.conv'}"
    m = re.match(r"^codec\.dec\.blocks\.(\d+)\.units\.(\d+)\.(snake|conv)([12])$", name)
    if m:
        part = f"act{m.group(4)}" if m.group(3) == "snake" else f"conv{m.group(4)}.conv"
        return "speech", f"decoder.decoder.{int(m.group(1)) + 1}.block.{int(m.group(2)) + 2}.{part}"
    m = re.match(r"^enc\.stages\.(\d+)\.(res\.conv1|res\.conv2|down)$", name)
    if m:
        i = int(m.group(1))
        part = {"res.conv1": f"{1 + 3 * i}.block.1.conv", "res.conv2": f"{1 + 3 * i}.block.3.conv", "down": f"{3 + 3 * i}.conv"}[m.group(2)]
        return "speech", f"encoder.encoder.layers.{part}"
    if name == "enc.conv_in":
        return "speech", "encoder.encoder.layers.0.conv"
    if name == "enc.conv_out":
        return "speech", f"encoder.encoder.layers.{2 + 3 * _N_ENC_STAGES[0]}.conv"
    if name == "enc.downsample":
        return "speech", "encoder.downsample.conv"
    m = re.match(r"^enc\.tf\.layers\.(\d+)\.(\w+)$", name)
    if m:
        part = {"attn_scale": "self_attn_layer_scale", "mlp_scale": "mlp_layer_scale", "fc1": "mlp.fc1", "fc2": "mlp.fc2"}.get(m.group(2)) \
            or _INV_LAYER[m.group(2)]
        return "speech", f"encoder.encoder_transformer.layers.{m.group(1)}.{part}"
    m = re.match(r"^enc\.rvq\.(semantic|acoustic)\.in_proj$", name)
    if m:
        return "speech", f"encoder.quantizer.{m.group(1)}_residual_vector_quantizer.input_proj"
    m = re.match(r"^enc\.rvq\.(semantic|acoustic)\.codebooks\.(\d+)\.(embed_sum|cluster_usage)$", name)
    if m:
        return "speech", f"encoder.quantizer.{m.group(1)}_residual_vector_quantizer.layers.{m.group(2)}.codebook.{m.group(3)}"
    m = re.match(r"^spk\.blocks\.(\d+)\.(tdnn[12]\.conv|res2net\.(\d+)\.conv|se\.conv[12])$", name)
    if m:
        part = m.group(2)
        part = part.replace("res2net.", "res2net_block.blocks.").replace("se.", "se_block.")
        return "model", f"speaker_encoder.blocks.{m.group(1)}.{part}"
    if name.startswith("spk."):
        return "model", "speaker_encoder." + name[len("spk."):]
    raise KeyError(name)


_N_ENC_STAGES = [4]


def export_mlx_checkpoint(ws: WeightStore, model_path: str, extra_config: Optional[dict] = None) -> None:
    """Writes `ws` as `<model_path>/{config.json, model.safetensors, speech_tokenizer/model.safetensors}` in the MLX
    affine format and the key spelling of KEY_RULES (used by the round-trip tests and to hand-build fixtures)."""
    if save_file is None:
        raise RuntimeError("the safetensors package is required")
    out = {"model": {}, "speech": {}}
    _N_ENC_STAGES[0] = len(ws.cfg.enc.ratios)
    for stem, (q, s, b) in ws.q.items():
        f, key = _export_name(stem)
        w, sc, bi = pack_mlx_affine(q.cpu(), s.cpu(), b.cpu())
        out[f][key + ".weight"] = w.view(torch.int32).view(torch.uint32) if w.dtype != torch.uint32 else w
        out[f][key + ".scales"], out[f][key + ".biases"] = sc, bi
    for name, t in ws.fp.items():
        stem, suf = name, ""
        for s_ in (".weight", ".bias", ".alpha", ".beta", ".gamma"):
            if name.endswith(s_):
                stem, suf = name[: -len(s_)], s_
                break
        if suf == ".weight" and stem in ws.q:
            continue                                   # de-quantised copy of a W8 linear
        if name.endswith(".attn_scale") or name.endswith(".mlp_scale"):
            f, key = _export_name(name)
            out[f][key + ".scale"] = t.cpu().contiguous()
            continue
        if re.search(r"\.(embed_sum|cluster_usage)$", name):
            f, key = _export_name(name)
            out[f][key] = t.cpu().contiguous()
            continue
        f, key = _export_name(stem)
        if suf == "":
            suf = ".weight"                            # embedding tables
        out[f][key + suf] = t.cpu().contiguous()
    os.makedirs(os.path.join(model_path, "speech_tokenizer"), exist_ok=True)
    save_file(out["model"], os.path.join(model_path, "model.safetensors"))
    if out["speech"]:
        save_file(out["speech"], os.path.join(model_path, "speech_tokenizer", "model.safetensors"))
    # the checkpoint's own config layout (talker_config / code_predictor_config / spk_id / ... + speech_tokenizer/config.json):
    # load_model reads the architecture back through config.from_hf_config, exactly as it would for a downloaded folder
    from .config import to_hf_config
    meta, speech = to_hf_config(ws.cfg)
    meta.update(extra_config or {})
    with open(os.path.join(model_path, "config.json"), "w") as f:
        json.dump(meta, f)
    with open(os.path.join(model_path, "speech_tokenizer", "config.json"), "w") as f:
        json.dump(speech, f)
