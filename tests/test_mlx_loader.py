"""MLX-quantised safetensors loader: format, key rules, round trip.  CPU only."""
import json
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qwen3-tts-apple-silicon_b200"))
from qwen3_tts_b200 import config as Cfg
from qwen3_tts_b200 import mlx_loader as ML
from qwen3_tts_b200.weights import dequantize_w8, make_weights, quantize_w8


def test_affine_word_packing_is_little_endian_along_the_input_axis():
    # element i of a uint32 word sits in bits [8i, 8i+8)
    q = torch.arange(2 * 64, dtype=torch.int64).reshape(2, 64).to(torch.uint8)
    s = torch.ones(2, 1, dtype=torch.bfloat16)
    b = torch.zeros(2, 1, dtype=torch.bfloat16)
    w, _, _ = ML.pack_mlx_affine(q, s, b)
    assert w.dtype == torch.uint32 and tuple(w.shape) == (2, 16)
    words = w.view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    for i in range(4):
        assert torch.equal(((words >> (8 * i)) & 0xFF).to(torch.uint8), q[:, i::4])
    q2, s2, b2 = ML.unpack_mlx_affine(w, s, b)
    assert torch.equal(q2, q) and torch.equal(s2, s) and torch.equal(b2, b)


def test_unpack_matches_the_published_dequantisation():
    torch.manual_seed(0)
    w = torch.randn(32, 256) * 0.02
    q, s, b = quantize_w8(w)
    wq, ws_, wb = ML.pack_mlx_affine(q, s, b)
    # MLX dequantisation written out from the format description: w = scale * code + bias per group of 64
    codes = torch.stack([((wq.view(torch.int32).to(torch.int64) >> (8 * i)) & 0xFF) for i in range(4)], -1).reshape(32, 256).float()
    ref = codes.reshape(32, 4, 64) * ws_.float().unsqueeze(-1) + wb.float().unsqueeze(-1)
    assert torch.equal(ref.reshape(32, 256), dequantize_w8(*ML.unpack_mlx_affine(wq, ws_, wb)))


@pytest.mark.parametrize("key,name", [
    ("talker.model.layers.3.self_attn.q_proj.scales", "talker.layers.3.q_proj.scales"),
    ("talker.model.layers.27.post_attention_layernorm.weight", "talker.layers.27.post_norm.weight"),
    ("talker.model.layers.0.self_attn.k_norm.weight", "talker.layers.0.k_norm.weight"),
    ("talker.model.text_embedding.weight", "talker.text_embedding.weight"),
    ("talker.text_projection.linear_fc2.bias", "talker.text_projection.fc2.bias"),
    ("talker.codec_head.biases", "talker.codec_head.biases"),
    ("talker.code_predictor.small_to_mtp_projection.bias", "cp.proj.bias"),
    ("talker.code_predictor.model.codec_embedding.14.weight", "cp.embeddings.14.weight"),
    ("talker.code_predictor.lm_head.0.weight", "cp.heads.0.weight"),
    ("talker.code_predictor.model.layers.4.mlp.down_proj.weight", "cp.layers.4.down_proj.weight"),
    ("decoder.quantizer.rvq_rest.vq.layers.7._codebook.embedding_sum", "codec.rvq.acoustic.codebooks.7.embed_sum"),
    ("decoder.quantizer.semantic_residual_vector_quantizer.layers.0.codebook.cluster_usage", "codec.rvq.semantic.codebooks.0.cluster_usage"),
    ("decoder.pre_transformer.layers.2.self_attn_layer_scale.scale", "codec.tf.layers.2.attn_scale"),
    ("decoder.upsample.1.1.gamma", "codec.up.1.cnx.gamma"),
    ("decoder.upsample.0.0.conv.weight", "codec.up.0.tconv.weight"),
    ("decoder.decoder.2.block.3.act2.beta", "codec.dec.blocks.1.units.1.snake2.beta"),
    ("decoder.decoder.4.block.1.conv.bias", "codec.dec.blocks.3.tconv.bias"),
    ("decoder.decoder.6.conv.weight", "codec.dec.conv_out.weight"),
])
def test_key_rules(key, name):
    assert ML.map_key(key) == name


def test_unknown_keys_are_not_guessed():
    assert ML.map_key("talker.model.layers.0.self_attn.rotary_fn.weight") is None


@pytest.fixture(scope="module")
def exported(tmp_path_factory):
    cfg = Cfg.small("custom_voice")
    ws = make_weights(cfg, seed=3, device="cpu", keep_fp=True, keep_q=True)
    d = str(tmp_path_factory.mktemp("Qwen3-TTS-12Hz-small-CustomVoice-8bit"))
    ML.export_mlx_checkpoint(ws, d)
    return cfg, ws, d


def test_round_trip_is_bit_exact(exported):
    cfg, ws, d = exported
    assert os.path.exists(os.path.join(d, "model.safetensors")) and os.path.exists(os.path.join(d, "speech_tokenizer", "model.safetensors"))
    got = ML.load_mlx_checkpoint(d, cfg, device="cpu", keep_fp=True)
    assert set(got.q) == set(ws.q)
    for k, (q, s, b) in ws.q.items():
        assert torch.equal(got.q[k][0], q) and torch.equal(got.q[k][1], s) and torch.equal(got.q[k][2], b), k
    for k, t in ws.fp.items():
        assert k in got.fp, k
        assert torch.equal(got.fp[k], t), k


def test_channel_last_conv_weights_are_detected(exported, tmp_path):
    # MLX keeps conv weights channel-last; the loader recognises the layout from the expected shape
    from safetensors.torch import load_file, save_file
    cfg, ws, d = exported
    import shutil
    d2 = str(tmp_path / "m")
    shutil.copytree(d, d2)
    fn = os.path.join(d2, "speech_tokenizer", "model.safetensors")
    st = {k: v.clone() for k, v in load_file(fn).items()}      # not mmap-backed: the file is rewritten below
    for k in list(st):
        if st[k].dim() == 3 and ("decoder.decoder" in k or "upsample" in k or "pre_conv" in k):
            name = ML.map_key(k)
            if ".tconv." in name:
                st[k] = st[k].permute(1, 2, 0).contiguous()      # [Cin, Cout, k] -> [Cout, k, Cin]
            else:
                st[k] = st[k].permute(0, 2, 1).contiguous()      # [Cout, Cin, k] -> [Cout, k, Cin]
    save_file(st, fn)
    got = ML.load_mlx_checkpoint(d2, cfg, device="cpu", keep_fp=True)
    for k, t in ws.fp.items():
        if k.startswith("codec."):
            assert torch.equal(got.fp[k], t), k


def test_missing_and_unmapped_tensors_are_reported(exported, tmp_path):
    from safetensors.torch import load_file, save_file
    import shutil
    cfg, ws, d = exported
    d2 = str(tmp_path / "m")
    shutil.copytree(d, d2)
    fn = os.path.join(d2, "model.safetensors")
    st = {k: v.clone() for k, v in load_file(fn).items()}      # not mmap-backed: the file is rewritten below
    del st["talker.codec_head.scales"], st["talker.codec_head.biases"], st["talker.codec_head.weight"]
    save_file(st, fn)
    with pytest.raises(ValueError, match="missing"):
        ML.load_mlx_checkpoint(d2, cfg, device="cpu")
    st["some.new.module.weight"] = torch.zeros(4)
    save_file(st, fn)
    with pytest.raises(ValueError, match="no rule maps"):
        ML.load_mlx_checkpoint(d2, cfg, device="cpu")
    with open(os.path.join(d2, "b200_key_map.json"), "w") as f:
        json.dump({"some.new.module.weight": "talker.codec_head.weight"}, f)
    with pytest.raises((ValueError, AssertionError)):
        ML.load_mlx_checkpoint(d2, cfg, device="cpu")       # mapped now, but the shape check refuses it
