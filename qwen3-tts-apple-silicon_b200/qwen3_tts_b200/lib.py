"""ctypes binding of libq3tts_b200.so (the C ABI declared in include/q3tts_b200.h).

There is NO fallback: if the CUDA extension is missing the import of the product path fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# Q3T_LIB selects another build of the same library (kernel-tuning experiments); the default is the in-tree build
LIB_PATH = os.environ.get("Q3T_LIB") or os.path.join(_HERE, "libq3tts_b200.so")

TILE_BYTES = 4352
KV_PAGE = 16
PRO_RAW, PRO_RMSNORM, PRO_SWIGLU = 0, 1, 2
ACT_NONE, ACT_SILU, ACT_GELU, ACT_SNAKE, ACT_SWIGLU_PAIR, ACT_ELU, ACT_RELU, ACT_SIGMOID, ACT_TANH = range(9)

vp, i32, f32, i64, u64 = C.c_void_p, C.c_int, C.c_float, C.c_longlong, C.c_ulonglong


class W8(C.Structure):
    _fields_ = [("w", vp), ("N", i32), ("K", i32), ("lin_bias", vp)]


class GemvArgs(C.Structure):
    _fields_ = [("w", W8), ("M", i32), ("prologue", i32), ("x", vp), ("x_stride", i64), ("norm_w", vp), ("eps", f32),
                ("gather_idx", vp), ("gather_idx_stride", i32), ("gather_row_stride", i64), ("act", i32),
                ("resid", vp), ("resid_stride", i64), ("y", vp), ("y_stride", i64)]


class GemmArgs(C.Structure):
    _fields_ = [("w", W8), ("M", i32), ("prologue", i32), ("x", vp), ("x_stride", i64), ("norm_w", vp), ("eps", f32),
                ("gather_idx", vp), ("gather_idx_stride", i32), ("gather_row_stride", i64), ("act", i32),
                ("swiglu_out", i32), ("resid", vp), ("resid_stride", i64), ("y", vp), ("y_stride", i64), ("xb", vp),
                ("splitk_ws", vp), ("splitk_ws_floats", i64), ("splitk_counters", vp), ("x_bf16", vp), ("y_bf16", vp),
                ("y_norm_w", vp), ("y_rowss", vp), ("x_rowss", vp), ("x_rowss_parts", i32)]


class AttnArgs(C.Structure):
    _fields_ = [("qkv", vp), ("q_norm_w", vp), ("k_norm_w", vp), ("eps", f32), ("inv_freq", vp), ("kv_pool", vp),
                ("block_tbl", vp), ("max_pages", i32), ("pos", vp), ("out", vp), ("work", vp), ("counters", vp),
                ("B", i32), ("H", i32), ("Hkv", i32), ("D", i32), ("nsplit", i32), ("mode", i32), ("seq_of_row", vp), ("out_bf16", vp)]


class AttnPrefillArgs(C.Structure):
    _fields_ = [("qkv", vp), ("q_norm_w", vp), ("eps", f32), ("inv_freq", vp), ("kv_pool", vp), ("block_tbl", vp),
                ("max_pages", i32), ("pos", vp), ("seq_of_row", vp), ("blocks", vp), ("n_blocks", i32), ("out", vp),
                ("out_bf16", vp), ("H", i32), ("Hkv", i32), ("D", i32), ("k_norm_w", vp), ("M", i32)]


class Sampling(C.Structure):
    _fields_ = [("do_sample", i32), ("temperature", f32), ("top_k", i32), ("top_p", f32),
                ("repetition_penalty", f32), ("min_new_tokens", i32), ("suppress_lo", i32), ("suppress_hi", i32),
                ("eos_id", i32), ("seed", u64)]


class SampleArgs(C.Structure):
    _fields_ = [("logits", vp), ("B", i32), ("V", i32), ("logits_stride", i64), ("sp", Sampling), ("seen", vp),
                ("step", vp), ("rng_stream", i32), ("uniforms", vp), ("out", vp), ("out_stride", i64),
                ("fo_stride", i64), ("fo_step_stride", i64), ("forced", vp), ("own", vp), ("done", vp), ("step_stride", i32)]


class Layer(C.Structure):
    _fields_ = [("input_norm", vp), ("qkv", W8), ("q_norm", vp), ("k_norm", vp), ("o", W8), ("post_norm", vp),
                ("gate_up", W8), ("down", W8)]


class Stack(C.Structure):
    _fields_ = [("hidden", i32), ("n_layers", i32), ("n_heads", i32), ("n_kv_heads", i32), ("head_dim", i32),
                ("inter", i32), ("eps", f32), ("layers_host", C.POINTER(Layer)), ("layers_dev", vp), ("final_norm", vp), ("inv_freq", vp),
                ("kv_pool", vp), ("kv_layer_stride_bytes", i64), ("block_tbl", vp), ("max_pages", i32),
                ("attn_nsplit", i32)]


class FrameArgs(C.Structure):
    _fields_ = [("B", i32), ("talker", Stack), ("codec_head", W8), ("talker_vocab", i32), ("codec_embedding", vp),
                ("talker_sp", Sampling), ("cp", Stack), ("cp_proj", W8), ("cp_embeddings_host", C.POINTER(vp)),
                ("cp_embeddings_dev", vp), ("cp_heads_host", C.POINTER(W8)), ("cp_vocab", i32), ("n_groups", i32),
                ("cp_sp", Sampling), ("x", vp), ("hidden", vp), ("logits", vp), ("cp_logits", vp),
                ("keep_cp_logits", i32), ("xc", vp), ("qkv", vp), ("attn", vp), ("gu", vp), ("attn_work", vp),
                ("attn_counters", vp), ("pos", vp), ("cp_pos", vp), ("step", vp), ("cur_codes", vp), ("codes", vp),
                ("own_codes", vp), ("max_frames", i32), ("seen", vp), ("done", vp), ("trailing", vp),
                ("n_trailing", i32), ("forced_codes", vp), ("gemm_xb", vp), ("gemm_ws", vp), ("gemm_ws_floats", i64), ("gemm_counters", vp), ("use_mega", i32), ("cp_heads_dev", vp), ("ll_work", vp),
                ("ll_work_bytes", i64), ("ll_state", vp), ("ll_timing", vp), ("gemm_xb2", vp), ("cp_proj_rows_dev", vp), ("cp_qkv0_rows_dev", vp), ("step_per_row", i32), ("active", vp),
                ("gemm_rowss", vp)]


class PrefillArgs(C.Structure):
    _fields_ = [("f", C.POINTER(FrameArgs)), ("M", i32), ("x", vp), ("pos", vp), ("seq_of_row", vp), ("qkv", vp), ("attn", vp),
                ("gu", vp), ("xb", vp), ("attn_work", vp), ("attn_counters", vp), ("blocks", vp), ("n_blocks", i32), ("xb2", vp), ("rowss", vp)]


class StackPassArgs(C.Structure):
    _fields_ = [("stack", Stack), ("head", W8), ("pos", vp), ("x_in", vp), ("hidden_out", vp), ("logits_out", vp),
                ("ll_work", vp), ("ll_work_bytes", i64), ("ll_state", vp), ("timing", vp)]


class TapGemmArgs(C.Structure):
    _fields_ = [("A", vp), ("B", i32), ("T_in", i32), ("Cin", i32), ("W", vp), ("bias", vp), ("taps", i32),
                ("shift", i32 * 8), ("up", i32), ("Cout", i32), ("T_out_rows", i32), ("scale", vp), ("resid", vp),
                ("out_raw", vp), ("out_act", vp), ("act", i32), ("act_a", vp), ("act_b", vp), ("force_fp32", i32),
                ("a_f16", i32), ("act_f16", i32)]


# every symbol include/q3tts_b200.h declares (tests check the .so exports all of them)
SYMBOLS = ["q3t_abi_version", "q3t_last_error", "q3t_launch_count", "q3t_w8_gemv", "q3t_w8_gemv_rows", "q3t_w8_gemm", "q3t_rmsnorm", "q3t_attn_decode", "q3t_attn_prefill",
           "q3t_sample", "q3t_stack_pass", "q3t_ll_work_bytes", "q3t_talker_step", "q3t_frame", "q3t_talker_prefill", "q3t_talker_tail", "q3t_rvq_gather_sum", "q3t_tapgemm", "q3t_dwconv_ln",
           "q3t_tapgemm_stats", "q3t_window_attn", "q3t_snake", "q3t_conv_out_clamp", "q3t_conv_out_clamp_h", "q3t_tapgemm_tc_eligible", "q3t_clamp_pcm16",
           "q3t_rvq_encode", "q3t_time_stats", "q3t_softmax_time", "q3t_eltwise", "q3t_mel", "q3t_layernorm"]

_lib = None


class Q3TError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the CUDA extension; raise (never fall back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Q3TError(f"{LIB_PATH} is missing: build it with qwen3-tts-apple-silicon_b200/csrc/build.sh "
                       "(or __graft_entry__.build()); there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    lib.q3t_last_error.restype = C.c_char_p
    lib.q3t_launch_count.restype = u64
    lib.q3t_w8_gemv.argtypes = [C.POINTER(GemvArgs), vp]
    lib.q3t_w8_gemv_rows.argtypes = [C.POINTER(GemvArgs), i32, vp]
    lib.q3t_w8_gemm.argtypes = [C.POINTER(GemmArgs), vp]
    lib.q3t_rmsnorm.argtypes = [vp, vp, vp, i32, i32, f32, vp]
    lib.q3t_attn_decode.argtypes = [C.POINTER(AttnArgs), vp]
    lib.q3t_attn_prefill.argtypes = [C.POINTER(AttnPrefillArgs), vp]
    lib.q3t_sample.argtypes = [C.POINTER(SampleArgs), vp]
    lib.q3t_stack_pass.argtypes = [C.POINTER(StackPassArgs), vp]
    lib.q3t_ll_work_bytes.argtypes = [C.POINTER(Stack), C.POINTER(Stack), i32]
    lib.q3t_ll_work_bytes.restype = i64
    lib.q3t_talker_step.argtypes = [C.POINTER(FrameArgs), i32, vp]
    lib.q3t_frame.argtypes = [C.POINTER(FrameArgs), vp]
    lib.q3t_talker_prefill.argtypes = [C.POINTER(PrefillArgs), vp]
    lib.q3t_talker_tail.argtypes = [C.POINTER(FrameArgs), vp]
    lib.q3t_rvq_gather_sum.argtypes = [vp, C.POINTER(vp), i32, i32, i32, i32, i32, i32, i32, vp, vp]
    lib.q3t_tapgemm.argtypes = [C.POINTER(TapGemmArgs), vp]
    lib.q3t_tapgemm_stats.argtypes = [C.POINTER(u64 * 3), i32]
    lib.q3t_tapgemm_stats.restype = None
    lib.q3t_dwconv_ln.argtypes = [vp, vp, vp, vp, vp, f32, i32, i32, i32, i32, vp, vp]
    lib.q3t_window_attn.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp, vp]
    lib.q3t_snake.argtypes = [vp, vp, vp, i64, i32, vp, vp]
    lib.q3t_conv_out_clamp.argtypes = [vp, i32, i32, i32, vp, vp, i32, vp, vp]
    lib.q3t_conv_out_clamp_h.argtypes = [vp, i32, i32, i32, vp, vp, i32, vp, vp]
    lib.q3t_tapgemm_tc_eligible.argtypes = [C.POINTER(TapGemmArgs)]
    lib.q3t_clamp_pcm16.argtypes = [vp, i64, vp, vp, vp]
    lib.q3t_rvq_encode.argtypes = [vp, C.POINTER(vp), i32, i32, i32, i32, i64, vp, vp, vp]
    lib.q3t_time_stats.argtypes = [vp, vp, i32, i32, i32, f32, vp, vp, vp]
    lib.q3t_softmax_time.argtypes = [vp, i32, i32, i32, vp, vp]
    lib.q3t_eltwise.argtypes = [i32, vp, vp, vp, i64, i32, i64, vp, vp]
    lib.q3t_mel.argtypes = [vp, i64, i32, i32, vp, i32, vp, vp]
    lib.q3t_layernorm.argtypes = [vp, vp, vp, i64, i32, f32, vp, vp]
    if lib.q3t_abi_version() != 2:
        raise Q3TError("libq3tts_b200.so ABI version mismatch")
    _lib = lib
    return lib


def tapgemm_stats(reset: bool = False):
    """(tcgen05 launches, FP32-pipe launches of tensor-core-eligible channel counts, FP32-pipe launches of other shapes)."""
    out = (u64 * 3)()
    load().q3t_tapgemm_stats(C.byref(out), int(reset))
    return tuple(int(v) for v in out)


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise Q3TError(f"{what}: {load().q3t_last_error().decode()}")


def ptr(t) -> int:
    """Device pointer of a torch tensor (or 0 for None)."""
    return 0 if t is None else t.data_ptr()


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
