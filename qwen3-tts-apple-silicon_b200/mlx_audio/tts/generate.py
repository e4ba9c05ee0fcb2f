"""`from mlx_audio.tts.generate import generate_audio` -> B200 backend.

Keyword sets used by the reference: custom.py:163-170 (model, text, voice, instruct, speed, output_path),
design.py:76-81 (model, text, instruct, output_path), clone.py:218-224 (model, text, ref_audio, ref_text,
output_path).  Contract: on success `<output_path>/audio_000.wav` exists (io.py:156-158), mono 24 kHz."""
import os

import numpy as np

from qwen3_tts_b200.model import write_wav


def generate_audio(text, model=None, voice=None, instruct=None, speed=1.0, lang_code="auto", ref_audio=None,
                   ref_text=None, output_path=None, file_prefix="audio", audio_format="wav", verbose=False,
                   join_audio=True, **kwargs):
    if model is None:
        raise ValueError("generate_audio: `model` is required (pass the object returned by load_model)")
    if text is None or not str(text).strip():
        raise ValueError("generate_audio: empty text")
    out_dir = output_path or "."
    os.makedirs(out_dir, exist_ok=True)
    chunks, last = [], None
    for res in model.generate(text=text, voice=voice, instruct=instruct, speed=speed, lang_code=lang_code,
                              ref_audio=ref_audio, ref_text=ref_text, verbose=verbose, **kwargs):
        chunks.append(res.audio)
        last = res
    # the reference only ever picks up audio_000.wav (io.py:156): always emit ONE joined file
    audio = np.concatenate(chunks) if chunks else np.zeros(0, dtype=np.float32)
    path = os.path.join(out_dir, f"{file_prefix}_000.{audio_format}")
    write_wav(path, audio, model.sample_rate)
    if verbose and last is not None:
        print(f"frames={last.token_count} audio={last.audio_duration:.2f}s rtfx={last.real_time_factor:.1f}")
    return path
