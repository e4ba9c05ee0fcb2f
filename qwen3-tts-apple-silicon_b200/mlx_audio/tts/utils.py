"""`from mlx_audio.tts.utils import load_model` -> B200 backend (reference io.py:111-112)."""
from qwen3_tts_b200.model import load_model  # noqa: F401
