"""Drop-in shim: the reference app imports `mlx_audio.tts.utils.load_model` and
`mlx_audio.tts.generate.generate_audio` (reference app.py:50-51, io.py:111, sessions/custom.py:28).
Putting this directory on sys.path routes both names to the B200 backend; nothing of MLX is here."""
