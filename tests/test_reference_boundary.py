"""Drives the UNMODIFIED reference session loops (read from /root/reference, CPU container only) against our
`mlx_audio` shim, following the scripted recipe of SURVEY.md 8b.  The model object is a host stub here (no GPU in
this container): what is under test is the boundary -- module paths, keyword sets, the audio_000.wav contract and
error behaviour.  The real CUDA model behind the same shim is exercised by tests/test_gpu_pipeline.py."""
import os
import subprocess
import sys
import textwrap
import wave

import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "qwen3_tts")),
                                reason="reference checkout not present (GPU box)")

DRIVER = textwrap.dedent('''
    import json, os, sys
    import numpy as np
    import mlx_audio.tts.utils as U
    import qwen3_tts_b200.model as M
    calls = []
    class StubModel:
        sample_rate = 24000
        def generate(self, **kw):
            calls.append(dict((k, v) for k, v in kw.items() if v is not None))
            n = 2400
            yield M.GenerationResult(audio=(0.1 * np.sin(np.arange(n) / 10)).astype(np.float32), sample_rate=24000, samples=n,
                                     segment_idx=0, token_count=1, audio_duration=0.1, processing_time_seconds=0.01,
                                     real_time_factor=10.0)
    loaded = []
    def fake_load(path, **kw):
        loaded.append(path)
        if os.environ.get("FAIL_LOAD"):
            raise ValueError("boom")
        return StubModel()
    U.load_model = fake_load
    import qwen3_tts.io as rio
    rio.time.sleep = lambda s: None
    from qwen3_tts.sessions import run_custom_session, run_design_session, run_clone_manager
    {call}
    print("RESULT" + json.dumps({{"calls": calls, "loaded": loaded}}))
''')


def _run(tmp_path, call, stdin, folder, env_extra=None):
    (tmp_path / "models" / folder).mkdir(parents=True)
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "qwen3-tts-apple-silicon_b200"), os.path.join(REF, "src"),
                                         os.path.join(ROOT, "tests", "stubs")])
    env.update(env_extra or {})
    p = subprocess.run([sys.executable, "-c", DRIVER.replace('{call}', call).replace('{{', '{').replace('}}', '}')], input=stdin, text=True, capture_output=True,
                       cwd=tmp_path, env=env, timeout=120)
    assert p.returncode == 0, p.stderr[-2000:]
    import json
    line = [l for l in p.stdout.splitlines() if "RESULT{" in l][-1]
    return json.loads(line[line.index("RESULT{") + len("RESULT"):]), p.stdout


def _wavs(tmp_path, sub):
    d = tmp_path / "outputs" / sub
    return sorted(d.rglob("*.wav")) if d.exists() else []


def test_design_session_kwargs_and_output(tmp_path):
    res, _ = _run(tmp_path, 'run_design_session("2")', "A calm narrator\nHello world this is a test\nq\n",
                  "Qwen3-TTS-12Hz-1.7B-VoiceDesign-8bit")
    assert res["loaded"] == [str(tmp_path / "models" / "Qwen3-TTS-12Hz-1.7B-VoiceDesign-8bit")]
    assert res["calls"] == [{"text": "Hello world this is a test", "instruct": "A calm narrator", "speed": 1.0,
                             "lang_code": "auto", "verbose": False}]
    wavs = _wavs(tmp_path, "VoiceDesign")
    assert len(wavs) == 1 and wavs[0].name.endswith("_Hello_world_this_is.wav")
    with wave.open(str(wavs[0])) as w:
        assert (w.getframerate(), w.getnchannels(), w.getsampwidth(), w.getnframes()) == (24000, 1, 2, 2400)


def test_custom_session_passes_lowercased_voice_and_speed(tmp_path):
    res, _ = _run(tmp_path, 'run_custom_session("1")', "7\n3\n2\nHello there\nq\n", "Qwen3-TTS-12Hz-1.7B-CustomVoice-8bit")
    (c,) = res["calls"]
    assert c["voice"] == "uncle_fu" and c["speed"] == 1.3 and c["text"] == "Hello there"
    assert c["instruct"] == "Excited and happy, speaking very fast"
    assert len(_wavs(tmp_path, "CustomVoice")) == 1


def test_clone_session_passes_ref_audio_path(tmp_path):
    (tmp_path / "voices").mkdir()
    with wave.open(str(tmp_path / "voices" / "Boss.wav"), "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(24000); w.writeframes(b"\\0\\0" * 2400)
    (tmp_path / "voices" / "Boss.txt").write_text("This is what the boss says.")
    res, _ = _run(tmp_path, 'run_clone_manager("3")', "1\n1\nClone me please\nq\n", "Qwen3-TTS-12Hz-1.7B-Base-8bit")
    (c,) = res["calls"]
    assert c["ref_audio"] == str(tmp_path / "voices" / "Boss.wav") and c["ref_text"] == "This is what the boss says."
    assert c["text"] == "Clone me please"
    assert len(_wavs(tmp_path, "Clones")) == 1


def test_load_failure_is_reported_not_raised(tmp_path):
    res, out = _run(tmp_path, 'run_design_session("2")', "x\ny\nq\n", "Qwen3-TTS-12Hz-1.7B-VoiceDesign-8bit",
                    {"FAIL_LOAD": "1"})
    assert res["calls"] == [] and "Failed to load model" in out      # reference io.py:115-117


def test_reference_smoke_test_passes_against_shim():
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "qwen3-tts-apple-silicon_b200"), os.path.join(REF, "src"),
                                         os.path.join(ROOT, "tests", "stubs")])
    p = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "-o", "addopts=",
                        os.path.join(REF, "tests", "test_sessions_smoke.py")], capture_output=True, text=True, env=env,
                       cwd="/tmp", timeout=300)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
