"""Minimal stand-in for prompt_toolkit (absent offline) so the UNMODIFIED reference sessions import; the reference
falls back to rich.Prompt on a non-tty stdin (ui.py:44-48, 67-79), so none of this is exercised."""


def prompt(*a, **k):
    return input()
