"""Config objects for the vocoder path.

Mirrors the reference's config surface (reference ``utils/util.py:57-122``):
``load_config(path) -> JsonHParams`` where the file is JSON5-flavoured
(``//`` comments, trailing commas; reference ``config/config.json:3,37``) and may
inherit from a ``basic_config`` resolved against ``$WORD_DIR``
(reference ``utils/util.py:72-76``).

The reference parses with the third-party ``json5`` package, which is not a
dependency here: the reader below handles the JSON5 subset the reference's config
actually uses (line comments, block comments, trailing commas) with the stdlib
``json`` module.
"""
from __future__ import annotations

import json
import os

__all__ = ["JsonHParams", "load_config", "loads_json5_subset", "override_config", "save_audio"]


def _strip_json5(text: str) -> str:
    """Remove ``//`` and ``/* */`` comments and trailing commas outside of strings."""
    out = []
    i, n = 0, len(text)
    in_str = False
    quote = ""
    while i < n:
        ch = text[i]
        if in_str:
            out.append(ch)
            if ch == "\\" and i + 1 < n:
                out.append(text[i + 1])
                i += 2
                continue
            if ch == quote:
                in_str = False
            i += 1
            continue
        if ch in "\"'":
            in_str, quote = True, ch
            out.append(ch)
            i += 1
        elif ch == "/" and i + 1 < n and text[i + 1] == "/":
            while i < n and text[i] != "\n":
                i += 1
        elif ch == "/" and i + 1 < n and text[i + 1] == "*":
            end = text.find("*/", i + 2)
            i = n if end < 0 else end + 2
        else:
            out.append(ch)
            i += 1
    text = "".join(out)
    # trailing commas: a comma followed only by whitespace then } or ]
    res = []
    i, n = 0, len(text)
    in_str = False
    while i < n:
        ch = text[i]
        if in_str:
            res.append(ch)
            if ch == "\\" and i + 1 < n:
                res.append(text[i + 1])
                i += 2
                continue
            if ch == quote:
                in_str = False
            i += 1
            continue
        if ch in "\"'":
            in_str, quote = True, ch
            res.append(ch)
        elif ch == ",":
            j = i + 1
            while j < n and text[j] in " \t\r\n":
                j += 1
            if j < n and text[j] in "}]":
                i += 1
                continue
            res.append(ch)
        else:
            res.append(ch)
        i += 1
    return "".join(res)


def loads_json5_subset(text: str):
    return json.loads(_strip_json5(text))


def override_config(base: dict, new: dict) -> dict:
    """Recursive dict merge, ``new`` wins (reference ``utils/util.py:57-65``)."""
    for key, val in new.items():
        if isinstance(val, dict):
            base[key] = override_config(base.get(key, {}) if isinstance(base.get(key), dict) else {}, val)
        else:
            base[key] = val
    return base


def _load_config_dict(path: str) -> dict:
    with open(path, "r") as f:
        cfg = loads_json5_subset(f.read())
    if "basic_config" in cfg:
        root = os.getenv("WORD_DIR")  # (sic) the reference's env var name, util.py:73
        if root is None:
            raise KeyError("config has 'basic_config' but $WORD_DIR is not set")
        parent = _load_config_dict(os.path.join(root, cfg["basic_config"]))
        cfg = override_config(parent, cfg)
    return cfg


class JsonHParams:
    """Attribute-access view of a nested dict (reference ``utils/util.py:92-122``)."""

    def __init__(self, **kwargs):
        for key, val in kwargs.items():
            if type(val) is dict:
                val = JsonHParams(**val)
            self[key] = val

    def keys(self):
        return self.__dict__.keys()

    def items(self):
        return self.__dict__.items()

    def values(self):
        return self.__dict__.values()

    def __len__(self):
        return len(self.__dict__)

    def __getitem__(self, key):
        return getattr(self, key)

    def __setitem__(self, key, value):
        setattr(self, key, value)

    def __contains__(self, key):
        return key in self.__dict__

    def __repr__(self):
        return repr(self.__dict__)


def load_config(config_fn: str) -> JsonHParams:
    return JsonHParams(**_load_config_dict(config_fn))


def save_audio(path, waveform, fs, add_silence=True, turn_up=True, volume_peak=0.9):
    """Reference ``utils/util.py:20-37``: peak-normalise to ``volume_peak``, pad ``fs // 20`` samples of
    silence on each side, write 16-bit PCM.  ``waveform`` is either the float array
    ``synthesis_audios`` returns (processed here on the host like the reference does) or the
    ``int16`` array ``synthesis_pcm16`` returns (already normalised, padded and quantised on the GPU:
    written as is).  The reference encodes with ``torchaudio.save``; this writes the same container
    with the stdlib ``wave`` module (no torchcodec / sox dependency)."""
    import wave

    import numpy as np

    w = np.asarray(waveform)
    if w.dtype != np.int16:
        w = w.astype(np.float32)
        if turn_up:
            ratio = volume_peak / max(w.max(), abs(w.min()))
            w = w * np.float32(ratio)
        if add_silence:
            silence = np.zeros((fs // 20,), dtype=w.dtype)
            w = np.concatenate([silence, w, silence])
        w = np.clip(np.rint(w * np.float32(32768.0)), -32768, 32767).astype(np.int16)
    with wave.open(path, "wb") as f:
        f.setnchannels(1)
        f.setsampwidth(2)
        f.setframerate(int(fs))
        f.writeframes(np.ascontiguousarray(w.reshape(-1)).tobytes())
