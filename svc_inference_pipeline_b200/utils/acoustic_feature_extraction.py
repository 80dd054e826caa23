"""Mel de-normalisation between the acoustic model and the vocoder
(reference ``utils/acoustic_feature_extraction.py:83-97``, called by ``infer.py:80``).

``denormalize_mel_channel(mel, cfg)`` keeps the reference's host-side call (numpy expression, same
operation order and dtype rules).  The B200 path can skip it: ``Generator.set_mel_denorm(mel_min,
mel_max)`` fuses the same affine map into the head kernel (``bvg_pack_mel``).

``load_mel_min_max(cfg)`` honours the reference's config contract (``utils/acoustic_feature_extraction.py:66-72``):
the two pickles named by ``cfg.min_mel_file`` / ``cfg.max_mel_file`` (``config/config.json:13-14``) are read when
the config carries them; then ``cfg.mel_range_path`` (an ``.npz`` with ``mel_min`` / ``mel_max``, this repo's
extension); only a config with neither (or no config) gets the copy of the reference's statistics shipped in
``config/mel_range.npz`` -- with a warning when a config was given, because a model trained with other
statistics would be de-normalised wrongly.
"""
from __future__ import annotations

import os
import pickle
import warnings

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_DEFAULT = os.path.join(os.path.dirname(_HERE), "config", "mel_range.npz")


def _cfg_get(cfg, key):
    if cfg is None:
        return None
    if isinstance(cfg, dict):
        return cfg.get(key)
    return getattr(cfg, key, None)


def load_mel_min_max(cfg=None):
    """Per-band ``(mel_min, mel_max)`` as the reference's ``load_mel_min_max`` returns them (whatever the pickles
    hold -- float32 arrays for the reference's own files, so the numpy expression below stays in float32)."""
    lo_file, hi_file = _cfg_get(cfg, "min_mel_file"), _cfg_get(cfg, "max_mel_file")
    if lo_file and hi_file:
        with open(lo_file, "rb") as f:
            mel_min = pickle.load(f)
        with open(hi_file, "rb") as f:
            mel_max = pickle.load(f)
        return np.asarray(mel_min), np.asarray(mel_max)
    path = _cfg_get(cfg, "mel_range_path")
    if path is None and cfg is not None:
        warnings.warn("load_mel_min_max: the config names neither min_mel_file / max_mel_file nor mel_range_path; "
                      "using the copy of the reference's statistics shipped in config/mel_range.npz", stacklevel=2)
    d = np.load(path or _DEFAULT)
    return d["mel_min"].astype(np.float32), d["mel_max"].astype(np.float32)


def denormalize_mel_channel(mel, cfg=None):
    """``mel`` [n_mels, T] in [-1, 1] -> log-mel in the vocoder's range, as a tensor on ``mel.device``."""
    device = mel.device
    mel = mel.cpu().numpy()
    ZERO = 1e-12
    mel_min, mel_max = load_mel_min_max(cfg)
    mel_min = np.expand_dims(mel_min, -1)
    mel_max = np.expand_dims(mel_max, -1)
    mel_norm = (mel + 1) / 2 * (mel_max - mel_min + ZERO) + mel_min
    return torch.as_tensor(mel_norm, device=device)
