"""Mel de-normalisation between the acoustic model and the vocoder
(reference ``utils/acoustic_feature_extraction.py:83-97``, called by ``infer.py:80``).

``denormalize_mel_channel(mel, cfg)`` keeps the reference's host-side call (numpy expression, same
operation order and dtype rules).  The B200 path can skip it: ``Generator.set_mel_denorm(mel_min,
mel_max)`` fuses the same affine map into the head kernel (``bvg_pack_mel``).  The reference reads
``mel_min`` / ``mel_max`` from two pickles named in the config; here they come from
``cfg.mel_min_max_stats_dir`` (``mel_min.npy`` / ``mel_max.npy`` or a ``mel_range.npz``) or, by
default, the copy of the reference's statistics shipped in ``config/mel_range.npz``.
"""
from __future__ import annotations

import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_DEFAULT = os.path.join(os.path.dirname(_HERE), "config", "mel_range.npz")


def load_mel_min_max(cfg=None):
    path = getattr(cfg, "mel_range_path", None) if cfg is not None else None
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
