"""Vocoder inference glue with the reference's call surface
(reference ``modules/bigvgan_inference.py:19-44``).

``synthesis_audios(model, mel, cfg)`` is what the reference's ``infer.py:86`` calls: mel
``[n_mels, T]`` in, ``np.float32[T * hop_length]`` out, trimmed and with the last 20 hops faded
linearly to zero.  ``vocoder_inference`` is the batched inner call (``[B, n_mels, T]`` ->
CPU ``[B, T*hop]``).  ``f0s``, ``batch_size`` and ``fast_inference`` are accepted and ignored
exactly like the reference does.
"""
from __future__ import annotations

import torch


def vocoder_inference(cfg, model, mels, device, fast_inference=False):
    model.eval()
    with torch.no_grad():
        mels = mels.to(device)
        output = model.forward(mels)
    return output.squeeze(1).detach().cpu()


def synthesis_audios(model, mel, cfg, f0s=None, batch_size=None, fast_inference=False):
    device = next(model.parameters()).device
    frame = mel.shape[-1]
    audio = vocoder_inference(cfg, model, mel.unsqueeze(0), device, fast_inference).squeeze(0)
    fade_out = torch.linspace(1, 0, steps=20 * cfg.hop_length)
    audio_length = frame * cfg.hop_length
    audio = audio[:audio_length]
    audio[-20 * cfg.hop_length :] *= fade_out  # raises for fewer than 20 frames, as the reference does
    return audio.numpy()
