"""Vocoder inference glue with the reference's call surface
(reference ``modules/bigvgan_inference.py:19-44``).

``synthesis_audios(model, mel, cfg)`` is what the reference's ``infer.py:86`` calls: mel
``[n_mels, T]`` in, ``np.float32[T * hop_length]`` out, trimmed and with the last 20 hops faded
linearly to zero.  ``vocoder_inference`` is the batched inner call (``[B, n_mels, T]`` ->
CPU ``[B, T*hop]``).  ``f0s``, ``batch_size`` and ``fast_inference`` are accepted and ignored
exactly like the reference does.

``synthesis_pcm16`` is the fused form of what ``infer.py:86-90`` does next on the host
(``synthesis_audios`` fade-out + ``save_audio``'s peak normalisation, silence padding and 16-bit
quantisation, reference ``utils/util.py:20-37``): the waveform never leaves the GPU as fp32, the
device -> host copy is int16.
"""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib as L


def vocoder_inference(cfg, model, mels, device, fast_inference=False):
    model.eval()
    with torch.no_grad():
        mels = mels.to(device)
        # the B200 Generator hands out its own output buffer here (no device-side copy): .cpu() below is the
        # consumer and is synchronous; any other model goes through forward() like in the reference
        output = getattr(model, "forward_borrowed", model.forward)(mels)
    output = output.squeeze(1).detach()
    if not output.is_cuda:
        return output.cpu()
    # device -> host into page-locked memory from PyTorch's caching host allocator (a DMA at PCIe rate instead of a
    # staged pageable copy: 15 MB per B16 x 10 s batch); the result is an ordinary CPU tensor owned by the caller
    host = torch.empty(output.shape, dtype=output.dtype, pin_memory=True)
    host.copy_(output, non_blocking=True)
    torch.cuda.current_stream(output.device).synchronize()
    return host


def synthesis_audios(model, mel, cfg, f0s=None, batch_size=None, fast_inference=False):
    device = next(model.parameters()).device
    frame = mel.shape[-1]
    audio = vocoder_inference(cfg, model, mel.unsqueeze(0), device, fast_inference).squeeze(0)
    fade_out = torch.linspace(1, 0, steps=20 * cfg.hop_length)
    audio_length = frame * cfg.hop_length
    audio = audio[:audio_length]
    audio[-20 * cfg.hop_length :] *= fade_out  # raises for fewer than 20 frames, as the reference does
    return audio.numpy()


def synthesis_pcm16(model, mel, cfg, add_silence=True, turn_up=True, volume_peak=0.9):
    """``mel [n_mels, T]`` (or a batch ``[B, n_mels, T]``) -> ``np.int16`` PCM ``[T*hop (+ 2 * fs // 20)]``
    (``[B, ...]`` for a batch): generator forward, fade-out of the last 20 hops, peak normalisation
    to ``volume_peak``, ``fs // 20`` samples of silence on each side, 16-bit quantisation -- the
    arguments are ``save_audio``'s.  Every item of a batch is normalised by its own peak."""
    device = next(model.parameters()).device
    single = mel.dim() == 2
    mels = mel.unsqueeze(0) if single else mel
    B, _, frames = mels.shape
    model.eval()
    with torch.no_grad():
        wave = getattr(model, "forward_borrowed", model.forward)(mels.to(device))  # [B, 1, T*hop] fp32 on the device
    Ln = frames * cfg.hop_length
    fade = 20 * cfg.hop_length
    if fade > Ln:
        raise RuntimeError(f"synthesis needs at least 20 mel frames (got {frames}): the fade-out spans 20 hops")
    silence = (cfg.fs // 20) if add_silence else 0
    pcm = torch.empty(B, Ln + 2 * silence, dtype=torch.int16, device=device)
    peak = torch.empty(B, dtype=torch.float32, device=device)
    d = L.TailDesc()
    d.d_wave, d.d_pcm, d.d_peak = wave.data_ptr(), pcm.data_ptr(), peak.data_ptr()
    d.B, d.L, d.fade_len, d.silence = B, Ln, fade, silence
    d.volume_peak = float(volume_peak) if turn_up else 0.0
    with torch.cuda.device(device):
        L.check(L.lib().bvg_tail_fwd(C.byref(d), torch.cuda.current_stream(device).cuda_stream), "tail_fwd")
    out = pcm.cpu().numpy()
    return out[0] if single else out
