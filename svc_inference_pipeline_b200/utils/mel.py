"""Log-mel analysis with the reference's call surface (reference ``utils/mel.py:130-174``), on the GPU.

``mel_spectrogram(y, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, center=False)`` keeps the
reference's signature and semantics (reflect pad ``(n_fft - hop) / 2``, periodic hann window, ``center=False``
framing, ``sqrt(re^2 + im^2 + 1e-9)``, mel basis, ``log(clamp(., 1e-5))``) and returns ``[B, num_mels, frames]`` on
``y``'s device.  The arithmetic is one CUDA kernel (``bvg_logmel_fwd``: two frames per complex radix-2 FFT in
shared memory); there is no CPU path -- ``y`` must live on a B200.

The reference takes its mel basis from ``librosa.filters.mel`` (``utils/mel.py:14,140``), a dependency it does not
pin and this image does not have; ``mel_filterbank`` restates librosa's published default (Slaney scale, Slaney
area normalisation) and is checked against an independent implementation (``tests/golden/make_golden.py``).
Used by the parity metrics (log-mel L1 of the bf16 gate, ``bench.py``) and for analysis -> synthesis round trips.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _lib as L

__all__ = ["mel_filterbank", "mel_spectrogram", "log_mel_l1", "dynamic_range_compression_torch", "spectral_normalize_torch"]

_F_SP, _MIN_LOG_HZ = 200.0 / 3, 1000.0
_MIN_LOG_MEL, _LOGSTEP = _MIN_LOG_HZ / _F_SP, np.log(6.4) / 27.0


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    return np.where(f >= _MIN_LOG_HZ, _MIN_LOG_MEL + np.log(np.maximum(f, _MIN_LOG_HZ) / _MIN_LOG_HZ) / _LOGSTEP, f / _F_SP)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    return np.where(m >= _MIN_LOG_MEL, _MIN_LOG_HZ * np.exp(_LOGSTEP * (m - _MIN_LOG_MEL)), _F_SP * m)


def mel_filterbank(sr, n_fft, n_mels, fmin=0.0, fmax=None) -> np.ndarray:
    """``librosa.filters.mel(sr=, n_fft=, n_mels=, fmin=, fmax=)`` with librosa's defaults (``htk=False``,
    ``norm="slaney"``): float32 ``[n_mels, 1 + n_fft // 2]`` (the call at reference ``utils/mel.py:140``)."""
    fmax = sr / 2.0 if fmax is None else fmax
    n_bins = 1 + n_fft // 2
    fft_f = np.linspace(0.0, sr / 2.0, n_bins)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fft_f[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    w = np.maximum(0.0, np.minimum(lower, upper))
    w *= (2.0 / (mel_f[2 : n_mels + 2] - mel_f[:n_mels]))[:, None]
    return w.astype(np.float32)


_basis_cache = {}


def _device_basis(key, device):
    k = key + (str(device),)
    hit = _basis_cache.get(k)
    if hit is None:
        sr, n_fft, n_mels, fmin, fmax = key
        fb = mel_filterbank(sr, n_fft, n_mels, fmin, fmax)
        nz = fb != 0
        first = np.where(nz.any(axis=1), nz.argmax(axis=1), 0)
        last = np.where(nz.any(axis=1), fb.shape[1] - nz[:, ::-1].argmax(axis=1), 0)
        band = np.stack([first, last], axis=1).astype(np.int32)
        hit = (torch.from_numpy(fb).to(device), torch.from_numpy(band).to(device))
        _basis_cache[k] = hit
    return hit


def dynamic_range_compression_torch(x, C=1, clip_val=1e-5):
    return torch.log(torch.clamp(x, min=clip_val) * C)


def spectral_normalize_torch(magnitudes):
    return dynamic_range_compression_torch(magnitudes)


@torch.no_grad()
def mel_spectrogram(y, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, center=False):
    """Reference ``utils/mel.py:130-174``: ``y [B, n]`` (or ``[n]``) float waveform on the GPU -> ``[B, num_mels, frames]``."""
    if center:
        raise NotImplementedError("mel_spectrogram: the reference path runs with center=False (utils/mel.py:170)")
    if y.device.type != "cuda":
        raise RuntimeError("svc_inference_pipeline_b200.utils.mel has no CPU path: move the waveform to the GPU")
    if y.dim() == 1:
        y = y.unsqueeze(0)
    y = y.contiguous().float()
    B, n = y.shape
    pad = int((n_fft - hop_size) / 2)
    frames = 1 + (n + 2 * pad - n_fft) // hop_size
    basis, band = _device_basis((int(sampling_rate), int(n_fft), int(num_mels), float(fmin), None if fmax is None else float(fmax)), y.device)
    out = torch.empty(B, num_mels, max(frames, 0), dtype=torch.float32, device=y.device)
    d = L.LogmelDesc()
    d.d_wave, d.wave_stride, d.d_out = y.data_ptr(), y.stride(0), out.data_ptr()
    d.d_basis, d.d_band = basis.data_ptr(), band.data_ptr()
    d.B, d.n, d.n_fft, d.hop, d.win, d.n_mels, d.frames = B, n, int(n_fft), int(hop_size), int(win_size), int(num_mels), frames
    d.clip = 1e-5
    with torch.cuda.device(y.device):
        L.require_sm100(y.device.index if y.device.index is not None else torch.cuda.current_device())
        L.check(L.lib().bvg_logmel_fwd(C.byref(d), torch.cuda.current_stream(y.device).cuda_stream), "logmel_fwd")
    return out


def log_mel_l1(ref_wave, wave, n_fft=1024, num_mels=100, sampling_rate=24000, hop_size=256, win_size=1024, fmin=0, fmax=12000) -> float:
    """Mean absolute difference of the two waveforms' log-mels under the reference's analysis: the "log-mel L1" of
    the bf16 parity gate (BASELINE.json north_star), computed on the device."""
    a = mel_spectrogram(ref_wave.reshape(1, -1), n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax)
    b = mel_spectrogram(wave.reshape(1, -1), n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax)
    return float((a - b).abs().mean())
