"""CPU oracle for the log-mel analysis of the reference (TEST INFRASTRUCTURE ONLY -- never imported by
the product path; see tests/test_host_logic.py::test_no_product_import_of_oracle).

Restates reference ``utils/mel.py:130-174`` (``mel_spectrogram``: reflect pad, hann window, magnitude
STFT with ``center=False``, mel projection, ``log(clamp(., 1e-5))``) in numpy, float64 or float32.

The mel basis itself lives in a third-party dependency that is absent from ``/root/reference`` and from
this image: ``librosa.filters.mel`` (``utils/mel.py:14,140``; the reference pins no version -- it ships no
requirements file).  Its published algorithm (librosa >= 0.6 defaults ``htk=False, norm="slaney"``) is
restated in ``slaney_mel_filterbank`` below:

* Slaney's Auditory-Toolbox mel scale: linear (200/3 Hz per mel) below 1 kHz, logarithmic above with
  ``logstep = ln(6.4) / 27``;
* ``n_mels + 2`` band edges equally spaced on that scale between ``fmin`` and ``fmax``;
* triangular weights ``max(0, min((f - lo) / (ce - lo), (hi - f) / (hi - ce)))`` on the FFT bin centres;
* "slaney" area normalisation: each triangle times ``2 / (hi - lo)`` (Hz).

Pin: ``tests/golden/make_golden.py`` checks this restatement against an independent implementation of the
same published algorithm that *is* in the image (``transformers.audio_utils.mel_filter_bank(norm="slaney",
mel_scale="slaney")``) and then runs the UNMODIFIED reference ``utils.mel.mel_spectrogram`` with this
function supplied as ``librosa.filters.mel``; the result is committed as ``tests/golden/logmel.npz``.
"""
from __future__ import annotations

import numpy as np

__all__ = ["hz_to_mel_slaney", "mel_to_hz_slaney", "slaney_mel_filterbank", "hann_periodic", "mel_spectrogram", "log_mel_l1"]

_F_SP = 200.0 / 3
_MIN_LOG_HZ = 1000.0
_MIN_LOG_MEL = _MIN_LOG_HZ / _F_SP
_LOGSTEP = np.log(6.4) / 27.0


def hz_to_mel_slaney(f):
    f = np.asarray(f, dtype=np.float64)
    lin = f / _F_SP
    log = _MIN_LOG_MEL + np.log(np.maximum(f, _MIN_LOG_HZ) / _MIN_LOG_HZ) / _LOGSTEP
    return np.where(f >= _MIN_LOG_HZ, log, lin)


def mel_to_hz_slaney(m):
    m = np.asarray(m, dtype=np.float64)
    lin = _F_SP * m
    log = _MIN_LOG_HZ * np.exp(_LOGSTEP * (m - _MIN_LOG_MEL))
    return np.where(m >= _MIN_LOG_MEL, log, lin)


def slaney_mel_filterbank(sr, n_fft, n_mels=128, fmin=0.0, fmax=None, dtype=np.float32):
    """``librosa.filters.mel(sr=, n_fft=, n_mels=, fmin=, fmax=)`` with its defaults (slaney scale, slaney
    norm): ``[n_mels, 1 + n_fft // 2]``.  Called by the reference at ``utils/mel.py:140``."""
    if fmax is None:
        fmax = sr / 2.0
    n_bins = 1 + n_fft // 2
    fftfreqs = np.linspace(0.0, sr / 2.0, n_bins)
    mel_f = mel_to_hz_slaney(np.linspace(hz_to_mel_slaney(fmin), hz_to_mel_slaney(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    weights = np.zeros((n_mels, n_bins), dtype=np.float64)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2 : n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, None]
    return weights.astype(dtype)


def hann_periodic(n, dtype=np.float64):
    """``torch.hann_window(n)`` (periodic=True): ``0.5 - 0.5 cos(2 pi k / n)`` (``utils/mel.py:146``)."""
    k = np.arange(n, dtype=np.float64)
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * k / n)).astype(dtype)


def mel_spectrogram(y, n_fft=1024, num_mels=100, sampling_rate=24000, hop_size=256, win_size=1024, fmin=0, fmax=12000, dtype=np.float64, basis=None):
    """Reference ``utils/mel.py:130-174`` for ``y[B, n]`` (``center=False``): returns ``[B, num_mels, frames]``
    with ``frames = 1 + (n + 2 * pad - n_fft) // hop_size``, ``pad = (n_fft - hop_size) // 2``."""
    y = np.asarray(y, dtype=dtype)
    if y.ndim == 1:
        y = y[None]
    pad = int((n_fft - hop_size) / 2)
    y = np.pad(y, ((0, 0), (pad, pad)), mode="reflect")                     # :148-153
    if basis is None:
        basis = slaney_mel_filterbank(sampling_rate, n_fft, num_mels, fmin, fmax)  # :140-142 (float32)
    basis = np.asarray(basis, dtype=dtype)
    win = hann_periodic(win_size, dtype)                                      # :146
    if win_size < n_fft:  # torch.stft centres a shorter window inside n_fft
        left = (n_fft - win_size) // 2
        win = np.pad(win, (left, n_fft - win_size - left))
    frames = 1 + (y.shape[1] - n_fft) // hop_size
    idx = np.arange(n_fft)[None, :] + hop_size * np.arange(frames)[:, None]
    seg = y[:, idx] * win[None, None, :]                                      # [B, frames, n_fft]
    spec = np.fft.rfft(seg.astype(np.float64), axis=-1)                       # :156-167
    mag = np.sqrt(spec.real**2 + spec.imag**2 + 1e-9).astype(dtype)           # :169
    mel = np.einsum("mf,btf->bmt", basis, mag)                                # :171
    return np.log(np.maximum(mel, dtype(1e-5) if dtype is np.float32 else 1e-5))  # :172, :25-26


def log_mel_l1(ref_wave, wave, **kw):
    """Mean absolute difference of the reference-analysis log-mels of two waveforms (the north_star's
    "log-mel L1" of the bf16 gate)."""
    a = mel_spectrogram(np.asarray(ref_wave, np.float64).reshape(1, -1), **kw)
    b = mel_spectrogram(np.asarray(wave, np.float64).reshape(1, -1), **kw)
    return float(np.abs(a - b).mean())
