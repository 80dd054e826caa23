"""CPU oracle for the BigVGAN vocoder forward path -- TEST INFRASTRUCTURE ONLY.

This file is a plain-numpy restatement of the algorithm the reference implements in
``modules/bigvgan.py`` (generator half, lines 1-632) and the waveform tail of
``modules/bigvgan_inference.py:29-44``.  It exists to check the CUDA path; it is never
imported by the product package (``svc_inference_pipeline_b200``).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg may import it.

Parity pin: the reference ships no tests or golden vectors for this path,
so the oracle is pinned against outputs of the reference module itself, imported and run
in the build container by ``tests/golden/make_golden.py``; the resulting vectors are committed
under ``tests/golden/`` and ``tests/test_oracle_golden.py`` checks every function here against
them.  The arithmetic itself lives in PyTorch (third-party; unpinned by the reference; the
goldens were produced with torch 2.11.0+cu128, CPU/oneDNN).

All functions compute in the dtype of their inputs (float32 or float64) and use the
reference's tensor layout ``[B, C, T]``.
"""
from __future__ import annotations

import math

import numpy as np

# --------------------------------------------------------------------------------------
# a1: kaiser-windowed sinc low-pass design            (reference modules/bigvgan.py:162-193)
# --------------------------------------------------------------------------------------


def kaiser_sinc_filter1d(cutoff: float, half_width: float, kernel_size: int, dtype=np.float32) -> np.ndarray:
    """12-tap (in practice) kaiser-sinc FIR, normalised to unit DC gain.  Returns ``[kernel_size]``.

    bigvgan.py:165-166 even/half_size; :169-176 kaiser beta from the attenuation estimate;
    :177 ``torch.kaiser_window(periodic=False)`` (float32); :180-183 time axis; :187-190 the
    windowed sinc and its normalisation.  The reference evaluates all of this in float32
    tensors; so does this function (``dtype`` selects the arithmetic type).
    """
    even = kernel_size % 2 == 0
    half_size = kernel_size // 2
    delta_f = 4 * half_width
    att = 2.285 * (half_size - 1) * math.pi * delta_f + 7.95
    if att > 50.0:
        beta = 0.1102 * (att - 8.7)
    elif att >= 21.0:
        beta = 0.5842 * (att - 21) ** 0.4 + 0.07886 * (att - 21.0)
    else:
        beta = 0.0
    n = np.arange(kernel_size, dtype=np.float64)
    # symmetric kaiser window: I0(beta * sqrt(1 - (2n/(N-1) - 1)^2)) / I0(beta)
    ratio = 2.0 * n / (kernel_size - 1) - 1.0
    window = np.i0(beta * np.sqrt(np.maximum(0.0, 1.0 - ratio * ratio))) / np.i0(beta)
    window = window.astype(dtype)
    if even:
        time = (np.arange(-half_size, half_size) + 0.5).astype(dtype)
    else:
        time = (np.arange(kernel_size) - half_size).astype(dtype)
    if cutoff == 0:
        return np.zeros_like(time)
    arg = (2 * cutoff * time).astype(dtype)
    filt = (dtype(2 * cutoff) * window * np.sinc(arg).astype(dtype)).astype(dtype)
    filt = filt / filt.sum(dtype=dtype)
    return filt.astype(dtype)


# --------------------------------------------------------------------------------------
# a2: Snake / SnakeBeta                                (reference modules/bigvgan.py:84-95, 146-159)
# --------------------------------------------------------------------------------------


def snake(x: np.ndarray, alpha: np.ndarray, beta: np.ndarray | None, logscale: bool) -> np.ndarray:
    """``x + 1/(b + 1e-9) * sin(a*x)^2`` with per-channel ``a, b`` on ``[B, C, T]``.

    Snake (bigvgan.py:90-93): ``a = b = alpha``; SnakeBeta (:152-157): ``a = alpha, b = beta``;
    ``logscale`` exponentiates both first (:91-92, :154-156).
    """
    dt = x.dtype
    a = alpha.astype(dt)
    b = a if beta is None else beta.astype(dt)
    if logscale:
        a = np.exp(a)
        b = np.exp(b)
    a = a[None, :, None]
    b = b[None, :, None]
    return x + (dt.type(1.0) / (b + dt.type(1e-9))) * np.sin(x * a) ** 2


# --------------------------------------------------------------------------------------
# a3 / a4: anti-aliasing resamplers                     (reference modules/bigvgan.py:196-307)
# --------------------------------------------------------------------------------------


def _replicate_pad(x: np.ndarray, left: int, right: int) -> np.ndarray:
    return np.pad(x, ((0, 0), (0, 0), (left, right)), mode="edge")


def upsample1d(x: np.ndarray, filt: np.ndarray, ratio: int = 2) -> np.ndarray:
    """``UpSample1d.forward`` (bigvgan.py:278-287), general ratio / kernel size.

    replicate-pad ``pad`` each side (:281), depthwise transposed conv with stride ``ratio``
    (:282-284), times ``ratio``, crop ``pad_left`` / ``pad_right`` (:285) with the constants of
    ``__init__`` (:262-271).
    """
    ksz = filt.shape[0]
    pad = ksz // ratio - 1
    pad_left = pad * ratio + (ksz - ratio) // 2
    pad_right = pad * ratio + (ksz - ratio + 1) // 2
    xp = _replicate_pad(x, pad, pad)
    bsz, ch, lp = xp.shape
    full = np.zeros((bsz, ch, (lp - 1) * ratio + ksz), dtype=x.dtype)
    f = filt.astype(x.dtype)
    for k in range(ksz):  # transposed conv == scatter-add of each tap
        full[:, :, k : k + lp * ratio : ratio] += xp * f[k]
    full *= x.dtype.type(ratio)
    return full[:, :, pad_left : full.shape[-1] - pad_right]


def lowpass_downsample1d(s: np.ndarray, filt: np.ndarray, ratio: int = 2) -> np.ndarray:
    """``DownSample1d`` -> ``LowPassFilter1d.forward`` (bigvgan.py:304-307, 224-231).

    replicate-pad ``K/2 - even`` left and ``K/2`` right (:215-216, :227), depthwise ``conv1d``
    (cross-correlation) with stride ``ratio`` (:229).
    """
    ksz = filt.shape[0]
    even = ksz % 2 == 0
    sp = _replicate_pad(s, ksz // 2 - int(even), ksz // 2)
    lout = (sp.shape[-1] - ksz) // ratio + 1
    out = np.zeros(s.shape[:2] + (lout,), dtype=s.dtype)
    f = filt.astype(s.dtype)
    for k in range(ksz):
        out += sp[:, :, k : k + (lout - 1) * ratio + 1 : ratio] * f[k]
    return out


def upsample2x_closed_form(x: np.ndarray, filt: np.ndarray) -> np.ndarray:
    """Polyphase closed form of ``UpSample1d`` for ratio 2 / 12 taps.

    ``u[2i] = 2*sum_t f[2t+1] x[clamp(i+2-t)]``, ``u[2i+1] = 2*sum_t f[2t] x[clamp(i+3-t)]``.
    This is the formulation the CUDA kernel uses; kept here so it is pinned by the same goldens.
    """
    assert filt.shape[0] == 12
    bsz, ch, ln = x.shape
    f = filt.astype(x.dtype)
    idx = np.arange(ln)
    out = np.zeros((bsz, ch, 2 * ln), dtype=x.dtype)
    ev = np.zeros_like(x)
    od = np.zeros_like(x)
    for t in range(6):
        ev += f[2 * t + 1] * x[:, :, np.clip(idx + 2 - t, 0, ln - 1)]
        od += f[2 * t] * x[:, :, np.clip(idx + 3 - t, 0, ln - 1)]
    out[:, :, 0::2] = 2 * ev
    out[:, :, 1::2] = 2 * od
    return out


def downsample2x_closed_form(s: np.ndarray, filt: np.ndarray) -> np.ndarray:
    """``z[i] = sum_k f[k] s[clamp(2i + k - 5, 0, 2L-1)]``."""
    assert filt.shape[0] == 12
    l2 = s.shape[-1]
    ln = l2 // 2
    f = filt.astype(s.dtype)
    idx = 2 * np.arange(ln)
    out = np.zeros(s.shape[:2] + (ln,), dtype=s.dtype)
    for k in range(12):
        out += f[k] * s[:, :, np.clip(idx + k - 5, 0, l2 - 1)]
    return out


def activation1d(x, alpha, beta, logscale, filt_up, filt_down) -> np.ndarray:
    """``Activation1d.forward`` (bigvgan.py:251-256): upsample x2 -> snake -> downsample x2."""
    u = upsample1d(x, filt_up, 2)
    u = snake(u, alpha, beta, logscale)
    return lowpass_downsample1d(u, filt_down, 2)


# --------------------------------------------------------------------------------------
# a6 / a10 / a12: weight-normed dense convolutions
# --------------------------------------------------------------------------------------


def weight_norm_fold(v: np.ndarray, g: np.ndarray) -> np.ndarray:
    """``w = g * v / ||v||`` with the norm over every dim except dim 0, no epsilon.

    ``torch.nn.utils.weight_norm`` default ``dim=0`` (used at bigvgan.py:319-386, 529, 550,
    593).  Dim 0 is Cout for ``Conv1d`` and **Cin** for ``ConvTranspose1d``.
    """
    norm = np.sqrt((v.astype(np.float64) ** 2).sum(axis=tuple(range(1, v.ndim)), keepdims=True)).astype(v.dtype)
    return v * (g.astype(v.dtype) / norm)


def get_padding(kernel_size: int, dilation: int = 1) -> int:
    """bigvgan.py:32-33."""
    return int((kernel_size * dilation - dilation) / 2)


def conv1d(x, w, b, dilation: int = 1, padding: int = 0) -> np.ndarray:
    """``nn.Conv1d`` forward, stride 1, zero padding.  ``w`` is ``[Cout, Cin, K]``."""
    bsz, cin, ln = x.shape
    cout, _, ksz = w.shape
    xp = np.pad(x, ((0, 0), (0, 0), (padding, padding)))
    lout = ln + 2 * padding - dilation * (ksz - 1)
    out = np.zeros((bsz, cout, lout), dtype=x.dtype)
    for k in range(ksz):
        out += np.matmul(w[:, :, k].astype(x.dtype), xp[:, :, k * dilation : k * dilation + lout])
    if b is not None:
        out += b.astype(x.dtype)[None, :, None]
    return out


def conv_transpose1d(x, w, b, stride: int, padding: int) -> np.ndarray:
    """``nn.ConvTranspose1d`` forward.  ``w`` is ``[Cin, Cout, K]`` (bigvgan.py:549-559)."""
    bsz, cin, ln = x.shape
    _, cout, ksz = w.shape
    full = np.zeros((bsz, cout, (ln - 1) * stride + ksz), dtype=x.dtype)
    for k in range(ksz):
        full[:, :, k : k + ln * stride : stride] += np.matmul(w[:, :, k].astype(x.dtype).T, x)
    out = full[:, :, padding : full.shape[-1] - padding]
    if b is not None:
        out = out + b.astype(x.dtype)[None, :, None]
    return out


# --------------------------------------------------------------------------------------
# a7 / a8 / a11: blocks and the generator, driven by a reference-format state_dict
# --------------------------------------------------------------------------------------


def _cfg_get(cfg, key):
    return cfg[key] if isinstance(cfg, dict) else getattr(cfg, key)


def _folded(sd, prefix):
    if prefix + ".weight_v" in sd:
        return weight_norm_fold(sd[prefix + ".weight_v"], sd[prefix + ".weight_g"])
    return sd[prefix + ".weight"]


def _act(sd, prefix, x, cfg):
    """One ``Activation1d`` from state_dict keys ``{prefix}.act.alpha|beta``, ``.upsample.filter``,
    ``.downsample.lowpass.filter`` (key grammar: SURVEY.md section 8b)."""
    name = _cfg_get(cfg, "activation")
    if name not in ("snake", "snakebeta"):
        raise NotImplementedError("activation incorrectly specified. check the config file and look for 'activation'.")
    alpha = sd[prefix + ".act.alpha"]
    beta = sd[prefix + ".act.beta"] if name == "snakebeta" else None
    fu = sd[prefix + ".upsample.filter"].reshape(-1)
    fd = sd[prefix + ".downsample.lowpass.filter"].reshape(-1)
    return activation1d(x, alpha, beta, bool(_cfg_get(cfg, "snake_logscale")), fu, fd)


def amp_block1(sd, prefix, x, cfg, kernel_size, dilations) -> np.ndarray:
    """``AMPBlock1.forward`` (bigvgan.py:424-433): per dilation ``x = c2(a2(c1(a1(x)))) + x``."""
    for l, d in enumerate(dilations):
        xt = _act(sd, f"{prefix}.activations.{2 * l}", x, cfg)
        xt = conv1d(xt, _folded(sd, f"{prefix}.convs1.{l}"), sd[f"{prefix}.convs1.{l}.bias"], d, get_padding(kernel_size, d))
        xt = _act(sd, f"{prefix}.activations.{2 * l + 1}", xt, cfg)
        xt = conv1d(xt, _folded(sd, f"{prefix}.convs2.{l}"), sd[f"{prefix}.convs2.{l}.bias"], 1, get_padding(kernel_size, 1))
        x = xt + x
    return x


def amp_block2(sd, prefix, x, cfg, kernel_size, dilations) -> np.ndarray:
    """``AMPBlock2.forward`` (bigvgan.py:506-512): per dilation ``x = c(a(x)) + x``."""
    for l, d in enumerate(dilations):
        xt = _act(sd, f"{prefix}.activations.{l}", x, cfg)
        xt = conv1d(xt, _folded(sd, f"{prefix}.convs.{l}"), sd[f"{prefix}.convs.{l}.bias"], d, get_padding(kernel_size, d))
        x = xt + x
    return x


def generator_forward(sd: dict, cfg, mel: np.ndarray) -> np.ndarray:
    """``Generator.forward`` (bigvgan.py:600-622).  ``mel`` ``[B, input_dim, T]`` -> ``[B, 1, T*prod(rates)]``.

    ``sd`` maps the reference's state_dict key names to numpy arrays; ``cfg`` is the ``vocoder``
    config block (dict or attribute object).
    """
    dt = mel.dtype
    sd = {k: np.asarray(v).astype(dt) for k, v in sd.items()}
    rates = list(_cfg_get(cfg, "upsample_rates"))
    ksizes = list(_cfg_get(cfg, "upsample_kernel_sizes"))
    rks = list(_cfg_get(cfg, "resblock_kernel_sizes"))
    rds = list(_cfg_get(cfg, "resblock_dilation_sizes"))
    block = amp_block1 if _cfg_get(cfg, "resblock") == "1" else amp_block2
    x = conv1d(mel, _folded(sd, "conv_pre"), sd["conv_pre.bias"], 1, 3)
    for i, (u, k) in enumerate(zip(rates, ksizes)):
        x = conv_transpose1d(x, _folded(sd, f"ups.{i}.0"), sd[f"ups.{i}.0.bias"], u, (k - u) // 2)
        xs = None
        for j, (rk, rd) in enumerate(zip(rks, rds)):
            y = block(sd, f"resblocks.{i * len(rks) + j}", x, cfg, rk, rd)
            xs = y if xs is None else xs + y
        x = xs / dt.type(len(rks))
    x = _act(sd, "activation_post", x, cfg)
    x = conv1d(x, _folded(sd, "conv_post"), sd["conv_post.bias"], 1, 3)
    return np.tanh(x)


# --------------------------------------------------------------------------------------
# a14: waveform tail of synthesis_audios          (reference modules/bigvgan_inference.py:33-44)
# --------------------------------------------------------------------------------------


def synthesis_tail(audio: np.ndarray, frames: int, hop_length: int) -> np.ndarray:
    """Trim to ``frames*hop`` samples and fade the last ``20*hop`` samples linearly to zero."""
    audio = np.array(audio[: frames * hop_length], dtype=np.float32, copy=True)
    fade = np.linspace(1.0, 0.0, 20 * hop_length, dtype=np.float32)
    audio[-20 * hop_length :] *= fade  # raises on frames < 20 exactly like the reference's broadcast
    return audio


# --------------------------------------------------------------------------------------
# section 8f row 1: the host steps either side of the vocoder in infer.py:80-90
# --------------------------------------------------------------------------------------


def denormalize_mel_channel(mel: np.ndarray, mel_min: np.ndarray, mel_max: np.ndarray) -> np.ndarray:
    """``denormalize_mel_channel`` (utils/acoustic_feature_extraction.py:83-97): ``mel`` [n_mels, T] (or
    [B, n_mels, T]) in [-1, 1] -> ``(mel + 1) / 2 * (mel_max - mel_min + 1e-12) + mel_min``, evaluated
    in float32 with one rounding per operation (numpy semantics for float32 arrays and python scalars)."""
    f = np.float32
    mel = mel.astype(f)
    mn = mel_min.astype(f)[..., :, None]
    rng = (mel_max.astype(f) - mel_min.astype(f) + f(1e-12)).astype(f)[..., :, None]
    return ((((mel + f(1.0)) * f(0.5)).astype(f) * rng).astype(f) + mn).astype(f)


def linspace_1_0(n: int) -> np.ndarray:
    """``torch.linspace(1, 0, steps=n)`` in float32, scalar formulation of ATen's CPU kernel
    (RangeFactoriesKernel: ``step = (end - start) / (n - 1)``; first half ``start + step * i``, second
    half ``end - step * (n - 1 - i)``).  ATen's *vectorised* path forms ``base + step * lane`` per SIMD
    chunk, so ``torch.linspace`` itself differs from this by at most one ulp on some entries, depending
    on the host's vector width (tests/test_oracle_golden.py::test_linspace_restates_torch)."""
    f = np.float32
    if n == 1:
        return np.ones(1, dtype=f)
    step = f(f(-1.0) / f(n - 1))
    i = np.arange(n)
    up = (f(1.0) + (step * i.astype(f)).astype(f)).astype(f)
    dn = (f(0.0) - (step * (n - 1 - i).astype(f)).astype(f)).astype(f)
    return np.where(i < n // 2, up, dn).astype(f)


def synthesis_pcm16(audio: np.ndarray, hop_length: int, fs: int, add_silence=True, turn_up=True, volume_peak=0.9) -> np.ndarray:
    """``synthesis_audios`` fade-out (modules/bigvgan_inference.py:37-42) followed by ``save_audio``'s
    waveform processing (utils/util.py:20-37) and 16-bit quantisation ``clip(rint(v * 32768))``, on one
    float32 waveform ``[T * hop]``.  ``volume_peak / peak`` is a float32 division (numpy >= 2 scalar
    rules).  The quantiser of ``torchaudio.save(..., encoding="PCM_S", bits_per_sample=16)`` is not
    importable in this image (needs torchcodec): stated, not pinned."""
    f = np.float32
    w = audio.astype(f).copy()
    n = 20 * hop_length
    w[-n:] = (w[-n:] * linspace_1_0(n)).astype(f)
    if turn_up:
        peak = f(max(w.max(), abs(w.min())))
        w = (w * f(f(volume_peak) / peak)).astype(f) if peak > 0 else np.zeros_like(w)
    if add_silence:
        s = np.zeros(fs // 20, dtype=f)
        w = np.concatenate([s, w, s])
    return np.clip(np.rint((w * f(32768.0)).astype(f)), -32768, 32767).astype(np.int16)
