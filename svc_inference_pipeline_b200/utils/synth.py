"""Deterministic synthetic checkpoints and mel inputs for the vocoder path.

There is no network and the reference's checkpoints (``config/config.json:8-10``) are not in
its tree, so benchmarks and parity tests run on random-init weights of the exact reference
architecture.  Everything here is numpy-only and bit-reproducible across machines (uniform
doubles from PCG64 + an explicit Box-Muller), so the GPU box regenerates the very state_dict
the committed golden vectors were produced with.

Recipe (after SURVEY.md section 8d): ``weight_v, bias ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in))`` (PyTorch's
default conv init, which is what the reference's weight-normed layers keep: ``init_weights``
at ``modules/bigvgan.py:36-39`` only touches the derived ``.weight``), ``weight_g = ||v|| * 2^U(-.5,.5)``
and ``alpha, beta ~ N(0, 0.25^2)`` so that the weight-norm fold and the per-channel snake
parameters are actually exercised (the defaults ``g = ||v||``, ``alpha = beta = 0`` would not).
The spreads are half (in log scale) of SURVEY's suggestion: with the full spreads on *both* the
116-conv random net becomes error-amplifying (fp32-vs-fp64 self-noise 1.3e-5 instead of 2e-7..8e-7,
bf16-operand SNR 31 dB instead of 41 dB; tools/precision_probe.py, DESIGN.md section "precision"), i.e. no
longer comparable with the default-init network the parity tolerances were calibrated on.
"""
from __future__ import annotations

import math
import os
from collections import OrderedDict

import numpy as np

__all__ = [
    "state_dict_spec",
    "synthetic_state_dict",
    "synthetic_mel",
    "mel_range",
    "count_parameters",
    "aa_filter_taps",
    "diffsvc_state_dict_spec",
    "synthetic_diffsvc_state_dict",
]


def _get(cfg, key):
    return cfg[key] if isinstance(cfg, dict) else getattr(cfg, key)


def aa_filter_taps(dtype=np.float32) -> np.ndarray:
    """The 12-tap anti-aliasing FIR both resamplers of every ``Activation1d`` carry
    (reference ``modules/bigvgan.py:162-193`` with cutoff 0.25, half-width 0.3, K=12:
    kaiser beta 4.6638, ``time = arange(-6, 6) + 0.5``), evaluated in float32 like the reference."""
    ksz, cutoff, half_width = 12, 0.25, 0.3
    att = 2.285 * (ksz // 2 - 1) * math.pi * (4 * half_width) + 7.95
    beta = 0.1102 * (att - 8.7) if att > 50.0 else (0.5842 * (att - 21) ** 0.4 + 0.07886 * (att - 21.0) if att >= 21.0 else 0.0)
    n = np.arange(ksz, dtype=np.float64)
    r = 2.0 * n / (ksz - 1) - 1.0
    window = (np.i0(beta * np.sqrt(np.maximum(0.0, 1.0 - r * r))) / np.i0(beta)).astype(np.float32)
    time = (np.arange(-ksz // 2, ksz // 2) + 0.5).astype(np.float32)
    filt = np.float32(2 * cutoff) * window * np.sinc((2 * cutoff * time).astype(np.float32)).astype(np.float32)
    filt = (filt / filt.sum(dtype=np.float32)).astype(np.float32)
    return filt.astype(dtype)


def state_dict_spec(vcfg) -> "OrderedDict[str, tuple]":
    """``name -> (shape, kind)`` for every tensor in the reference generator's ``state_dict()``
    (key grammar: SURVEY.md section 8b; construction: reference ``modules/bigvgan.py:521-598``).

    ``kind`` is one of ``bias``, ``g``, ``v_conv`` (``[Cout, Cin, K]``), ``v_convT`` (``[Cin, Cout, K]``),
    ``alpha``, ``beta``, ``filter``.
    """
    spec: "OrderedDict[str, tuple]" = OrderedDict()
    act = _get(vcfg, "activation")
    if act not in ("snake", "snakebeta"):
        raise NotImplementedError("activation incorrectly specified. check the config file and look for 'activation'.")
    c0 = int(_get(vcfg, "upsample_initial_channel"))
    cin0 = int(_get(vcfg, "input_dim"))
    rates = list(_get(vcfg, "upsample_rates"))
    uks = list(_get(vcfg, "upsample_kernel_sizes"))
    rks = list(_get(vcfg, "resblock_kernel_sizes"))
    rds = list(_get(vcfg, "resblock_dilation_sizes"))
    block1 = _get(vcfg, "resblock") == "1"

    def wn(prefix, cout, cin, k, transposed=False):
        spec[prefix + ".bias"] = ((cout,), "bias")
        dim0 = cin if transposed else cout
        spec[prefix + ".weight_g"] = ((dim0, 1, 1), "g")
        spec[prefix + ".weight_v"] = ((cin, cout, k) if transposed else (cout, cin, k), "v_convT" if transposed else "v_conv")

    def act1d(prefix, ch):
        spec[prefix + ".act.alpha"] = ((ch,), "alpha")
        if act == "snakebeta":
            spec[prefix + ".act.beta"] = ((ch,), "beta")
        spec[prefix + ".upsample.filter"] = ((1, 1, 12), "filter")
        spec[prefix + ".downsample.lowpass.filter"] = ((1, 1, 12), "filter")

    wn("conv_pre", c0, cin0, 7)
    for i, (u, k) in enumerate(zip(rates, uks)):
        wn(f"ups.{i}.0", c0 // 2 ** (i + 1), c0 // 2**i, k, transposed=True)
    ch = c0
    for i in range(len(rates)):
        ch = c0 // 2 ** (i + 1)
        for j, (rk, rd) in enumerate(zip(rks, rds)):
            p = f"resblocks.{i * len(rks) + j}"
            nl = len(rd)
            if block1:
                for l in range(nl):
                    wn(f"{p}.convs1.{l}", ch, ch, rk)
                for l in range(nl):
                    wn(f"{p}.convs2.{l}", ch, ch, rk)
                for m in range(2 * nl):
                    act1d(f"{p}.activations.{m}", ch)
            else:
                for l in range(nl):
                    wn(f"{p}.convs.{l}", ch, ch, rk)
                for m in range(nl):
                    act1d(f"{p}.activations.{m}", ch)
    act1d("activation_post", ch)
    wn("conv_post", 1, ch, 7)
    return spec


def count_parameters(vcfg) -> int:
    """Learnable parameter count (filters are buffers).  112 446 290 for the repo config."""
    return sum(int(np.prod(shape)) for shape, kind in state_dict_spec(vcfg).values() if kind != "filter")


def _uniform(rng, shape, lo, hi):
    return (lo + (hi - lo) * rng.random(shape)).astype(np.float32)


def _normal(rng, shape, std):
    n = int(np.prod(shape))
    m = (n + 1) // 2
    u1 = 1.0 - rng.random(m)  # (0, 1]
    u2 = rng.random(m)
    rad = np.sqrt(-2.0 * np.log(u1))
    z = np.concatenate([rad * np.cos(2 * np.pi * u2), rad * np.sin(2 * np.pi * u2)])[:n]
    return (std * z).reshape(shape).astype(np.float32)


# checkpoint recipes: (log2 half-spread of weight_g around ||v||, alpha mean, alpha std, beta mean, beta std)
RECIPES = {
    # default: halved log-spreads (module docstring): behaves like the default-init net the gates were calibrated on
    "repo": (0.5, 0.0, 0.25, 0.0, 0.25),
    # SURVEY.md section 8d as written: weight_g *= U(0.5, 2) in log scale (2^U(-1, 1)), alpha, beta ~ N(0, 0.5^2).
    # Error-amplifying random net: the reference's own fp32 differs from its fp64 by ~1e-5 on it.
    "survey": (1.0, 0.0, 0.5, 0.0, 0.5),
    # larger snake frequencies, exp(alpha) ~ 2.1 (1.6 .. 2.7 at one sigma): the largest shift for which this
    # random-weight net is still a usable parity target.  Probed with the unmodified reference (48 frames, its own
    # fp32 vs its own fp64, max-abs): alpha ~ N(0.5, .25^2) 1.5e-6, N(0.75, .25^2) 7.0e-6, N(1, .25^2) 4.8e-5,
    # N(1, .5^2) 4.7e-4, N(1.5, .5^2) 1.04 (fully decorrelated: 116 random convolutions behind high-frequency
    # snakes are chaotic).  Larger arguments are pinned at the operator level (activation1d.npz, big_*).
    "large_alpha": (0.5, 0.75, 0.25, 0.0, 0.25),
}


def synthetic_state_dict(vcfg, seed: int = 0, recipe: str = "repo") -> "OrderedDict[str, np.ndarray]":
    """Seeded random-init state_dict in the reference's checkpoint key/shape format (float32).
    ``recipe`` selects the spread of ``weight_g`` and the snake parameters (``RECIPES``); the random
    stream is the same for every recipe, only the scales differ."""
    g_spread, a_mean, a_std, b_mean, b_std = RECIPES[recipe]
    rng = np.random.Generator(np.random.PCG64(seed))
    spec = state_dict_spec(vcfg)
    taps = aa_filter_taps().reshape(1, 1, 12)
    sd: "OrderedDict[str, np.ndarray]" = OrderedDict()
    bounds = {}
    for name, (shape, kind) in spec.items():
        if kind in ("v_conv", "v_convT"):
            # PyTorch's fan_in = size(1) * K for both layouts
            bound = 1.0 / math.sqrt(shape[1] * shape[2])
            bounds[name[: -len(".weight_v")]] = bound
    for name, (shape, kind) in spec.items():
        if kind in ("v_conv", "v_convT"):
            b = bounds[name[: -len(".weight_v")]]
            sd[name] = _uniform(rng, shape, -b, b)
        elif kind == "bias":
            b = bounds[name[: -len(".bias")]]
            sd[name] = _uniform(rng, shape, -b, b)
        elif kind == "g":
            sd[name] = None  # filled once v is known
        elif kind == "alpha":
            sd[name] = (_normal(rng, shape, 1.0) * np.float32(a_std) + np.float32(a_mean)).astype(np.float32)
        elif kind == "beta":
            sd[name] = (_normal(rng, shape, 1.0) * np.float32(b_std) + np.float32(b_mean)).astype(np.float32)
        elif kind == "filter":
            sd[name] = taps.copy()
    for name, (shape, kind) in spec.items():
        if kind == "g":
            v = sd[name[: -len("_g")] + "_v"]
            norm = np.sqrt((v.astype(np.float64) ** 2).sum(axis=(1, 2), keepdims=True))
            sd[name] = (norm * np.exp2(2.0 * g_spread * (rng.random(shape) - 0.5))).astype(np.float32)
    return sd


def diffsvc_state_dict_spec(mcfg) -> "OrderedDict[str, tuple]":
    """``name -> (shape, kind, fan_in)`` of the reference DiffSVC denoiser's ``state_dict()`` (``modules/diffsvc.py:235-282``;
    the step-embedding table is a non-persistent buffer and is not part of it)."""
    n_mel, ch = int(_get(mcfg, "n_mel")), int(_get(mcfg, "residual_channels"))
    fc, cond = int(_get(mcfg, "diffusion_fc_size")), int(_get(mcfg, "conditioner_size"))
    k, nl = int(_get(mcfg, "residual_kernel_size")), int(_get(mcfg, "residual_layer_num"))
    spec: "OrderedDict[str, tuple]" = OrderedDict()

    def conv(prefix, cout, cin, ks):
        spec[prefix + ".weight"] = ((cout, cin, ks), "conv_w", cin * ks)
        spec[prefix + ".bias"] = ((cout,), "bias", cin * ks)

    def linear(prefix, out, inp):
        spec[prefix + ".weight"] = ((out, inp), "lin_w", inp)
        spec[prefix + ".bias"] = ((out,), "bias", inp)

    conv("mel_preprocess.projection", ch, n_mel, 1)
    linear("diffusion_embedding.projection1", fc, 128)
    linear("diffusion_embedding.projection2", fc, fc)
    for i in range(nl):
        p = f"residual_layers.{i}"
        conv(p + ".dilated_conv", 2 * ch, ch, k)
        linear(p + ".diffusion_projection", ch, fc)
        conv(p + ".conditioner_projection", 2 * ch, cond, 1)
        conv(p + ".output_projection", 2 * ch, ch, 1)
    conv("skip_projection", ch, ch, 1)
    conv("output_projection", n_mel, ch, 1)
    return spec


def synthetic_diffsvc_state_dict(mcfg, seed: int = 0) -> "OrderedDict[str, np.ndarray]":
    """Seeded random-init DiffSVC denoiser state_dict (float32, numpy-only, bit-reproducible): convolutions
    ``N(0, 2 / fan_in)`` like the reference's ``kaiming_normal_`` (``modules/diffsvc.py:23-26``), linear layers and biases
    ``U(-1/sqrt(fan_in), 1/sqrt(fan_in))`` like PyTorch's defaults.  The final ``output_projection.weight`` -- zero at
    construction in the reference (``:280``), which would make the whole network invisible in the output -- is drawn
    ``N(0, 1 / fan_in)`` as a trained checkpoint would have it."""
    rng = np.random.Generator(np.random.PCG64(seed))
    sd: "OrderedDict[str, np.ndarray]" = OrderedDict()
    for name, (shape, kind, fan_in) in diffsvc_state_dict_spec(mcfg).items():
        if kind == "conv_w":
            std = math.sqrt((1.0 if name == "output_projection.weight" else 2.0) / fan_in)
            sd[name] = _normal(rng, shape, std)
        else:
            b = 1.0 / math.sqrt(fan_in)
            sd[name] = _uniform(rng, shape, -b, b)
    return sd


_RANGE_FILE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "config", "mel_range.npz")


def mel_range(n_mels: int = 100):
    """Per-band (min, max) of the log-mel the vocoder is fed.

    For 100 bands this is the range ``denormalize_mel_channel`` emits in the reference
    (``utils/acoustic_feature_extraction.py:83-97`` with ``config/mel_{min,max}.pkl``), stored in
    ``config/mel_range.npz`` by ``tests/golden/make_golden.py``.  Other band counts (the 128-band
    512x generator of BASELINE config 5) get the same curve resampled.
    """
    data = np.load(_RANGE_FILE)
    lo, hi = data["mel_min"].astype(np.float32), data["mel_max"].astype(np.float32)
    if n_mels != lo.shape[0]:
        src = np.linspace(0.0, 1.0, lo.shape[0])
        dst = np.linspace(0.0, 1.0, n_mels)
        lo = np.interp(dst, src, lo).astype(np.float32)
        hi = np.interp(dst, src, hi).astype(np.float32)
    return lo, hi


def synthetic_mel(batch: int, n_mels: int, frames: int, seed: int = 1234, dist: str = "logmel") -> np.ndarray:
    """``[B, n_mels, T]`` float32.  ``logmel``: ``min + U(0,1) * (max - min)`` per band; ``randn``: N(0,1)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    if dist == "randn":
        return _normal(rng, (batch, n_mels, frames), 1.0)
    lo, hi = mel_range(n_mels)
    u = rng.random((batch, n_mels, frames)).astype(np.float32)
    return (lo[None, :, None] + u * (hi - lo)[None, :, None]).astype(np.float32)
