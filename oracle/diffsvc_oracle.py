"""CPU oracle for the DiffSVC denoiser step (TEST INFRASTRUCTURE ONLY -- never imported by the product path).

Functional restatement of reference ``modules/diffsvc.py:284-321`` (``DiffSVC.forward``) with its sub-modules
``StepEncoder.forward`` (``:69-93``, integer and fractional steps), ``SpectrogramPreprocessor.forward`` (``:118-128``) and
``ResidualBlock.forward`` (``:192-232``), on the reference's own arithmetic library (PyTorch CPU), in float32 or
float64, from a reference-format ``state_dict``.  Pinned by ``tests/golden/diffsvc.npz``, which
``tests/golden/make_golden.py`` writes by running the UNMODIFIED reference module on the same seeded state_dict.
"""
from __future__ import annotations

from math import sqrt

import torch
import torch.nn.functional as F


def _get(cfg, key):
    return cfg[key] if isinstance(cfg, dict) else getattr(cfg, key)


def build_embedding(max_steps: int) -> torch.Tensor:
    """``StepEncoder.build_embedding`` (``:45-55``): float32 ``[max_steps, 128]``."""
    steps = torch.arange(max_steps).unsqueeze(1)
    dims = torch.arange(64).unsqueeze(0)
    table = steps * 10.0 ** (dims * 4.0 / 63.0)
    return torch.cat([torch.sin(table), torch.cos(table)], dim=1)


@torch.no_grad()
def denoiser_forward(sd: dict, cfg, mel_spec: torch.Tensor, conditioner: torch.Tensor, diffusion_step: torch.Tensor) -> torch.Tensor:
    """``mel_spec [B, L, n_mel]``, ``conditioner [B, L, cond]``, ``diffusion_step [B, 1]`` int -> ``[B, L, n_mel]``;
    computes in ``mel_spec.dtype`` (``sd`` tensors are cast)."""
    dt = mel_spec.dtype
    w = lambda k: sd[k].to(dt)
    nl = int(_get(cfg, "residual_layer_num"))
    cycle = int(_get(cfg, "dilation_cycle_length"))
    ks = int(_get(cfg, "residual_kernel_size"))
    table = build_embedding(int(_get(cfg, "noise_schedule_factors")[2])).to(device=mel_spec.device, dtype=dt)
    # SpectrogramPreprocessor (:118-128)
    x = F.relu(F.conv1d(mel_spec.transpose(1, 2), w("mel_preprocess.projection.weight"), w("mel_preprocess.projection.bias")))
    # StepEncoder, integer steps (:79-91)
    if diffusion_step.dtype in (torch.int32, torch.int64):
        e = table[diffusion_step]                                                 # [B, 1, 128]
    else:  # lerp_embedding (:57-67), per batch item (the reference's expression broadcasts [B, 1, 128] x [B, 1] to
        # [B, B, 128] for B > 1 and only makes sense for one item; identical to it for B = 1)
        t = diffusion_step.to(dt)
        lo_i, hi_i = torch.floor(t).long(), torch.ceil(t).long()
        e = table[lo_i] + (table[hi_i] - table[lo_i]) * (t - lo_i).unsqueeze(-1)
    e = F.silu(F.linear(e, w("diffusion_embedding.projection1.weight"), w("diffusion_embedding.projection1.bias")))
    e = F.silu(F.linear(e, w("diffusion_embedding.projection2.weight"), w("diffusion_embedding.projection2.bias")))
    skip = None
    cond = conditioner.transpose(1, 2)                                            # :216
    for i in range(nl):
        p = f"residual_layers.{i}"
        d = 2 ** (i % cycle)
        pad = (ks - 1) // 2 if d == 1 else d                                      # :150-170
        step = F.linear(e, w(p + ".diffusion_projection.weight"), w(p + ".diffusion_projection.bias"))  # [B, 1, C]
        y = x + step.transpose(1, 2)                                              # :213
        c = F.conv1d(cond, w(p + ".conditioner_projection.weight"), w(p + ".conditioner_projection.bias"))
        y = F.conv1d(y, w(p + ".dilated_conv.weight"), w(p + ".dilated_conv.bias"), padding=pad, dilation=d) + c  # :220
        gate, filt = torch.chunk(y, 2, dim=1)
        y = torch.sigmoid(gate) * torch.tanh(filt)                                # :227
        y = F.conv1d(y, w(p + ".output_projection.weight"), w(p + ".output_projection.bias"))
        residual, s = torch.chunk(y, 2, dim=1)
        x = (x + residual) / sqrt(2.0)                                            # :232
        skip = s if skip is None else s + skip                                    # :307
    x = skip / sqrt(nl)                                                           # :313
    x = F.relu(F.conv1d(x, w("skip_projection.weight"), w("skip_projection.bias")))
    x = F.conv1d(x, w("output_projection.weight"), w("output_projection.bias"))
    return x.transpose(1, 2)
