"""CPU port of the reference BigVGAN forward on the reference's own arithmetic library
(PyTorch CPU / oneDNN) -- TEST / BASELINE INFRASTRUCTURE ONLY.

The reference's path is a sequence of PyTorch library calls (``modules/bigvgan.py``); the reference
itself cannot travel to the GPU box (nothing there may read ``/root/reference``), so the CPU
baseline that ``bench.py`` reports (``cpu_baseline.kind = "port"``) times this functional
restatement, which issues the same calls per layer as the reference does:

* per-forward weight-norm recompute ``torch._weight_norm(v, g, 0)`` (the reference never removes
  the hooks: ``utils/load_models.py:52-79``),
* ``F.pad(replicate)`` + depthwise ``F.conv_transpose1d(stride 2, groups=C)`` * 2 + crop
  (``UpSample1d.forward`` ``:278-287``), snake (``:84-95``/``:146-159``), ``F.pad(replicate, 5|6)`` +
  depthwise ``F.conv1d(stride 2)`` (``:224-231``),
* ``F.conv1d`` / ``F.conv_transpose1d`` for the dense layers (``:428-431``, ``:602-620``).

Pinned against the same golden vectors as the numpy oracle (tests/test_oracle_golden.py).
Never imported by the product package.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _get(cfg, key):
    return cfg[key] if isinstance(cfg, dict) else getattr(cfg, key)


def _wn(sd, p):
    return torch._weight_norm(sd[p + ".weight_v"], sd[p + ".weight_g"], 0)


def _activation1d(sd, p, x, cfg):
    ch = x.shape[1]
    fu = sd[p + ".upsample.filter"].expand(ch, -1, -1)
    fd = sd[p + ".downsample.lowpass.filter"].expand(ch, -1, -1)
    # upsample x2, K=12: pad 5|5, transposed depthwise conv, gain 2, crop 15|15
    x = F.pad(x, (5, 5), mode="replicate")
    x = 2 * F.conv_transpose1d(x, fu, stride=2, groups=ch)
    x = x[..., 15:-15]
    a = sd[p + ".act.alpha"][None, :, None]
    b = sd[p + ".act.beta"][None, :, None] if _get(cfg, "activation") == "snakebeta" else a
    if _get(cfg, "snake_logscale"):
        a, b = torch.exp(a), torch.exp(b)
    x = x + (1.0 / (b + 1e-9)) * torch.pow(torch.sin(x * a), 2)
    x = F.pad(x, (5, 6), mode="replicate")
    return F.conv1d(x, fd, stride=2, groups=ch)


def _resblock(sd, p, x, cfg, k, dils):
    block1 = _get(cfg, "resblock") == "1"
    for l, d in enumerate(dils):
        if block1:
            xt = _activation1d(sd, f"{p}.activations.{2 * l}", x, cfg)
            xt = F.conv1d(xt, _wn(sd, f"{p}.convs1.{l}"), sd[f"{p}.convs1.{l}.bias"], dilation=d, padding=(k * d - d) // 2)
            xt = _activation1d(sd, f"{p}.activations.{2 * l + 1}", xt, cfg)
            xt = F.conv1d(xt, _wn(sd, f"{p}.convs2.{l}"), sd[f"{p}.convs2.{l}.bias"], padding=(k - 1) // 2)
        else:
            xt = _activation1d(sd, f"{p}.activations.{l}", x, cfg)
            xt = F.conv1d(xt, _wn(sd, f"{p}.convs.{l}"), sd[f"{p}.convs.{l}.bias"], dilation=d, padding=(k * d - d) // 2)
        x = xt + x
    return x


@torch.no_grad()
def generator_forward(sd: dict, cfg, mel: torch.Tensor) -> torch.Tensor:
    """``sd``: reference-format state_dict of torch tensors; ``mel`` ``[B, input_dim, T]``."""
    if _get(cfg, "activation") not in ("snake", "snakebeta"):
        raise NotImplementedError("activation incorrectly specified. check the config file and look for 'activation'.")
    rates, uks = _get(cfg, "upsample_rates"), _get(cfg, "upsample_kernel_sizes")
    rks, rds = _get(cfg, "resblock_kernel_sizes"), _get(cfg, "resblock_dilation_sizes")
    x = F.conv1d(mel, _wn(sd, "conv_pre"), sd["conv_pre.bias"], padding=3)
    for i, (u, k) in enumerate(zip(rates, uks)):
        x = F.conv_transpose1d(x, _wn(sd, f"ups.{i}.0"), sd[f"ups.{i}.0.bias"], stride=u, padding=(k - u) // 2)
        xs = None
        for j, (rk, rd) in enumerate(zip(rks, rds)):
            y = _resblock(sd, f"resblocks.{i * len(rks) + j}", x, cfg, rk, rd)
            xs = y if xs is None else xs + y
        x = xs / len(rks)
    x = _activation1d(sd, "activation_post", x, cfg)
    x = F.conv1d(x, _wn(sd, "conv_post"), sd["conv_post.bias"], padding=3)
    return torch.tanh(x)


def vocoder_inference(sd, cfg, mels: torch.Tensor) -> torch.Tensor:
    """Counterpart of reference ``modules/bigvgan_inference.py:19-26`` on CPU."""
    return generator_forward(sd, cfg, mels).squeeze(1).detach().cpu()
