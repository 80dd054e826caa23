"""Diffusion sampler of the DiffSVC acoustic model with the reference's call surface.

Drop-in for ``svc_model_inference(model, batch, cfg, fast_inference=False, speedup=10)`` of the reference's
``modules/diffsvcrepo_inference.py:153-240`` (called by ``infer.py:79``): ``model[0]`` maps the packed features to the
conditioner ``[N, T, cond]`` (the reference's condition encoders, any torch module), ``model[1]`` is this package's
``DiffSVC``.  The sample ``x`` stays in the denoiser's input buffer on the device for the whole run; one diffusion step
is a CUDA-graph replay of the denoiser program followed by (DDPM: captured with) one ``bvg_sample_fwd`` launch, where
the reference runs ~350 eager torch operations and recomputes the 20 conditioner projections.

* ``fast_inference=False``: ``p_sample`` (``:88-97``) for every step of ``cfg.mapper.noise_schedule``, last to first.  The
  initial sample and each step's noise are drawn with the reference's own calls (``torch.normal(0, 1 / 1.2, size=
  batch["y"].shape)``, ``torch.randn(x.shape)``), in the same order, so a seeded run consumes the device generator like the
  reference does.
* ``fast_inference=True``: ``p_sample_plms`` (``:100-150``), every ``speedup``-th step.  The reference's version of this
  branch calls ``.transpose`` on the ``(noise, stats)`` tuple its own denoiser returns and raises; this one takes the
  tuple's first element, which is what the formulas that follow expect (parity is pinned on the unmodified reference
  function driven with a denoiser that returns the tensor alone, ``tests/golden/make_golden.py::golden_sampler``).

There is no CPU path: a ``model[1]`` that is not the B200 ``DiffSVC`` is refused.
"""
from __future__ import annotations

import numpy as np
import torch

from .diffsvc import SCHEDULE_ROWS, DiffSVC

__all__ = ["svc_model_inference", "schedule_tables"]


def schedule_tables(noise_schedule) -> np.ndarray:
    """float32 ``[6, steps]`` in the order of ``SCHEDULE_ROWS``: the per-step constants the reference computes in float64
    numpy and rounds to float32 (``:163-197``)."""
    betas = np.array(noise_schedule)
    if betas.ndim != 1 or betas.size == 0:
        raise ValueError("cfg.mapper.noise_schedule must be a non-empty list of betas")
    alphas = 1.0 - betas
    alphas_cumprod = np.cumprod(alphas, axis=0)
    alphas_cumprod_prev = np.append(1.0, alphas_cumprod[:-1])
    posterior_variance = betas * (1.0 - alphas_cumprod_prev) / (1.0 - alphas_cumprod)
    rows = dict(
        sqrt_recip_alphas_cumprod=np.sqrt(1.0 / alphas_cumprod),
        sqrt_recipm1_alphas_cumprod=np.sqrt(1.0 / alphas_cumprod - 1),
        posterior_mean_coef1=betas * np.sqrt(alphas_cumprod_prev) / (1.0 - alphas_cumprod),
        posterior_mean_coef2=(1.0 - alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - alphas_cumprod),
        posterior_log_variance_clipped=np.log(np.maximum(posterior_variance, 1e-20)),
        alphas_cumprod=alphas_cumprod,
    )
    return np.stack([rows[k] for k in SCHEDULE_ROWS]).astype(np.float32)


def svc_model_inference(model, batch, cfg, fast_inference=False, speedup=10, *, noise=None):
    """Reference ``modules/diffsvcrepo_inference.py:153-240``.  Returns ``[n_mel, T]`` for one utterance (for N > 1 the
    reference's ``.T`` of ``[N, T, n_mel]``, i.e. ``[n_mel, T, N]``).

    ``noise`` (tests): ``{"x0": [N, T, n_mel], "steps": [steps, N, 1, n_mel, T]}`` replaces the generator draws."""
    denoiser = model[1]
    if not isinstance(denoiser, DiffSVC):
        raise TypeError("model[1] must be svc_inference_pipeline_b200.modules.diffsvc.DiffSVC (there is no CPU / eager path)")
    with torch.no_grad():
        cond = model[0](batch)
        device = cond.device
        tables = schedule_tables(cfg.mapper.noise_schedule)
        t = tables.shape[1]
        state = denoiser.sampler(cond.float().contiguous(), tables)
        shape = tuple(batch["y"].shape)  # [N, T, n_mel]
        if shape != tuple(state.x.shape):
            raise ValueError(f"batch['y'] {shape} does not match the conditioner's [N, T] and n_mel {tuple(state.x.shape)}")
        # TODO of the reference kept: the initial sample is N(0, (1 / 1.2)^2)
        x0 = torch.normal(0, 1 / 1.2, size=shape, device=device) if noise is None else noise["x0"].to(device)
        state.x.copy_(x0)
        if fast_inference:
            for i in reversed(range(0, t, speedup)):
                state.plms_step(i, speedup)
        else:
            for i in reversed(range(0, t)):
                if noise is None:
                    state.noise.normal_()          # == torch.randn(x.shape, device=device)
                else:
                    state.noise.copy_(noise["steps"][i])
                state.ddpm_step(i)
        mels_output = state.x.clone().squeeze(0)
        return mels_output.permute(*reversed(range(mels_output.dim())))
