"""CPU oracle for the DiffSVC diffusion sampler (TEST INFRASTRUCTURE ONLY -- never imported by the product path).

Functional restatement of reference ``modules/diffsvcrepo_inference.py``: ``svc_model_inference`` (``:153-240``) with
``p_sample`` (``:88-97``; ``p_mean_variance`` ``:53-85``, ``predict_start_from_noise`` ``:32-36``, ``q_posterior`` ``:39-50``)
and ``p_sample_plms`` (``:100-150``), on the reference's own arithmetic library (PyTorch CPU) in float32 or float64,
without its module-level globals.  The denoiser is a callable ``denoise(x [N, T, n_mel], cond, t [N, 1]) -> eps`` (e.g.
``oracle.diffsvc_oracle.denoiser_forward`` bound to a state_dict); the random draws are arguments.  Pinned by
``tests/golden/sampler.npz``, which ``tests/golden/make_golden.py::golden_sampler`` writes by running the UNMODIFIED
reference function.
"""
from __future__ import annotations

import numpy as np
import torch


def schedule(noise_schedule, dtype=torch.float32) -> dict:
    """The per-step constants (``:163-197``): float64 numpy, then ``to_torch`` (float32 in the reference)."""
    betas = np.array(noise_schedule)
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    pv = betas * (1.0 - ac_prev) / (1.0 - ac)
    to = lambda a: torch.tensor(a, dtype=dtype)
    return dict(
        sqrt_recip=to(np.sqrt(1.0 / ac)), sqrt_recipm1=to(np.sqrt(1.0 / ac - 1)),
        coef1=to(betas * np.sqrt(ac_prev) / (1.0 - ac)), coef2=to((1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac)),
        logvar=to(np.log(np.maximum(pv, 1e-20))), alphas_cumprod=to(ac), steps=len(betas),
    )


def _ext(a, t, x):
    return a.gather(-1, t).reshape(t.shape[0], *((1,) * (x.dim() - 1)))


@torch.no_grad()
def p_sample(denoise, sch, x, t, cond, noise, clip_denoised=True):
    """``x [N, 1, n_mel, T]``, ``t [N]`` long, ``noise`` like ``x``."""
    eps = denoise(x.transpose(-1, -2).squeeze(1), cond, t.unsqueeze(1)).transpose(-1, -2).unsqueeze(1)
    x0 = _ext(sch["sqrt_recip"], t, x) * x - _ext(sch["sqrt_recipm1"], t, x) * eps
    if clip_denoised:
        x0 = x0.clamp(-1.0, 1.0)
    mean = _ext(sch["coef1"], t, x) * x0 + _ext(sch["coef2"], t, x) * x
    mask = (1 - (t == 0).to(x.dtype)).reshape(x.shape[0], *((1,) * (x.dim() - 1)))
    return mean + mask * (0.5 * _ext(sch["logvar"], t, x)).exp() * noise


def _x_pred(sch, x, e, t, interval):
    a_t = _ext(sch["alphas_cumprod"], t, x)
    a_prev = _ext(sch["alphas_cumprod"], torch.max(t - interval, torch.zeros_like(t)), x)
    st, sp = a_t.sqrt(), a_prev.sqrt()
    delta = (a_prev - a_t) * ((1 / (st * (st + sp))) * x - 1 / (st * (((1 - a_prev) * a_t).sqrt() + ((1 - a_t) * a_prev).sqrt())) * e)
    return x + delta


@torch.no_grad()
def p_sample_plms(denoise, sch, x, t, interval, cond, history):
    """One PLMS step; ``history`` is the list of earlier predictions (oldest first) and is appended to."""
    call = lambda x_, t_: denoise(x_.transpose(-1, -2).squeeze(1), cond, t_.unsqueeze(1)).transpose(-1, -2).unsqueeze(1)
    e = call(x, t)
    if len(history) == 0:
        xp = _x_pred(sch, x, e, t, interval)
        e_prev = call(xp, torch.full_like(t, max(int(t[0]) - interval, 0)))
        ep = (e + e_prev) / 2
    elif len(history) == 1:
        ep = (3 * e - history[-1]) / 2
    elif len(history) == 2:
        ep = (23 * e - 16 * history[-1] + 5 * history[-2]) / 12
    else:
        ep = (55 * e - 59 * history[-1] + 37 * history[-2] - 9 * history[-3]) / 24
    history.append(e)
    del history[:-4]
    return _x_pred(sch, x, ep, t, interval)


@torch.no_grad()
def svc_model_inference(denoise, cond, noise_schedule, x0, step_noise=None, fast_inference=False, speedup=10):
    """``x0 [N, T, n_mel]`` (the reference's ``torch.normal`` draw), ``step_noise [steps, N, 1, n_mel, T]`` indexed by the
    diffusion step.  Returns what the reference returns: ``[n_mel, T]`` for N = 1."""
    dt = x0.dtype
    sch = schedule(noise_schedule, dt)
    n = x0.shape[0]
    x = x0.transpose(-1, -2).unsqueeze(1)
    if fast_inference:
        history = []
        for i in reversed(range(0, sch["steps"], speedup)):
            x = p_sample_plms(denoise, sch, x, torch.full((n,), i, dtype=torch.long), speedup, cond, history)
    else:
        for i in reversed(range(0, sch["steps"])):
            x = p_sample(denoise, sch, x, torch.full((n,), i, dtype=torch.long), cond, step_noise[i].to(dt))
    out = x.transpose(-1, -2).squeeze(1).squeeze(0)
    return out.permute(*reversed(range(out.dim())))
