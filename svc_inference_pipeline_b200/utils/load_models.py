"""Checkpoint loading with the reference's call surface (reference ``utils/load_models.py:52-79``).

``vocoder_model_loader(cfg)`` builds ``Generator(cfg.vocoder)``, reads
``torch.load(cfg.vocoder_model_path)["generator_state_dict"]``, strips a ``module.`` prefix
(``k.split("module.")[-1]``), keeps only tensors whose name *and* shape match, loads them, moves the
model to CUDA iff ``cfg.device == "cuda"`` and puts it in eval mode.  The one deliberate
difference: the reference drops mismatched tensors silently (they keep their random init);
this loader does the same but *reports* them (``model.load_report``) and warns.
"""
from __future__ import annotations

import warnings

import torch

from ..modules.bigvgan import Generator
from ..modules.diffsvc import DiffSVC


def filter_state_dict(pretrained: dict, target: dict):
    """The reference's name+shape filter (``utils/load_models.py:63-70``) plus a report."""
    kept, wrong_shape, unknown = {}, [], []
    for k, v in pretrained.items():
        name = k.split("module.")[-1]
        if name not in target:
            unknown.append(name)
        elif tuple(v.shape) != tuple(target[name].shape):
            wrong_shape.append((name, tuple(v.shape), tuple(target[name].shape)))
        else:
            kept[name] = v
    missing = [k for k in target if k not in kept]
    return kept, {"missing": missing, "wrong_shape": wrong_shape, "unknown": unknown}


def vocoder_model_loader(cfg, precision: str = "fp32"):
    print("Loading vocoder model from ", cfg.vocoder_model_path)
    model = Generator(cfg.vocoder, precision=precision)
    ckpt = torch.load(cfg.vocoder_model_path, map_location=torch.device(cfg.device))
    pretrained = ckpt["generator_state_dict"]
    generator_dict = model.state_dict()
    kept, report = filter_state_dict(pretrained, generator_dict)
    generator_dict.update(kept)
    model.load_state_dict(generator_dict)
    model.load_report = report
    if report["missing"] or report["wrong_shape"] or report["unknown"]:
        warnings.warn(
            f"vocoder checkpoint: {len(report['missing'])} tensors keep their init values, "
            f"{len(report['wrong_shape'])} had a mismatched shape, {len(report['unknown'])} were not recognised"
        )
    if cfg.device == "cuda":
        model = model.cuda()
    model = model.eval()
    return model


def denoiser_model_loader(cfg, precision: str = "fp32"):
    """The DiffSVC denoiser out of the reference's MAPPER checkpoint.

    ``svc_model_loader`` (reference ``utils/load_models.py:23-50``) builds ``ModuleList([EncoderFramework(cfg.mapper),
    DiffSVC(cfg.mapper)])`` (``:18-21``) and loads ``torch.load(cfg.svc_model_path)["state_dict"]`` with the same
    ``module.``-prefix strip and name+shape filter as the vocoder loader; the denoiser's tensors are therefore the
    keys that start with ``"1."`` (index 1 of the ModuleList).  This loader takes exactly those into the B200
    ``DiffSVC`` (the condition encoders, index 0, are out of scope) and reports what did not match."""
    print("Load mapper model from ", cfg.svc_model_path)
    model = DiffSVC(cfg.mapper, precision=precision)
    ckpt = torch.load(cfg.svc_model_path, map_location=torch.device(cfg.device))
    pretrained = {k.split("module.")[-1]: v for k, v in ckpt["state_dict"].items()}
    mine = {k[2:]: v for k, v in pretrained.items() if k.startswith("1.")}
    weights = model.state_dict()
    kept, report = filter_state_dict(mine, weights)
    report["other_modules"] = sorted({k.split(".")[0] for k in pretrained if not k.startswith("1.")})
    weights.update(kept)
    model.load_state_dict(weights)
    model.load_report = report
    if report["missing"] or report["wrong_shape"] or report["unknown"]:
        warnings.warn(
            f"mapper checkpoint: {len(report['missing'])} denoiser tensors keep their init values, "
            f"{len(report['wrong_shape'])} had a mismatched shape, {len(report['unknown'])} were not recognised"
        )
    if cfg.device == "cuda":
        model = model.cuda()
    return model.eval()
