#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Runs only in the build container (needs ``/root/reference``); the GPU box and the test-suite
never import the reference -- they read the ``.npz`` files this script writes.  The reference
modules are imported by path and executed on CPU (torch, oneDNN); nothing is copied.

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz and config/mel_range.npz
"""
from __future__ import annotations

import hashlib
import os
import pickle
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("BVG_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from modules import bigvgan as ref  # noqa: E402  (the reference, unmodified)
from modules import bigvgan_inference as ref_inf  # noqa: E402

from svc_inference_pipeline_b200.utils import synth  # noqa: E402
from svc_inference_pipeline_b200.utils.util import JsonHParams, load_config  # noqa: E402

torch.set_grad_enabled(False)


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrays.items()})
    print(f"wrote {name}: " + ", ".join(f"{k}{tuple(np.asarray(v).shape)}" for k, v in arrays.items()))


def rnd(shape, seed, scale=1.0):
    return (np.random.Generator(np.random.PCG64(seed)).standard_normal(shape) * scale).astype(np.float32)


# ---------------------------------------------------------------------------------------------
def golden_filters():
    out = {}
    for tag, (cut, hw, k) in {
        "aa12": (0.25, 0.3, 12),  # the only design the generator instantiates
        "odd9": (0.2, 0.25, 9),
        "k24": (0.125, 0.15, 24),
        "lowatt": (0.25, 0.05, 12),  # exercises the 21 <= A <= 50 beta branch
        "noatt": (0.25, 0.005, 12),  # exercises beta = 0
    }.items():
        out[tag] = ref.kaiser_sinc_filter1d(cut, hw, k).reshape(-1).numpy()
        out[tag + "_args"] = np.array([cut, hw, k], dtype=np.float64)
    save("filters.npz", **out)


def golden_activation():
    out = {}
    x = rnd((2, 5, 37), 11, 1.5)
    alpha = rnd((5,), 12, 0.5)
    beta = rnd((5,), 13, 0.5)
    out.update(x=x, alpha=alpha, beta=beta)
    up, down = ref.UpSample1d(2, 12), ref.DownSample1d(2, 12)
    out["up"] = up(torch.from_numpy(x)).numpy()
    out["down_of_up"] = down(up(torch.from_numpy(x))).numpy()
    out["up_f64"] = up.double()(torch.from_numpy(x).double()).numpy()
    for name, cls in (("snake", ref.Snake), ("snakebeta", ref.SnakeBeta)):
        for logscale in (False, True):
            act = cls(5, alpha_logscale=logscale)
            act.alpha.copy_(torch.from_numpy(alpha) if logscale else torch.from_numpy(1.0 + 0.3 * alpha))
            if name == "snakebeta":
                act.beta.copy_(torch.from_numpy(beta) if logscale else torch.from_numpy(1.0 + 0.3 * beta))
            tag = f"{name}_{'log' if logscale else 'lin'}"
            out[tag + "_act"] = act(torch.from_numpy(x)).numpy()
            a1d = ref.Activation1d(activation=act)
            out[tag + "_a1d"] = a1d(torch.from_numpy(x)).numpy()
            out[tag + "_a1d_f64"] = a1d.double()(torch.from_numpy(x).double()).numpy()
            a1d.float()
    # edge lengths: the replicate clamps dominate
    act = ref.SnakeBeta(3, alpha_logscale=True)
    act.alpha.copy_(torch.from_numpy(alpha[:3]))
    act.beta.copy_(torch.from_numpy(beta[:3]))
    a1d = ref.Activation1d(activation=act)
    for ln in (1, 2, 3, 5, 6, 11, 12, 13):
        xe = rnd((1, 3, ln), 100 + ln, 2.0)
        out[f"edge{ln}_x"] = xe
        out[f"edge{ln}_y"] = a1d(torch.from_numpy(xe)).numpy()
    # large-argument sin: |x * exp(alpha)| ~ 1e2
    act = ref.SnakeBeta(4, alpha_logscale=True)
    act.alpha.copy_(torch.tensor([3.0, 2.0, -1.0, 0.0]))
    act.beta.copy_(torch.tensor([0.5, -2.0, 1.0, 0.0]))
    xb = rnd((1, 4, 64), 21, 6.0)
    out["big_x"] = xb
    out["big_alpha"] = act.alpha.numpy().copy()
    out["big_beta"] = act.beta.numpy().copy()
    a1d_big = ref.Activation1d(activation=act)
    out["big_y"] = a1d_big(torch.from_numpy(xb)).numpy()
    out["big_y_f64"] = a1d_big.double()(torch.from_numpy(xb).double()).numpy()  # shows the reference's own fp32 noise at |a u| ~ 1e2
    save("activation1d.npz", **out)


def golden_convs():
    out = {}
    # weight-normed dilated Conv1d exactly as AMPBlock1 builds it (bigvgan.py:319-386)
    for tag, (c, k, d, ln) in {"c8k3d1": (8, 3, 1, 29), "c8k7d3": (8, 7, 3, 40), "c6k11d5": (6, 11, 5, 33), "c4k11d5_short": (4, 11, 5, 7)}.items():
        conv = ref.weight_norm(ref.Conv1d(c, c, k, 1, dilation=d, padding=ref.get_padding(k, d)))
        conv.weight_g.mul_(torch.from_numpy(0.5 + np.abs(rnd((c, 1, 1), 31))))
        x = rnd((2, c, ln), 32)
        out[tag + "_x"] = x
        out[tag + "_v"] = conv.weight_v.numpy().copy()
        out[tag + "_g"] = conv.weight_g.numpy().copy()
        out[tag + "_b"] = conv.bias.numpy().copy()
        out[tag + "_y"] = conv(torch.from_numpy(x)).numpy()
        out[tag + "_w"] = torch._weight_norm(conv.weight_v, conv.weight_g, 0).numpy()
        out[tag + "_args"] = np.array([c, k, d])
    # weight-normed ConvTranspose1d exactly as Generator builds it (bigvgan.py:547-561)
    for tag, (cin, cout, k, u, ln) in {"t8to4k8u4": (8, 4, 8, 4, 9), "t6to3k4u2": (6, 3, 4, 2, 13), "t4to2k16u8": (4, 2, 16, 8, 5), "t4to2k4u2_len1": (4, 2, 4, 2, 1)}.items():
        conv = ref.weight_norm(ref.ConvTranspose1d(cin, cout, k, u, padding=(k - u) // 2))
        conv.weight_g.mul_(torch.from_numpy(0.5 + np.abs(rnd((cin, 1, 1), 41))))
        x = rnd((2, cin, ln), 42)
        out[tag + "_x"] = x
        out[tag + "_v"] = conv.weight_v.numpy().copy()
        out[tag + "_g"] = conv.weight_g.numpy().copy()
        out[tag + "_b"] = conv.bias.numpy().copy()
        out[tag + "_y"] = conv(torch.from_numpy(x)).numpy()
        out[tag + "_w"] = torch._weight_norm(conv.weight_v, conv.weight_g, 0).numpy()
        out[tag + "_args"] = np.array([cin, cout, k, u])
    save("convs.npz", **out)


TINY = {
    "resblock_kernel_sizes": [3, 7],
    "upsample_rates": [4, 2],
    "input_dim": 10,
    "upsample_initial_channel": 32,
    "resblock": "1",
    "upsample_kernel_sizes": [8, 4],
    "resblock_dilation_sizes": [[1, 3, 5], [1, 3, 5]],
    "activation": "snakebeta",
    "snake_logscale": True,
}


def load_into_reference(vcfg_dict, seed):
    vcfg = JsonHParams(**vcfg_dict)
    model = ref.Generator(vcfg).eval()
    sd = synth.synthetic_state_dict(vcfg_dict, seed)
    ref_sd = model.state_dict()
    assert list(ref_sd.keys()) == list(sd.keys()), "state_dict key grammar/order mismatch"
    for k, v in sd.items():
        assert tuple(ref_sd[k].shape) == v.shape, (k, ref_sd[k].shape, v.shape)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    return model, sd


def golden_tiny():
    out = {}
    variants = {
        "b1_snakebeta_log": dict(),
        "b2_snake_lin": dict(resblock="2", activation="snake", snake_logscale=False, resblock_dilation_sizes=[[1, 3], [1, 3]]),
        "b1_snake_log": dict(activation="snake"),
        "b2_snakebeta_lin": dict(resblock="2", snake_logscale=False, resblock_dilation_sizes=[[1, 3], [1, 3]]),
    }
    for tag, delta in variants.items():
        cfgd = dict(TINY, **delta)
        model, sd = load_into_reference(cfgd, seed=7)
        if not cfgd["snake_logscale"]:
            # linear-scale alphas must stay away from 0: shift the N(0, .5) draw to 1 + 0.3 z
            fixed = {k: torch.from_numpy(1.0 + 0.6 * v) for k, v in sd.items() if k.endswith(".alpha") or k.endswith(".beta")}
            model.load_state_dict(fixed, strict=False)
        mel = synth.synthetic_mel(2, 10, 23, seed=99, dist="randn")
        out[tag + "_mel"] = mel
        out[tag + "_y"] = model(torch.from_numpy(mel)).numpy()
        out[tag + "_y_f64"] = model.double()(torch.from_numpy(mel).double()).numpy()
        model.float()
    # the inference glue (bigvgan_inference.py:29-44) on the default tiny variant, hop = 8
    model, _ = load_into_reference(TINY, seed=7)
    cfg = JsonHParams(hop_length=8)
    mel1 = synth.synthetic_mel(1, 10, 31, seed=98, dist="randn")[0]
    out["synth_mel"] = mel1
    out["synth_audio"] = ref_inf.synthesis_audios(model, torch.from_numpy(mel1), cfg)
    out["voc_inf"] = ref_inf.vocoder_inference(cfg, model, torch.from_numpy(mel1)[None], torch.device("cpu")).numpy()
    save("tiny_generator.npz", **out)


def golden_repo():
    """Full-size generators with the procedural checkpoint (weights are regenerated, not stored)."""
    cfg = load_config(os.path.join(REF, "config", "config.json"))
    vc = cfg.vocoder
    vcfg = {k: vc[k] for k in TINY}
    model, sd = load_into_reference(vcfg, seed=0)
    n_params = sum(p.numel() for p in model.parameters())
    keys_digest = hashlib.sha256("\n".join(f"{k}:{tuple(v.shape)}" for k, v in model.state_dict().items()).encode()).hexdigest()
    out = dict(n_params=n_params, n_tensors=len(model.state_dict()), keys_sha256=np.frombuffer(bytes.fromhex(keys_digest), dtype=np.uint8))
    mel = synth.synthetic_mel(1, 100, 24, seed=1235, dist="logmel")
    out["logmel_mel"] = mel
    out["logmel_y"] = model(torch.from_numpy(mel)).numpy()
    mel2 = synth.synthetic_mel(2, 100, 41, seed=1236, dist="randn")
    out["randn_mel"] = mel2
    out["randn_y"] = model(torch.from_numpy(mel2)).numpy()
    model.double()
    out["logmel_y_f64"] = model(torch.from_numpy(mel).double()).numpy()
    out["randn_y_f64"] = model(torch.from_numpy(mel2).double()).numpy()
    save("repo_generator.npz", **out)
    del model

    # BASELINE config 5: BigVGAN-v2-style 512x generator
    v2 = dict(vcfg, input_dim=128, upsample_rates=[8, 4, 2, 2, 2, 2], upsample_kernel_sizes=[16, 8, 4, 4, 4, 4])
    model, sd = load_into_reference(v2, seed=0)
    out = dict(n_params=sum(p.numel() for p in model.parameters()), n_tensors=len(model.state_dict()))
    mel = synth.synthetic_mel(1, 128, 12, seed=1239, dist="logmel")
    out["logmel_mel"] = mel
    out["logmel_y"] = model(torch.from_numpy(mel)).numpy()
    out["logmel_y_f64"] = model.double()(torch.from_numpy(mel).double()).numpy()
    save("v2_generator.npz", **out)


def mel_range_fixture():
    lo = pickle.load(open(os.path.join(REF, "config", "mel_min.pkl"), "rb"))
    hi = pickle.load(open(os.path.join(REF, "config", "mel_max.pkl"), "rb"))
    path = os.path.join(ROOT, "svc_inference_pipeline_b200", "config", "mel_range.npz")
    np.savez(path, mel_min=np.asarray(lo, np.float32), mel_max=np.asarray(hi, np.float32))
    print("wrote", path)


def _repo_vcfg():
    cfg = load_config(os.path.join(REF, "config", "config.json"))
    return {k: cfg.vocoder[k] for k in TINY}


def golden_bench_item():
    """The BENCHMARKED shape (BASELINE configs[1], bench.py): item 7 of the B16 x 938-frame batch bench.py feeds
    rank 0 (synthetic_mel(16, 100, 938, seed=1235)), through the unmodified reference in fp32 and fp64.  The mel is
    regenerated by the tests (its sha256 is stored); fp64 is stored as float32(y_f64) + the fp32 reference."""
    vcfg = _repo_vcfg()
    model, _ = load_into_reference(vcfg, seed=0)
    item = 7
    mel = synth.synthetic_mel(16, 100, 938, seed=1235)[item : item + 1]
    y32 = model(torch.from_numpy(mel)).numpy()
    y64 = model.double()(torch.from_numpy(mel).double()).numpy()
    save("bench_item.npz", item=item, batch=16, frames=938, seed=1235, mel_sha256=np.frombuffer(hashlib.sha256(mel.tobytes()).digest(), dtype=np.uint8),
         y=y32, y_f64=y64, ref_fp32_vs_fp64=np.abs(y32 - y64).max())


def golden_v2_long():
    """BASELINE configs[4] length class: one 30-s item ([1, 128, 2584], hop 512) of the 512x v2 generator: item 3 of
    the batch tools/time_forward.py --v2 and bench.py's v2 leg use (synthetic_mel(8, 128, 2584, seed=1235)).  The
    waveform has 1.32 M samples, so only float32(reference fp64) is stored, with the reference's own fp32-vs-fp64
    distance next to it."""
    vcfg = dict(_repo_vcfg(), input_dim=128, upsample_rates=[8, 4, 2, 2, 2, 2], upsample_kernel_sizes=[16, 8, 4, 4, 4, 4])
    model, _ = load_into_reference(vcfg, seed=0)
    item = 3
    mel = synth.synthetic_mel(8, 128, 2584, seed=1235)[item : item + 1]
    y32 = model(torch.from_numpy(mel)).numpy()
    y64 = model.double()(torch.from_numpy(mel).double()).numpy()
    save("v2_long.npz", item=item, batch=8, frames=2584, seed=1235, mel_sha256=np.frombuffer(hashlib.sha256(mel.tobytes()).digest(), dtype=np.uint8),
         y_f64_as_f32=y64.astype(np.float32), ref_fp32_vs_fp64=np.abs(y32 - y64).max(), ref_fp32_snr_db=10 * np.log10((y64**2).sum() / ((y32 - y64) ** 2).sum()))


def golden_recipes():
    """Repo generator under the two other checkpoint recipes of utils/synth.py (RECIPES): SURVEY.md section 8d as written
    ("survey") and trained-like large snake frequencies ("large_alpha"), 96 log-mel frames each, fp32 + fp64."""
    vcfg = _repo_vcfg()
    out = {}
    mel = synth.synthetic_mel(1, 100, 96, seed=1240)
    out["mel"] = mel
    for recipe in ("survey", "large_alpha"):
        vc = JsonHParams(**vcfg)
        model = ref.Generator(vc).eval()
        sd = synth.synthetic_state_dict(vcfg, 0, recipe=recipe)
        model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        y32 = model(torch.from_numpy(mel)).numpy()
        y64 = model.double()(torch.from_numpy(mel).double()).numpy()
        out[recipe + "_y"], out[recipe + "_y_f64"] = y32, y64
        print(f"{recipe}: reference fp32 vs fp64 max-abs {np.abs(y32 - y64).max():.3e}, |y|max {np.abs(y64).max():.3f}")
    save("recipes.npz", **out)


def golden_logmel():
    """The reference's log-mel analysis (utils/mel.py:130-174) run UNMODIFIED.  Its mel basis comes from librosa
    (absent here, unpinned by the reference); oracle/logmel_oracle.py restates librosa.filters.mel and is supplied in
    its place, after being checked against the independent implementation in transformers.audio_utils."""
    import types

    from oracle import logmel_oracle as LM

    fb = LM.slaney_mel_filterbank(24000, 1024, 100, 0, 12000)
    try:
        from transformers.audio_utils import mel_filter_bank

        other = mel_filter_bank(num_frequency_bins=513, num_mel_filters=100, min_frequency=0.0, max_frequency=12000.0, sampling_rate=24000, norm="slaney", mel_scale="slaney").T
        dev = np.abs(other - fb).max()
        print(f"slaney filterbank vs transformers.audio_utils.mel_filter_bank: max-abs {dev:.3e} (peak weight {fb.max():.3e})")
        assert dev < 1e-6 * max(1.0, fb.max())
    except ImportError:
        print("transformers not importable: filterbank restatement not cross-checked")
    librosa = types.ModuleType("librosa")
    librosa.filters = types.ModuleType("librosa.filters")
    librosa.filters.mel = lambda sr, n_fft, n_mels, fmin, fmax: LM.slaney_mel_filterbank(sr, n_fft, n_mels, fmin, fmax)
    audio_stub = types.ModuleType("utils.audio")  # utils/mel.py:13 imports the WAV loader (needs soundfile); unused by mel_spectrogram
    audio_stub.load_audio_torch = None
    sys.modules.update({"librosa": librosa, "librosa.filters": librosa.filters, "utils.audio": audio_stub})
    from utils import mel as ref_mel  # the reference, unmodified

    rng = np.random.Generator(np.random.PCG64(5))
    t = np.arange(24000 * 2) / 24000.0
    wave = (0.3 * np.sin(2 * np.pi * (220 + 200 * t) * t) + 0.1 * np.sin(2 * np.pi * 3100 * t) + 0.02 * rng.standard_normal(t.size)).astype(np.float32)
    wave[30000:33000] = 0.0  # a silent stretch: exercises the 1e-5 clip
    m = ref_mel.mel_spectrogram(torch.from_numpy(wave)[None], 1024, 100, 24000, 256, 1024, 0, 12000, center=False).numpy()
    save("logmel.npz", wave=wave, logmel=m, basis=fb)


def golden_diffsvc():
    """DiffSVC denoiser step: the UNMODIFIED reference ``modules/diffsvc.py::DiffSVC`` with the
    reference's own mapper hyper-parameters (config/config.json:54-65) and the procedural state_dict of
    utils/synth.py::synthetic_diffsvc_state_dict (the mapper checkpoint is absent), fp32 and fp64."""
    from modules import diffsvc as ref_d

    cfg = load_config(os.path.join(REF, "config", "config.json"))
    keys = ["noise_schedule_factors", "n_mel", "residual_channels", "diffusion_fc_size", "conditioner_size", "dilation_cycle_length", "residual_kernel_size", "residual_layer_num"]
    mcfg = {k: cfg.mapper[k] for k in keys}
    model = ref_d.DiffSVC(JsonHParams(**mcfg)).eval()
    sd = synth.synthetic_diffsvc_state_dict(mcfg, seed=3)
    ref_sd = model.state_dict()
    assert list(ref_sd.keys()) == list(sd.keys()), "DiffSVC state_dict key grammar/order mismatch"
    for k, v in sd.items():
        assert tuple(ref_sd[k].shape) == v.shape, (k, ref_sd[k].shape, v.shape)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    out = dict(n_params=sum(p.numel() for p in model.parameters()), keys_sha256=np.frombuffer(hashlib.sha256("\n".join(f"{k}:{tuple(v.shape)}" for k, v in ref_sd.items()).encode()).digest(), dtype=np.uint8))
    for tag, (B, Ln, steps) in {"b2": (2, 61, [17, 903]), "utt": (1, 379, [500])}.items():
        mel = rnd((B, Ln, mcfg["n_mel"]), 50 + B)
        cond = rnd((B, Ln, mcfg["conditioner_size"]), 60 + B)
        t = torch.tensor(steps, dtype=torch.long).unsqueeze(1)
        y32, _ = model(torch.from_numpy(mel), torch.from_numpy(cond), t)
        model.double()
        y64, _ = model(torch.from_numpy(mel).double(), torch.from_numpy(cond).double(), t)
        model.float()
        out.update({tag + "_mel": mel, tag + "_cond": cond, tag + "_steps": np.asarray(steps), tag + "_y": y32.numpy(), tag + "_y_f64": y64.numpy()})
        print(f"diffsvc {tag}: reference fp32 vs fp64 max-abs {np.abs(y32.numpy() - y64.numpy()).max():.3e}, |y|max {np.abs(y64.numpy()).max():.3f}")
    save("diffsvc.npz", **out)


def golden_sampler():
    """Diffusion sampler (the caller of the denoiser step): the UNMODIFIED reference ``modules/diffsvcrepo_inference.py::
    svc_model_inference`` around the unmodified reference DiffSVC (a 4-layer, 64-channel mapper: the sampler is
    independent of the denoiser's size), 12-step schedule, N = 2 and N = 1.  The reference draws its noise from the
    global generator: the draws are reproduced here by re-seeding and repeating its calls (``torch.normal`` of
    batch["y"].shape, then one ``torch.randn(x.shape)`` per step, last step first) and saved with the outputs (float32
    only: a float64 run of the reference would draw different numbers).
    ``fast``: the PLMS branch, which in the reference fails on its own denoiser's (noise, stats) tuple; it is run
    with a denoiser wrapper that returns the tensor alone (speedup 3 -> steps 9, 6, 3, 0 = all four branches)."""
    from modules import diffsvc as ref_d
    from modules import diffsvcrepo_inference as ref_s

    mcfg = dict(noise_schedule_factors=[0.0001, 0.02, 1000], n_mel=100, residual_channels=64, diffusion_fc_size=128, conditioner_size=64,
                dilation_cycle_length=4, residual_kernel_size=3, residual_layer_num=4)
    model = ref_d.DiffSVC(JsonHParams(**mcfg)).eval()
    sd = synth.synthetic_diffsvc_state_dict(mcfg, seed=5)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    sched = np.linspace(1e-4, 0.35, 12).tolist()
    cfg = JsonHParams(mapper=JsonHParams(noise_schedule=sched))
    out = dict(noise_schedule=np.asarray(sched))

    class Cond(torch.nn.Module):
        def forward(self, batch):
            return batch["cond"]

    class TensorOnly(torch.nn.Module):
        def __init__(self, m):
            super().__init__()
            self.m = m

        def forward(self, *a):
            return self.m(*a)[0]

    for tag, (N, T) in {"n2": (2, 61), "n1": (1, 96)}.items():
        cond = torch.from_numpy(rnd((N, T, mcfg["conditioner_size"]), 70 + N))
        batch = {"y": torch.zeros(N, T, mcfg["n_mel"]), "cond": cond}
        seed = 900 + N
        torch.manual_seed(seed)
        x0 = torch.normal(0, 1 / 1.2, size=batch["y"].shape)
        steps = torch.zeros(len(sched), N, 1, mcfg["n_mel"], T)
        for i in reversed(range(len(sched))):
            steps[i] = torch.randn(N, 1, mcfg["n_mel"], T)
        torch.manual_seed(seed)
        y = ref_s.svc_model_inference(torch.nn.ModuleList([Cond(), model]), batch, cfg)
        out[tag + "_y"] = y.numpy()
        torch.manual_seed(seed)
        yf = ref_s.svc_model_inference(torch.nn.ModuleList([Cond(), TensorOnly(model)]), batch, cfg, fast_inference=True, speedup=3)
        out.update({tag + "_cond": cond.numpy(), tag + "_x0": x0.numpy(), tag + "_noise": steps.numpy(), tag + "_fast": yf.numpy()})
        print(f"sampler {tag}: out {tuple(y.shape)}, clamped at the last step {float((y.abs() >= 1).float().mean()):.2f}, fast |y|max {float(yf.abs().max()):.3f}")
    save("sampler.npz", **out)


STEPS = {
    "mel_range": mel_range_fixture, "filters": golden_filters, "activation": golden_activation, "convs": golden_convs, "tiny": golden_tiny,
    "repo": golden_repo, "bench_item": golden_bench_item, "v2_long": golden_v2_long, "recipes": golden_recipes, "logmel": golden_logmel, "diffsvc": golden_diffsvc, "sampler": golden_sampler,
}

if __name__ == "__main__":
    torch.manual_seed(0)
    for name in (sys.argv[1:] or list(STEPS)):
        STEPS[name]()
