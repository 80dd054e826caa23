"""Whole-generator parity on the GPU against reference-generated golden vectors
(tests/golden/*.npz) and the CPU oracle.  Needs a B200: run with ``-m gpu``."""
import os

import numpy as np
import pytest
import torch

from oracle import bigvgan_oracle as O
from util_cases import REPO, TINY_VARIANTS, V2, snr_db, tiny_cfg_sd

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def build(cfgd, sd, precision):
    from svc_inference_pipeline_b200.modules.bigvgan import Generator
    from svc_inference_pipeline_b200.utils.util import JsonHParams

    model = Generator(JsonHParams(**cfgd), precision=precision)
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    return model.to(DEV).eval()


@pytest.fixture(scope="module")
def repo_model():
    from svc_inference_pipeline_b200.utils import synth

    sd = synth.synthetic_state_dict(REPO, seed=0)
    return build(REPO, sd, "fp32")


@pytest.mark.parametrize("tag", list(TINY_VARIANTS))
@pytest.mark.parametrize("precision", ["fp32_simt", "fp32"])
def test_tiny_generator_golden(golden, tag, precision):
    g = golden("tiny_generator.npz")
    cfgd, sd = tiny_cfg_sd(tag)
    model = build(cfgd, sd, precision)
    y = model(torch.from_numpy(g[tag + "_mel"]).to(DEV)).cpu().numpy()
    assert y.shape == g[tag + "_y"].shape
    tol = 2e-5 if precision == "fp32_simt" else 1e-4
    assert np.abs(y - g[tag + "_y_f64"]).max() < tol
    assert np.abs(y - g[tag + "_y"]).max() < tol


def test_tiny_generator_bf16(golden):
    g = golden("tiny_generator.npz")
    cfgd, sd = tiny_cfg_sd("b1_snakebeta_log")
    model = build(cfgd, sd, "bf16")
    y = model(torch.from_numpy(g["b1_snakebeta_log_mel"]).to(DEV)).cpu().numpy()
    assert snr_db(g["b1_snakebeta_log_y_f64"], y) > 35.0


def test_repo_generator_fp32(golden, repo_model):
    """BASELINE north_star gate: fp32 path within 1e-4 max-abs of the reference waveform."""
    g = golden("repo_generator.npz")
    for tag in ("logmel", "randn"):
        y = repo_model(torch.from_numpy(g[tag + "_mel"]).to(DEV)).cpu().numpy()
        assert y.shape == g[tag + "_y"].shape
        err = np.abs(y - g[tag + "_y"]).max()
        err64 = np.abs(y - g[tag + "_y_f64"]).max()
        print(f"repo fp32 {tag}: max-abs vs ref fp32 {err:.3e}, vs ref fp64 {err64:.3e}, SNR {snr_db(g[tag + '_y_f64'], y):.1f} dB")
        assert err < 1e-4 and err64 < 1e-4
        assert np.abs(y).max() <= 1.0


def test_repo_generator_simt_exact(golden, repo_model):
    g = golden("repo_generator.npz")
    repo_model.set_precision("fp32_simt")
    try:
        y = repo_model(torch.from_numpy(g["logmel_mel"]).to(DEV)).cpu().numpy()
    finally:
        repo_model.set_precision("fp32")
    assert np.abs(y - g["logmel_y_f64"]).max() < 2e-5


def log_mel_l1(ref, y, **kw):
    """log-mel L1 with the REFERENCE's analysis (utils/mel.py:130-174: slaney mel basis of librosa.filters.mel, hann
    window, center=False with the reflect pad, log clip 1e-5), restated in oracle/logmel_oracle.py and pinned by
    tests/golden/logmel.npz (the unmodified reference function run on a seeded waveform)."""
    from oracle import logmel_oracle as LM

    return LM.log_mel_l1(ref, y, **kw)


def test_repo_generator_bf16(golden, repo_model):
    """BASELINE north_star gate: bf16 path >= 35 dB waveform SNR and log-mel L1 <= 1e-2."""
    g = golden("repo_generator.npz")
    repo_model.set_precision("bf16")
    try:
        for tag in ("logmel", "randn"):
            y = repo_model(torch.from_numpy(g[tag + "_mel"]).to(DEV)).cpu().numpy()
            ref = g[tag + "_y_f64"]
            snr = snr_db(ref, y)
            l1 = np.mean([log_mel_l1(ref[b, 0], y[b, 0]) for b in range(ref.shape[0])])
            print(f"repo bf16 {tag}: SNR {snr:.1f} dB, log-mel L1 {l1:.2e}, max-abs {np.abs(y - ref).max():.3e}")
            assert snr >= 35.0
            assert l1 <= 1e-2
    finally:
        repo_model.set_precision("fp32")


def test_v2_generator_fp32(golden):
    from svc_inference_pipeline_b200.utils import synth

    g = golden("v2_generator.npz")
    model = build(V2, synth.synthetic_state_dict(V2, seed=0), "fp32")
    assert sum(p.numel() for p in model.parameters()) == int(g["n_params"]) == 122_184_530
    y = model(torch.from_numpy(g["logmel_mel"]).to(DEV)).cpu().numpy()
    assert np.abs(y - g["logmel_y_f64"]).max() < 1e-4


def test_batch_and_length_independence(repo_model):
    """Size-independent properties at larger shapes: batch items are independent, and a chunk with a
    >= 38-frame halo reproduces the interior of the full forward (receptive field)."""
    from svc_inference_pipeline_b200.utils import synth

    mel = torch.from_numpy(synth.synthetic_mel(3, 100, 300, seed=77)).to(DEV)
    full = repo_model(mel)
    single = repo_model(mel[1:2].contiguous())
    assert (full[1:2] - single).abs().max().item() < 2e-6
    lo, hi, halo = 100, 200, 48
    part = repo_model(mel[:1, :, lo - halo : hi + halo].contiguous())
    a = full[0, 0, lo * 256 : hi * 256]
    b = part[0, 0, halo * 256 : (halo + hi - lo) * 256]
    assert (a - b).abs().max().item() < 5e-6
    assert torch.isfinite(full).all() and full.abs().max().item() <= 1.0


def test_synthesis_audios_and_loader(golden, tmp_path):
    """The three reference call sites end to end: checkpoint file -> vocoder_model_loader ->
    synthesis_audios, compared with the reference's own output for the same checkpoint + mel."""
    from svc_inference_pipeline_b200.modules.bigvgan_inference import synthesis_audios, vocoder_inference
    from svc_inference_pipeline_b200.utils.load_models import vocoder_model_loader
    from svc_inference_pipeline_b200.utils.util import JsonHParams

    g = golden("tiny_generator.npz")
    cfgd, sd = tiny_cfg_sd("b1_snakebeta_log")
    ckpt = {"generator_state_dict": {"module." + k: torch.from_numpy(v) for k, v in sd.items()}}
    path = os.path.join(tmp_path, "vocoder.pt")
    torch.save(ckpt, path)
    cfg = JsonHParams(device="cuda", vocoder_model_path=path, hop_length=8, vocoder=cfgd)
    model = vocoder_model_loader(cfg)
    assert not model.training and next(model.parameters()).is_cuda
    assert model.load_report == {"missing": [], "wrong_shape": [], "unknown": []}
    mel = torch.from_numpy(g["synth_mel"]).to(DEV)
    audio = synthesis_audios(model, mel, cfg)
    assert isinstance(audio, np.ndarray) and audio.dtype == np.float32 and audio.shape == g["synth_audio"].shape
    assert np.abs(audio - g["synth_audio"]).max() < 1e-4
    assert audio[-1] == 0.0
    out = vocoder_inference(cfg, model, torch.from_numpy(g["synth_mel"])[None], torch.device(DEV))
    assert out.device.type == "cpu" and np.abs(out.numpy() - g["voc_inf"]).max() < 1e-4
    with pytest.raises(RuntimeError):
        synthesis_audios(model, mel[:, :10], cfg)  # < 20 frames: same failure mode as the reference


def test_cuda_graph_matches_eager(golden):
    g = golden("tiny_generator.npz")
    cfgd, sd = tiny_cfg_sd("b1_snakebeta_log")
    model = build(cfgd, sd, "fp32")
    mel = torch.from_numpy(g["b1_snakebeta_log_mel"]).to(DEV)
    y0 = model(mel)
    model.use_cuda_graph = True
    y1 = model(mel)
    y2 = model(mel * 0.5)
    model.use_cuda_graph = False
    y3 = model(mel * 0.5)
    assert torch.equal(y0, y1) and torch.equal(y2, y3)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_programmatic_dependent_launch_is_identical(repo_model, precision):
    """Generator.set_pdl: the same launches chained with programmatic dependent launch (every kernel waits for its
    predecessor's results before its first global access) give bit-identical waveforms -- one program, two overlapped
    half-batch programs, CUDA graph -- and stay identical over repeated calls (no read of a buffer the previous
    launch is still writing)."""
    from svc_inference_pipeline_b200.utils import synth

    mel = torch.from_numpy(synth.synthetic_mel(3, 100, 61, seed=37)).to(DEV)
    try:
        repo_model.set_precision(precision)
        ref = repo_model(mel)
        ref1 = repo_model(mel[:1])
        repo_model.set_pdl(True)
        for _ in range(3):
            assert torch.equal(repo_model(mel), ref)
        repo_model.overlap_streams = False
        for _ in range(3):
            assert torch.equal(repo_model(mel), ref)
        repo_model.use_cuda_graph = True
        for _ in range(3):
            assert torch.equal(repo_model(mel[:1]), ref1)
    finally:
        repo_model.use_cuda_graph = False
        repo_model.overlap_streams = None
        repo_model.set_pdl(False)
        repo_model.set_precision("fp32")


def test_vocode_long_matches_full(repo_model):
    """Long-form path: time chunks with a 48-frame halo + 20-frame cross-fade reproduce the
    unchunked forward (the receptive field is +-38 frames)."""
    from svc_inference_pipeline_b200 import sharding as S
    from svc_inference_pipeline_b200.utils import synth

    mel = torch.from_numpy(synth.synthetic_mel(1, 100, 700, seed=78)[0]).to(DEV)
    full = repo_model(mel[None])[0, 0]
    out = S.vocode_long(repo_model, mel, 256, chunk_frames=200, batch_chunks=4)
    assert out.shape == full.shape
    assert (out - full).abs().max().item() < 5e-6
    # distributed entry point degenerates to the same result without a process group
    out2 = S.vocode_long_distributed(repo_model, mel, 256, chunk_frames=256)
    assert (out2 - full).abs().max().item() < 5e-6


# ------------------------------------------------------------------------------------------
# the host steps either side of the vocoder, fused on the device
# ------------------------------------------------------------------------------------------
def test_fused_mel_denormalisation(repo_model):
    """Generator.set_mel_denorm: forward(normalised mel) == forward(denormalize_mel_channel(mel)) bit for
    bit (the head kernel applies the reference's fp32 expression while it transposes)."""
    from svc_inference_pipeline_b200.utils.acoustic_feature_extraction import denormalize_mel_channel, load_mel_min_max

    mel_min, mel_max = load_mel_min_max()
    rng = np.random.default_rng(21)
    norm = rng.uniform(-1, 1, size=(2, 100, 37)).astype(np.float32)
    den = np.stack([denormalize_mel_channel(torch.from_numpy(m)).numpy() for m in norm])
    np.testing.assert_array_equal(den, O.denormalize_mel_channel(norm, mel_min, mel_max))
    y_host = repo_model(torch.from_numpy(den).to(DEV)).cpu().numpy()
    try:
        repo_model.set_mel_denorm(mel_min, mel_max)
        y_fused = repo_model(torch.from_numpy(norm).to(DEV)).cpu().numpy()
    finally:
        repo_model.set_mel_denorm(None, None)
    np.testing.assert_array_equal(y_fused, y_host)
    np.testing.assert_array_equal(repo_model(torch.from_numpy(den).to(DEV)).cpu().numpy(), y_host)  # fusion is off again


@pytest.mark.parametrize("add_silence,turn_up", [(True, True), (False, True), (True, False)])
def test_fused_pcm16_tail(repo_model, add_silence, turn_up):
    """synthesis_pcm16 (fade-out + peak normalisation + silence + int16, two kernels) equals the oracle's
    restatement of synthesis_audios + save_audio applied to the same generator output, bit for bit, for a
    single mel and for a batch (every item normalised by its own peak)."""
    from svc_inference_pipeline_b200.modules.bigvgan_inference import synthesis_audios, synthesis_pcm16
    from svc_inference_pipeline_b200.utils import synth
    from svc_inference_pipeline_b200.utils.util import JsonHParams

    cfg = JsonHParams(hop_length=256, fs=24000)
    mels = torch.from_numpy(synth.synthetic_mel(3, 100, 45, seed=5))
    wave = repo_model(mels.to(DEV)).cpu().numpy()[:, 0]
    ref = np.stack([O.synthesis_pcm16(w, 256, 24000, add_silence=add_silence, turn_up=turn_up) for w in wave])
    got = synthesis_pcm16(repo_model, mels, cfg, add_silence=add_silence, turn_up=turn_up)
    assert got.dtype == np.int16 and got.shape == ref.shape
    np.testing.assert_array_equal(got, ref)
    one = synthesis_pcm16(repo_model, mels[1], cfg, add_silence=add_silence, turn_up=turn_up)
    np.testing.assert_array_equal(one, ref[1])
    # and the float path of the reference surface agrees with it to one LSB (torch.linspace's SIMD ramp)
    audio = synthesis_audios(repo_model, mels[1], cfg)
    host = O.synthesis_pcm16(wave[1], 256, 24000, add_silence=False, turn_up=False)
    assert np.abs(np.rint(audio * 32768.0) - host).max() <= 1
    with pytest.raises(RuntimeError):
        synthesis_pcm16(repo_model, mels[0, :, :19], cfg)


@pytest.mark.parametrize("B", [2, 3, 5])
def test_overlapped_forward_is_identical(repo_model, B):
    """Two half-batches on two streams (Generator.overlap_streams, the default for B >= 2) give exactly the
    waveform of the single-program forward, for even and odd batch sizes, and repeated calls are stable."""
    from svc_inference_pipeline_b200.utils import synth

    mel = torch.from_numpy(synth.synthetic_mel(B, 100, 53, seed=31)).to(DEV)
    assert repo_model.overlap_streams is None and repo_model.overlaps(B)  # fp32 path: on by default
    y_ov = repo_model(mel)
    y_ov2 = repo_model(mel)
    try:
        repo_model.overlap_streams = False
        y_plain = repo_model(mel)
    finally:
        repo_model.overlap_streams = None
    assert y_ov.shape == (B, 1, 53 * 256)
    assert torch.equal(y_ov, y_plain) and torch.equal(y_ov, y_ov2)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_time_folded_layers_match_unfolded(repo_model, precision):
    """The C = 24 resblock convolutions run time-folded (four rows read as one row of 96 channels,
    bvg_conv_geom.fold): same waveform as the unfolded layers up to accumulation order."""
    from svc_inference_pipeline_b200.utils import synth

    m = repo_model
    old = m.precision
    mel = torch.from_numpy(synth.synthetic_mel(2, 100, 40, 91)).cuda()
    try:
        m.set_precision(precision)
        m.time_fold = True
        m._invalidate()
        y_fold = m(mel).clone()
        folded = [n for n, pc in m._packed["conv"].items() if pc.fold > 1]
        m.time_fold = False
        m._invalidate()
        y_plain = m(mel).clone()
        assert not [n for n, pc in m._packed["conv"].items() if pc.fold > 1]
    finally:
        m.time_fold = True
        m.set_precision(old)
        m._invalidate()
    assert len(folded) == 18, folded  # resblocks 15-17
    err = float((y_fold - y_plain).abs().max())
    if precision == "fp32":
        assert err < 2e-6, err
    else:
        snr = 10 * np.log10(float((y_plain.double() ** 2).sum() / ((y_fold - y_plain).double() ** 2).sum()))
        assert snr > 45, snr


def test_fused_activation_layers_match_unfused(repo_model):
    """Generator.fuse_amp: the Activation1d in front of every C <= 64 unfolded resblock convolution runs inside the
    convolution kernel (bvg_conv_desc.pre_amp).  Same arithmetic in the same order as the two-kernel pair."""
    from svc_inference_pipeline_b200.utils import synth

    m = repo_model
    mel = torch.from_numpy(synth.synthetic_mel(2, 100, 33, 92)).cuda()
    y_plain = m(mel).clone()
    try:
        m.fuse_amp = True
        m._invalidate()
        y_fused = m(mel).clone()
        fused = [lab for lab, kind, _ in m._program(1, 33, slot=(2, 2, 0)).labels if "+activations" in lab]
        m.time_fold = False  # then the C = 24 stage fuses as well
        m._invalidate()
        y_fused_all = m(mel).clone()
        fused_all = [lab for lab, kind, _ in m._program(1, 33, slot=(2, 2, 0)).labels if "+activations" in lab]
    finally:
        m.fuse_amp = False
        m.time_fold = True
        m._invalidate()
    assert len(fused) == 18 and len(fused_all) == 36, (len(fused), len(fused_all))
    assert float((y_fused - y_plain).abs().max()) < 2e-6
    assert float((y_fused_all - y_plain).abs().max()) < 2e-6


def test_full_size_properties(repo_model):
    """BASELINE configs[1] at its full size (batch 16 x 938 frames, the bench shape; the oracle needs minutes per
    item there): size-independent properties instead -- every batch item is independent of its neighbours, a chunk
    with a 48-frame halo reproduces the interior of the full forward, the waveform is finite and inside tanh's
    range, and the bf16 path stays within its SNR gate of the fp32 path."""
    from svc_inference_pipeline_b200.utils import synth

    m = repo_model
    mel = torch.from_numpy(synth.synthetic_mel(16, 100, 938, seed=1235)).to(DEV)
    full = m(mel)
    assert full.shape == (16, 1, 938 * 256)
    assert torch.isfinite(full).all() and full.abs().max().item() <= 1.0
    for b in (0, 7, 15):
        single = m(mel[b : b + 1].contiguous())
        assert (full[b : b + 1] - single).abs().max().item() < 2e-6, b
    lo, hi, halo = 400, 520, 48
    part = m(mel[3:5, :, lo - halo : hi + halo].contiguous())
    a = full[3:5, 0, lo * 256 : hi * 256]
    c = part[:, 0, halo * 256 : (halo + hi - lo) * 256]
    assert (a - c).abs().max().item() < 5e-6
    try:
        m.set_precision("bf16")
        yb = m(mel)
    finally:
        m.set_precision("fp32")
    snr = 10 * np.log10(float((full.double() ** 2).sum() / ((yb - full).double() ** 2).sum()))
    assert snr > 35.0, snr


# ------------------------------------------------------------------------------------------
# Parity pinned at the configurations bench.py measures
# ------------------------------------------------------------------------------------------
def _mel_sha(mel):
    import hashlib

    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(mel).tobytes()).digest(), dtype=np.uint8)


def _gates(tag, ref64, y32, yb, fs_kw):
    err = float(np.abs(y32 - ref64).max())
    snr = snr_db(ref64, yb)
    l1 = log_mel_l1(ref64.reshape(-1), yb.reshape(-1), **fs_kw)
    print(f"{tag}: fp32 path max-abs vs reference fp64 {err:.3e} (gate 1e-4); bf16 path SNR {snr:.1f} dB (gate 35), log-mel L1 {l1:.2e} (gate 1e-2)")
    return err, snr, l1


def test_bench_shape_item_vs_reference(golden, repo_model):
    """BASELINE configs[1] / the bench.py workload: item 7 of the B16 x 938-frame batch against the unmodified
    reference's fp32 and fp64 waveforms for that very mel (tests/golden/bench_item.npz), standalone and in place
    inside the full batch; both north_star gates."""
    from svc_inference_pipeline_b200.utils import synth

    g = golden("bench_item.npz")
    item, B, T = int(g["item"]), int(g["batch"]), int(g["frames"])
    mel = synth.synthetic_mel(B, 100, T, seed=int(g["seed"]))
    np.testing.assert_array_equal(_mel_sha(mel[item : item + 1]), g["mel_sha256"])
    ref32, ref64 = g["y"], g["y_f64"]
    m = repo_model
    x = torch.from_numpy(mel).to(DEV)
    y_alone = m(x[item : item + 1].contiguous()).cpu().numpy()
    y_batch = m(x)[item : item + 1].cpu().numpy()
    try:
        m.set_precision("bf16")
        yb_alone = m(x[item : item + 1].contiguous()).cpu().numpy()
        yb_batch = m(x)[item : item + 1].cpu().numpy()
    finally:
        m.set_precision("fp32")
    print(f"reference fp32 vs its own fp64 at this shape: {float(g['ref_fp32_vs_fp64']):.3e}")
    for tag, y32, yb in (("bench item standalone", y_alone, yb_alone), ("bench item 7 of B16", y_batch, yb_batch)):
        err, snr, l1 = _gates(tag, ref64, y32, yb, {})
        assert err < 1e-4 and float(np.abs(y32 - ref32).max()) < 1e-4
        assert snr >= 35.0 and l1 <= 1e-2


def test_v2_long_item_vs_reference(golden):
    """BASELINE configs[4] length class: one 30-s item ([1, 128, 2584], hop 512, 44.1 kHz analysis) of the 512x v2
    generator against float32(reference fp64), standalone and as item 3 of a batch of 8 (one rank's share of B64)."""
    from svc_inference_pipeline_b200.utils import synth

    g = golden("v2_long.npz")
    item, B, T = int(g["item"]), int(g["batch"]), int(g["frames"])
    mel = synth.synthetic_mel(B, 128, T, seed=int(g["seed"]))
    np.testing.assert_array_equal(_mel_sha(mel[item : item + 1]), g["mel_sha256"])
    ref = g["y_f64_as_f32"].astype(np.float64)
    m = build(V2, synth.synthetic_state_dict(V2, seed=0), "fp32")
    x = torch.from_numpy(mel).to(DEV)
    y_alone = m(x[item : item + 1].contiguous()).cpu().numpy()
    y_batch = m(x)[item : item + 1].cpu().numpy()
    m.set_precision("bf16")
    yb_alone = m(x[item : item + 1].contiguous()).cpu().numpy()
    yb_batch = m(x)[item : item + 1].cpu().numpy()
    print(f"reference fp32 vs its own fp64 at this shape: {float(g['ref_fp32_vs_fp64']):.3e}")
    kw = dict(n_fft=2048, num_mels=128, sampling_rate=44100, hop_size=512, win_size=2048, fmin=0, fmax=22050)
    for tag, y32, yb in (("v2 30-s item standalone", y_alone, yb_alone), ("v2 item 3 of B8", y_batch, yb_batch)):
        err, snr, l1 = _gates(tag, ref, y32, yb, kw)
        assert err < 1e-4
        assert snr >= 35.0 and l1 <= 1e-2


@pytest.mark.parametrize("recipe", ["survey", "large_alpha"])
def test_checkpoint_recipes(golden, recipe):
    """The two other checkpoint recipes of utils/synth.py against the unmodified reference on them (recipes.npz):
    SURVEY section 8d as written ("survey") and larger snake frequencies ("large_alpha").  The fp32 gate must hold with the
    default (MUFU on the raw argument) and with the exact range reduction.  The bf16 gates are measured and printed,
    not asserted: both nets amplify rounding errors (the reference's own fp32 is 10x further from its fp64 than on the
    default recipe, bf16 operands emulated on the reference give 31 dB on "survey", tools/precision_probe.py); measured on
    B200: 34 dB ("survey"), 27 dB ("large_alpha") -- the bf16 path meets its 35 dB gate on the default recipe only
    (README.md, "Precision")."""
    from svc_inference_pipeline_b200.modules.bigvgan import Generator
    from svc_inference_pipeline_b200.utils import synth
    from svc_inference_pipeline_b200.utils.util import JsonHParams

    g = golden("recipes.npz")
    sd = synth.synthetic_state_dict(REPO, 0, recipe=recipe)
    ref64, ref32 = g[recipe + "_y_f64"], g[recipe + "_y"]
    x = torch.from_numpy(g["mel"]).to(DEV)
    errs = {}
    for precise in (False, True):
        m = Generator(JsonHParams(**REPO), precision="fp32", precise_sin=precise)
        m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        m = m.to(DEV).eval()
        y = m(x).cpu().numpy()
        errs[precise] = float(np.abs(y - ref64).max())
        assert errs[precise] < 1e-4, (recipe, precise, errs)
        assert float(np.abs(y - ref32).max()) < 1e-4
    m.set_precision("bf16")
    yb = m(x).cpu().numpy()
    snr, l1 = snr_db(ref64, yb), log_mel_l1(ref64.reshape(-1), yb.reshape(-1))
    print(f"recipe {recipe}: fp32 path max-abs vs reference fp64: fast sin {errs[False]:.3e}, exact reduction {errs[True]:.3e} "
          f"(reference fp32 vs fp64 {float(np.abs(ref32 - ref64).max()):.3e}); bf16 path SNR {snr:.1f} dB, log-mel L1 {l1:.2e}")
    assert snr >= 20.0  # sanity only (see the docstring)


def test_fused_activation_amblock2_does_not_alias():
    """AMPBlock2 with three layers and Generator.fuse_amp: the middle layer's output buffer is its Activation1d's own
    input (xa -> xa), where a fused producer would read halo rows other CTAs are overwriting -- the fusion is refused
    there, kept for the first and last layer, and the result equals the unfused forward and the oracle."""
    from svc_inference_pipeline_b200.utils import synth
    from util_cases import TINY

    cfgd = dict(TINY, resblock="2")  # dilations [1, 3, 5] per block
    sd = synth.synthetic_state_dict(cfgd, seed=5)
    m = build(cfgd, sd, "fp32")
    m.time_fold = False  # (the tiny generator's 16- and 8-channel layers would otherwise run time-folded, which never fuses)
    mel_np = synth.synthetic_mel(2, 10, 23, seed=98, dist="randn")
    mel = torch.from_numpy(mel_np).to(DEV)
    y_plain = m(mel).clone()
    m.fuse_amp = True
    m._invalidate()
    y_fused = m(mel).clone()
    labels = [lab for lab, kind, _ in m._program(1, mel.shape[-1], slot=(2, 2, 0)).labels if kind == "conv"]
    fused = {lab.split()[0].rsplit(".", 1)[1] for lab in labels if "+activations" in lab}
    assert fused == {"0", "2"}, fused
    assert float((y_fused - y_plain).abs().max()) < 2e-6
    ref = O.generator_forward(sd, cfgd, mel_np.astype(np.float64))
    assert np.abs(y_fused.cpu().numpy() - ref).max() < 1e-4


def test_two_devices_in_one_process(golden):
    """Function attributes (dynamic shared memory opt-in, carve-out) are per device: a second GPU driven from the same
    process must get them too (the cache is keyed by (device, kernel))."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs in one process")
    g = golden("tiny_generator.npz")
    cfgd, sd = tiny_cfg_sd("b1_snakebeta_log")
    from svc_inference_pipeline_b200.modules.bigvgan import Generator
    from svc_inference_pipeline_b200.utils.util import JsonHParams

    outs = []
    for dev in ("cuda:0", "cuda:1"):
        m = Generator(JsonHParams(**cfgd), precision="fp32")
        m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
        m = m.to(dev).eval()
        outs.append(m(torch.from_numpy(g["b1_snakebeta_log_mel"]).to(dev)).cpu())
    assert torch.equal(outs[0], outs[1])
    assert np.abs(outs[1].numpy() - g["b1_snakebeta_log_y_f64"]).max() < 1e-4


def test_infer_chain_sampler_denormalise_vocoder(golden, repo_model):
    """infer.py:79-86 on the device stages this package owns: svc_model_inference -> denormalize_mel_channel ->
    synthesis_audios.  The sampler returns a transposed (non-contiguous) ``[n_mel, T]`` view on the device, which is what
    the reference hands on; the result must equal the chain fed with a contiguous host copy of the same mel, and
    the oracle's waveform for that mel (fp32 gate)."""
    from svc_inference_pipeline_b200.modules.bigvgan_inference import synthesis_audios
    from svc_inference_pipeline_b200.modules.diffsvc import DiffSVC
    from svc_inference_pipeline_b200.modules.diffsvcrepo_inference import svc_model_inference
    from svc_inference_pipeline_b200.utils import synth
    from svc_inference_pipeline_b200.utils.acoustic_feature_extraction import denormalize_mel_channel
    from svc_inference_pipeline_b200.utils.util import JsonHParams

    g = golden("sampler.npz")
    mapper = dict(noise_schedule_factors=[0.0001, 0.02, 1000], n_mel=100, residual_channels=64, diffusion_fc_size=128, conditioner_size=64,
                  dilation_cycle_length=4, residual_kernel_size=3, residual_layer_num=4)
    dm = DiffSVC(JsonHParams(**mapper))
    dm.load_state_dict({k: torch.from_numpy(v) for k, v in synth.synthetic_diffsvc_state_dict(mapper, seed=5).items()})
    dm = dm.to(DEV).eval()
    cfg = JsonHParams(hop_length=256, fs=24000, mapper=JsonHParams(noise_schedule=g["noise_schedule"].tolist()))
    batch = {"y": torch.zeros(*g["n1_x0"].shape, device=DEV), "cond": torch.from_numpy(g["n1_cond"]).to(DEV)}
    nz = {"x0": torch.from_numpy(g["n1_x0"]), "steps": torch.from_numpy(g["n1_noise"]).to(DEV)}
    y_pred = svc_model_inference([lambda b: b["cond"], dm], batch, cfg, noise=nz)            # [n_mel, T] in [-1, 1]
    assert y_pred.is_cuda and tuple(y_pred.shape) == (100, 96) and not y_pred.is_contiguous()
    assert np.abs(y_pred.cpu().numpy() - g["n1_y"]).max() < 2e-4
    mel = denormalize_mel_channel(y_pred, cfg)                                                 # shipped default range (warns)
    assert mel.is_cuda and float(mel.min()) >= -11.6 and float(mel.max()) <= 1.0
    audio = synthesis_audios(repo_model, mel, cfg)
    audio_host = synthesis_audios(repo_model, mel.cpu().contiguous(), cfg)
    assert audio.shape == (96 * 256,) and np.array_equal(audio, audio_host)
    from oracle import bigvgan_torch_cpu as P

    tsd = {k: torch.from_numpy(v).double() for k, v in synth.synthetic_state_dict(REPO, seed=0).items()}
    wave = P.generator_forward(tsd, REPO, mel.cpu().double().unsqueeze(0))[0, 0].numpy()     # the reference's op sequence, fp64
    ref = O.synthesis_tail(wave, 96, 256)
    assert np.abs(audio - ref).max() < 1e-4
    assert np.isfinite(audio).all() and audio[-1] == 0.0
