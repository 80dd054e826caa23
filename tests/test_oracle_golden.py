"""Pin the CPU oracle against vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU-only."""
import numpy as np
import pytest

from oracle import bigvgan_oracle as O
from svc_inference_pipeline_b200.utils import synth

from util_cases import REPO, TINY, TINY_VARIANTS  # noqa: E402

def tiny_sd(tag):
    cfg = dict(TINY, **TINY_VARIANTS[tag])
    sd = synth.synthetic_state_dict(cfg, seed=7)
    if not cfg["snake_logscale"]:
        for k in sd:
            if k.endswith(".alpha") or k.endswith(".beta"):
                sd[k] = (1.0 + 0.6 * sd[k]).astype(np.float32)
    return cfg, sd


@pytest.mark.parametrize("tag", ["aa12", "odd9", "k24", "lowatt", "noatt"])
def test_filter_design(golden, tag):
    g = golden("filters.npz")
    cut, hw, k = g[tag + "_args"]
    f = O.kaiser_sinc_filter1d(float(cut), float(hw), int(k))
    np.testing.assert_allclose(f, g[tag], rtol=0, atol=2e-7)
    assert abs(f.sum() - 1.0) < 1e-6


def test_filter_known_taps(golden):
    # SURVEY.md section 8 a1 golden constants
    taps = [0.0020289647, 0.0093894657, -0.0255434588, -0.0576573834, 0.1285725832, 0.4432097971]
    f = synth.aa_filter_taps()
    np.testing.assert_allclose(f[:6], taps, atol=2e-7)
    np.testing.assert_allclose(f[::-1], f, atol=0)
    np.testing.assert_allclose(f, golden("filters.npz")["aa12"], atol=2e-7)


def test_resamplers(golden):
    g = golden("activation1d.npz")
    f = golden("filters.npz")["aa12"]
    x = g["x"]
    up = O.upsample1d(x, f)
    np.testing.assert_allclose(up, g["up"], atol=2e-6)
    np.testing.assert_allclose(O.upsample2x_closed_form(x, f), g["up"], atol=2e-6)
    np.testing.assert_allclose(O.upsample1d(x.astype(np.float64), f.astype(np.float64)), g["up_f64"], atol=1e-13)
    np.testing.assert_allclose(O.upsample2x_closed_form(x.astype(np.float64), f.astype(np.float64)), g["up_f64"], atol=1e-13)
    np.testing.assert_allclose(O.lowpass_downsample1d(g["up"], f), g["down_of_up"], atol=2e-6)
    np.testing.assert_allclose(O.downsample2x_closed_form(g["up"], f), g["down_of_up"], atol=2e-6)


@pytest.mark.parametrize("name", ["snake", "snakebeta"])
@pytest.mark.parametrize("scale", ["lin", "log"])
def test_activation1d(golden, name, scale):
    g = golden("activation1d.npz")
    f = golden("filters.npz")["aa12"]
    log = scale == "log"
    alpha = g["alpha"] if log else 1.0 + 0.3 * g["alpha"]
    beta = None if name == "snake" else (g["beta"] if log else 1.0 + 0.3 * g["beta"])
    tag = f"{name}_{scale}"
    np.testing.assert_allclose(O.snake(g["x"], alpha.astype(np.float32), None if beta is None else beta.astype(np.float32), log), g[tag + "_act"], atol=3e-6, rtol=1e-6)
    y = O.activation1d(g["x"], alpha.astype(np.float32), None if beta is None else beta.astype(np.float32), log, f, f)
    np.testing.assert_allclose(y, g[tag + "_a1d"], atol=5e-6, rtol=1e-6)
    y64 = O.activation1d(g["x"].astype(np.float64), alpha.astype(np.float32).astype(np.float64), None if beta is None else beta.astype(np.float32).astype(np.float64), log, f.astype(np.float64), f.astype(np.float64))
    np.testing.assert_allclose(y64, g[tag + "_a1d_f64"], atol=1e-12)


@pytest.mark.parametrize("ln", [1, 2, 3, 5, 6, 11, 12, 13])
def test_activation1d_edges(golden, ln):
    g = golden("activation1d.npz")
    f = golden("filters.npz")["aa12"]
    y = O.activation1d(g[f"edge{ln}_x"], g["alpha"][:3], g["beta"][:3], True, f, f)
    np.testing.assert_allclose(y, g[f"edge{ln}_y"], atol=1e-5, rtol=1e-6)


def test_activation1d_large_argument(golden):
    g = golden("activation1d.npz")
    f = golden("filters.npz")["aa12"]
    y = O.activation1d(g["big_x"], g["big_alpha"], g["big_beta"], True, f, f)
    np.testing.assert_allclose(y, g["big_y"], atol=2e-4, rtol=1e-5)  # sin of |arg| ~ 1e2 in fp32


@pytest.mark.parametrize("tag", ["c8k3d1", "c8k7d3", "c6k11d5", "c4k11d5_short"])
def test_conv1d(golden, tag):
    g = golden("convs.npz")
    c, k, d = g[tag + "_args"]
    w = O.weight_norm_fold(g[tag + "_v"], g[tag + "_g"])
    np.testing.assert_allclose(w, g[tag + "_w"], atol=1e-6, rtol=1e-6)
    y = O.conv1d(g[tag + "_x"], w, g[tag + "_b"], int(d), O.get_padding(int(k), int(d)))
    np.testing.assert_allclose(y, g[tag + "_y"], atol=5e-6, rtol=1e-5)


@pytest.mark.parametrize("tag", ["t8to4k8u4", "t6to3k4u2", "t4to2k16u8", "t4to2k4u2_len1"])
def test_conv_transpose1d(golden, tag):
    g = golden("convs.npz")
    cin, cout, k, u = g[tag + "_args"]
    w = O.weight_norm_fold(g[tag + "_v"], g[tag + "_g"])
    np.testing.assert_allclose(w, g[tag + "_w"], atol=1e-6, rtol=1e-6)
    y = O.conv_transpose1d(g[tag + "_x"], w, g[tag + "_b"], int(u), int(k - u) // 2)
    assert y.shape == g[tag + "_y"].shape
    np.testing.assert_allclose(y, g[tag + "_y"], atol=5e-6, rtol=1e-5)


@pytest.mark.parametrize("tag", list(TINY_VARIANTS))
def test_tiny_generator(golden, tag):
    g = golden("tiny_generator.npz")
    cfg, sd = tiny_sd(tag)
    y = O.generator_forward(sd, cfg, g[tag + "_mel"])
    np.testing.assert_allclose(y, g[tag + "_y"], atol=2e-5)
    y64 = O.generator_forward(sd, cfg, g[tag + "_mel"].astype(np.float64))
    np.testing.assert_allclose(y64, g[tag + "_y_f64"], atol=1e-11)


def test_synthesis_tail(golden):
    g = golden("tiny_generator.npz")
    cfg, sd = tiny_sd("b1_snakebeta_log")
    mel = g["synth_mel"]
    wav = O.generator_forward(sd, cfg, mel[None])[0, 0]
    np.testing.assert_allclose(wav, g["voc_inf"][0], atol=2e-5)
    audio = O.synthesis_tail(wav, mel.shape[-1], 8)
    assert audio.shape == g["synth_audio"].shape and audio.dtype == np.float32
    np.testing.assert_allclose(audio, g["synth_audio"], atol=2e-5)
    assert audio[-1] == 0.0
    with pytest.raises(ValueError):
        O.synthesis_tail(wav[:80], 10, 8)  # fewer than 20 frames: the reference's broadcast fails too


def test_unknown_activation_raises():
    with pytest.raises(NotImplementedError):
        synth.state_dict_spec(dict(TINY, activation="relu"))


def test_repo_structure(golden):
    g = golden("repo_generator.npz")
    assert int(g["n_params"]) == 112_446_290 == synth.count_parameters(REPO)
    assert int(g["n_tensors"]) == 784 == len(synth.state_dict_spec(REPO))
    import hashlib
    digest = hashlib.sha256("\n".join(f"{k}:{tuple(s)}" for k, (s, _) in synth.state_dict_spec(REPO).items()).encode()).digest()
    assert bytes(g["keys_sha256"].tobytes()) == digest
    v2 = dict(REPO, input_dim=128, upsample_rates=[8, 4, 2, 2, 2, 2], upsample_kernel_sizes=[16, 8, 4, 4, 4, 4])
    assert int(golden("v2_generator.npz")["n_params"]) == 122_184_530 == synth.count_parameters(v2)


def test_repo_generator_oracle(golden):
    """Full-size repo generator, procedural checkpoint, fp32 oracle vs fp32 and fp64 reference."""
    g = golden("repo_generator.npz")
    sd = synth.synthetic_state_dict(REPO, seed=0)
    y = O.generator_forward(sd, REPO, g["logmel_mel"])
    assert y.shape == (1, 1, 24 * 256)
    assert np.abs(y - g["logmel_y"]).max() < 2e-5
    assert np.abs(y - g["logmel_y_f64"]).max() < 2e-5


def test_torch_cpu_port_matches_goldens(golden):
    """The PyTorch-CPU port used as bench.py's CPU baseline is pinned to the same vectors."""
    import torch

    from oracle import bigvgan_torch_cpu as P

    g = golden("tiny_generator.npz")
    for tag in TINY_VARIANTS:
        cfg, sd = tiny_sd(tag)
        tsd = {k: torch.from_numpy(v) for k, v in sd.items()}
        y = P.generator_forward(tsd, cfg, torch.from_numpy(g[tag + "_mel"])).numpy()
        np.testing.assert_allclose(y, g[tag + "_y"], atol=2e-6)
    g = golden("repo_generator.npz")
    tsd = {k: torch.from_numpy(v) for k, v in synth.synthetic_state_dict(REPO, seed=0).items()}
    y = P.generator_forward(tsd, REPO, torch.from_numpy(g["logmel_mel"])).numpy()
    assert np.abs(y - g["logmel_y"]).max() < 2e-6


def test_linspace_restates_torch():
    """The fade-out ramp: the oracle's scalar restatement of torch.linspace(1, 0, n) equals torch's own
    output up to one ulp (torch's vectorised CPU path depends on the host SIMD width), exactly for
    short ramps, and has the reference's end points."""
    import torch

    for n in (1, 2, 3, 7, 20, 256, 5120, 20 * 512):
        ours = O.linspace_1_0(n)
        ref = torch.linspace(1, 0, steps=n).numpy()
        assert ours[0] == 1.0 and (n == 1 or ours[-1] == 0.0)
        assert np.abs(ours - ref).max() <= 2**-23
        if n <= 20:
            np.testing.assert_array_equal(ours, ref)


def test_denormalize_mel_channel_matches_numpy_expression():
    """denormalize_mel_channel against the reference's literal numpy expression
    (utils/acoustic_feature_extraction.py:93) on the reference's own mel_min / mel_max statistics."""
    import os

    d = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "svc_inference_pipeline_b200", "config", "mel_range.npz"))
    mel_min, mel_max = d["mel_min"], d["mel_max"]
    rng = np.random.default_rng(3)
    mel = rng.uniform(-1, 1, size=(100, 57)).astype(np.float32)
    ZERO = 1e-12
    ref = (mel + 1) / 2 * (np.expand_dims(mel_max, -1) - np.expand_dims(mel_min, -1) + ZERO) + np.expand_dims(mel_min, -1)
    got = O.denormalize_mel_channel(mel, mel_min, mel_max)
    assert ref.dtype == np.float32
    np.testing.assert_array_equal(got, ref)
    from svc_inference_pipeline_b200.utils.acoustic_feature_extraction import denormalize_mel_channel
    import torch

    np.testing.assert_array_equal(denormalize_mel_channel(torch.from_numpy(mel)).numpy(), ref)


def test_synthesis_pcm16_oracle_vs_reference_steps(tmp_path):
    """The fused tail restated in the oracle equals the reference's host sequence -- synthesis_audios'
    fade (torch.linspace) then save_audio's numpy arithmetic -- within one PCM LSB, and
    utils.util.save_audio writes that PCM into a readable 16-bit WAV."""
    import wave

    import torch

    from svc_inference_pipeline_b200.utils.util import save_audio

    rng = np.random.default_rng(4)
    hop, fs, frames = 256, 24000, 41
    audio = (rng.standard_normal(frames * hop) * 0.2).astype(np.float32)
    pcm = O.synthesis_pcm16(audio, hop, fs)
    a = torch.from_numpy(audio.copy())
    a[-20 * hop:] *= torch.linspace(1, 0, steps=20 * hop)
    w = a.numpy()
    ratio = 0.9 / max(w.max(), abs(w.min()))
    w = w * ratio
    sil = np.zeros((fs // 20,), dtype=w.dtype)
    w = np.concatenate([sil, w, sil])
    ref = np.clip(np.rint(w.astype(np.float64) * 32768.0), -32768, 32767).astype(np.int16)
    assert pcm.shape == ref.shape == (frames * hop + 2 * (fs // 20),)
    assert np.abs(pcm.astype(np.int32) - ref.astype(np.int32)).max() <= 1
    assert np.abs(pcm).max() in (29491, 29492)  # 0.9 * 32768
    path = str(tmp_path / "x.wav")
    save_audio(path, pcm, fs)
    with wave.open(path) as f:
        assert (f.getframerate(), f.getsampwidth(), f.getnchannels(), f.getnframes()) == (fs, 2, 1, pcm.size)
        np.testing.assert_array_equal(np.frombuffer(f.readframes(pcm.size), dtype=np.int16), pcm)
    save_audio(path, audio, fs)  # float input: processed on the host like the reference
    with wave.open(path) as f:
        assert f.getnframes() == audio.size + 2 * (fs // 20)


# ------------------------------------------------------------------------------------------
# log-mel analysis (reference utils/mel.py:130-174) and the bench-shape pins
# ------------------------------------------------------------------------------------------
def test_logmel_oracle_vs_reference(golden):
    """oracle/logmel_oracle.py against the unmodified reference mel_spectrogram run by make_golden.py on a seeded
    2-s waveform (with a silent stretch: the 1e-5 clip), and the slaney mel basis against the one committed there."""
    from oracle import logmel_oracle as LM

    g = golden("logmel.npz")
    fb = LM.slaney_mel_filterbank(24000, 1024, 100, 0, 12000)
    assert fb.shape == (100, 513) and fb.dtype == np.float32
    np.testing.assert_array_equal(fb, g["basis"])
    # slaney area normalisation: each triangle integrates to ~1 over Hz (bin width 24000 / 1024)
    area = fb.sum(axis=1) * (24000 / 1024)
    assert np.abs(area[5:] - 1.0).max() < 0.15 and np.abs(area[40:] - 1.0).max() < 0.05  # narrow low bands sample the triangle coarsely
    m64 = LM.mel_spectrogram(g["wave"], dtype=np.float64)
    assert m64.shape == g["logmel"].shape == (1, 100, 187)
    # the reference computes in fp32: bins at the clip floor are exact, the rest agree to fp32 rounding of the STFT
    assert np.abs(m64 - g["logmel"]).max() < 1e-4   # measured 1.3e-5
    assert np.abs(m64 - g["logmel"]).mean() < 5e-6  # measured 4.8e-7
    assert (g["logmel"] == np.float32(np.log(np.float32(1e-5)))).any()  # the silent stretch hits the clip
    assert LM.log_mel_l1(g["wave"], g["wave"]) == 0.0
    assert 0 < LM.log_mel_l1(g["wave"], g["wave"] * 1.01) < 0.011


def test_bench_shape_golden_is_the_bench_input(golden):
    """tests/golden/bench_item.npz belongs to the mel bench.py feeds rank 0 (item 7 of synthetic_mel(16, 100, 938, 1235))."""
    import hashlib

    g = golden("bench_item.npz")
    mel = synth.synthetic_mel(int(g["batch"]), 100, int(g["frames"]), seed=int(g["seed"]))[int(g["item"])][None]
    assert (int(g["batch"]), int(g["frames"]), int(g["seed"])) == (16, 938, 1235)
    np.testing.assert_array_equal(np.frombuffer(hashlib.sha256(mel.tobytes()).digest(), dtype=np.uint8), g["mel_sha256"])
    assert g["y"].shape == g["y_f64"].shape == (1, 1, 938 * 256)
    assert float(g["ref_fp32_vs_fp64"]) < 1e-5 and np.abs(g["y_f64"]).max() <= 1.0


@pytest.mark.parametrize("recipe", ["survey", "large_alpha"])
def test_torch_cpu_port_vs_reference_on_recipes(golden, recipe):
    """The PyTorch-CPU port (the CPU baseline bench.py times) reproduces the unmodified reference on the two other
    checkpoint recipes of utils/synth.py, 96 log-mel frames of the full repo generator."""
    import torch

    from oracle import bigvgan_torch_cpu as port

    g = golden("recipes.npz")
    sd = {k: torch.from_numpy(v) for k, v in synth.synthetic_state_dict(REPO, 0, recipe=recipe).items()}
    y = port.generator_forward(sd, REPO, torch.from_numpy(g["mel"])).numpy()
    assert np.abs(y - g[recipe + "_y"]).max() < 2e-5
    assert np.abs(y - g[recipe + "_y_f64"]).max() < 5e-5
