"""Log-mel front end on the GPU (bvg_logmel_fwd) against the unmodified reference's
mel_spectrogram output (tests/golden/logmel.npz) and the fp64 CPU oracle.  Needs a B200: run with ``-m gpu``."""
import numpy as np
import pytest
import torch

from oracle import logmel_oracle as LM

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_logmel_vs_reference_golden(golden):
    from svc_inference_pipeline_b200.utils.mel import mel_spectrogram

    g = golden("logmel.npz")
    y = torch.from_numpy(g["wave"]).to(DEV)
    m = mel_spectrogram(y[None], 1024, 100, 24000, 256, 1024, 0, 12000, center=False).cpu().numpy()
    assert m.shape == g["logmel"].shape == (1, 100, 187)
    d = np.abs(m - g["logmel"])
    m64 = LM.mel_spectrogram(g["wave"], dtype=np.float64)
    print(f"log-mel vs the reference's fp32 output: max {d.max():.2e} mean {d.mean():.2e}; vs fp64 oracle: max {np.abs(m - m64).max():.2e} "
          f"(reference fp32 vs fp64 oracle: max {np.abs(g['logmel'] - m64).max():.2e})")
    assert d.max() < 2e-4 and d.mean() < 5e-6
    assert np.abs(m - m64).max() < 2e-4
    # the clip floor is exact
    floor = np.float32(np.log(np.float32(1e-5)))
    assert ((m == floor) == (g["logmel"] == floor)).mean() > 0.999


@pytest.mark.parametrize("cfg", [
    dict(n=24000, B=2, n_fft=1024, mels=100, sr=24000, hop=256, win=1024, fmin=0, fmax=12000),
    dict(n=4099, B=3, n_fft=1024, mels=100, sr=24000, hop=256, win=1024, fmin=0, fmax=12000),     # odd frame count, ragged tail
    dict(n=700, B=1, n_fft=1024, mels=100, sr=24000, hop=256, win=1024, fmin=0, fmax=12000),       # shorter than one window: all reflect
    dict(n=66150, B=2, n_fft=2048, mels=128, sr=44100, hop=512, win=2048, fmin=0, fmax=22050),     # the 512x generator's analysis
    dict(n=9000, B=1, n_fft=512, mels=80, sr=16000, hop=160, win=400, fmin=20, fmax=7600),         # window shorter than n_fft
])
def test_logmel_vs_oracle(cfg):
    from svc_inference_pipeline_b200.utils.mel import mel_spectrogram

    rng = np.random.default_rng(cfg["n"])
    t = np.arange(cfg["n"]) / cfg["sr"]
    wave = (0.4 * np.sin(2 * np.pi * 330 * t)[None] * rng.uniform(0.2, 1, (cfg["B"], 1)) + 0.05 * rng.standard_normal((cfg["B"], cfg["n"]))).astype(np.float32)
    args = (cfg["n_fft"], cfg["mels"], cfg["sr"], cfg["hop"], cfg["win"], cfg["fmin"], cfg["fmax"])
    m = mel_spectrogram(torch.from_numpy(wave).to(DEV), *args).cpu().numpy()
    ref = LM.mel_spectrogram(wave, *args, dtype=np.float64)
    assert m.shape == ref.shape
    assert np.abs(m - ref).max() < 2e-4, np.abs(m - ref).max()
    one = mel_spectrogram(torch.from_numpy(wave[0]).to(DEV), *args).cpu().numpy()  # 1-D input
    np.testing.assert_array_equal(one[0], m[0])


def test_logmel_errors_and_l1():
    from svc_inference_pipeline_b200 import _lib as L
    from svc_inference_pipeline_b200.utils.mel import log_mel_l1, mel_spectrogram

    y = torch.zeros(1, 300, device=DEV)
    with pytest.raises(L.BvgError):
        mel_spectrogram(y, 1024, 100, 24000, 256, 1024, 0, 12000)  # reflect pad (384) needs more than 300 samples
    with pytest.raises(RuntimeError):
        mel_spectrogram(torch.zeros(1, 4096), 1024, 100, 24000, 256, 1024, 0, 12000)  # CPU tensor: no fallback
    with pytest.raises(NotImplementedError):
        mel_spectrogram(torch.zeros(1, 4096, device=DEV), 1024, 100, 24000, 256, 1024, 0, 12000, center=True)
    rng = np.random.default_rng(0)
    a = (rng.standard_normal(24000) * 0.1).astype(np.float32)
    b = (a * 1.02).astype(np.float32)
    got = log_mel_l1(torch.from_numpy(a).to(DEV), torch.from_numpy(b).to(DEV))
    want = LM.log_mel_l1(a, b)
    assert abs(got - want) < 1e-5 and abs(got - np.log(1.02)) < 1e-3
