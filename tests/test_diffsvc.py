"""DiffSVC denoiser step: oracle vs the unmodified reference (CPU), and the B200 module vs both
(``-m gpu``).  Goldens: tests/golden/diffsvc.npz, written by make_golden.py from reference modules/diffsvc.py::DiffSVC."""
import numpy as np
import pytest
import torch

from oracle import diffsvc_oracle as DO
from svc_inference_pipeline_b200.utils import synth
from svc_inference_pipeline_b200.utils.util import JsonHParams
from util_cases import snr_db

MAPPER = dict(noise_schedule_factors=[0.0001, 0.02, 1000], n_mel=100, residual_channels=384, diffusion_fc_size=128, conditioner_size=384,
              dilation_cycle_length=4, residual_kernel_size=3, residual_layer_num=20)
DEV = "cuda:0"


def test_oracle_vs_reference(golden):
    g = golden("diffsvc.npz")
    sd = {k: torch.from_numpy(v) for k, v in synth.synthetic_diffsvc_state_dict(MAPPER, seed=3).items()}
    assert sum(v.numel() for v in sd.values()) == int(g["n_params"])
    for tag in ("b2", "utt"):
        mel, cond = torch.from_numpy(g[tag + "_mel"]), torch.from_numpy(g[tag + "_cond"])
        t = torch.from_numpy(g[tag + "_steps"]).long().unsqueeze(1)
        y32 = DO.denoiser_forward(sd, MAPPER, mel, cond, t).numpy()
        y64 = DO.denoiser_forward(sd, MAPPER, mel.double(), cond.double(), t).numpy()
        assert y32.shape == g[tag + "_y"].shape
        assert np.abs(y64 - g[tag + "_y_f64"]).max() < 1e-11
        assert np.abs(y32 - g[tag + "_y"]).max() < 2e-5
    np.testing.assert_array_equal(DO.build_embedding(1000).numpy(), __import__("svc_inference_pipeline_b200.modules.diffsvc", fromlist=["StepEncoder"]).StepEncoder.build_embedding(1000).numpy())


def test_module_matches_reference_grammar(golden):
    """Same state_dict keys / shapes / order as the reference module (digest recorded when the goldens were made); the
    final projection starts at zero like the reference's (modules/diffsvc.py:280); no CPU path."""
    import hashlib

    from svc_inference_pipeline_b200.modules.diffsvc import DiffSVC

    m = DiffSVC(JsonHParams(**MAPPER))
    sd = m.state_dict()
    digest = hashlib.sha256("\n".join(f"{k}:{tuple(v.shape)}" for k, v in sd.items()).encode()).digest()
    np.testing.assert_array_equal(np.frombuffer(digest, dtype=np.uint8), golden("diffsvc.npz")["keys_sha256"])
    assert list(sd.keys()) == list(synth.diffsvc_state_dict_spec(MAPPER).keys())
    assert sum(p.numel() for p in m.parameters()) == int(golden("diffsvc.npz")["n_params"])
    assert float(m.output_projection.weight.abs().max()) == 0.0
    assert "diffusion_embedding.embedding" not in sd and m.diffusion_embedding.embedding.shape == (1000, 128)
    assert len(m.noise_schedule) == 1000 and [rl.dilated_conv.dilation for rl in m.residual_layers][:5] == [1, 2, 4, 8, 1]
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 8, 100), torch.zeros(1, 8, 384), torch.zeros(1, 1, dtype=torch.long))


def _gpu_model(precision):
    from svc_inference_pipeline_b200.modules.diffsvc import DiffSVC

    m = DiffSVC(JsonHParams(**MAPPER), precision=precision)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.synthetic_diffsvc_state_dict(MAPPER, seed=3).items()})
    return m.to(DEV).eval()


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32_simt", "fp32", "bf16"])
def test_denoiser_step_vs_reference(golden, precision):
    g = golden("diffsvc.npz")
    m = _gpu_model(precision)
    for tag in ("b2", "utt"):
        mel, cond = torch.from_numpy(g[tag + "_mel"]).to(DEV), torch.from_numpy(g[tag + "_cond"]).to(DEV)
        t = torch.from_numpy(g[tag + "_steps"]).long().unsqueeze(1).to(DEV)
        y, stats = m(mel, cond, t)
        y = y.cpu().numpy()
        ref64, ref32 = g[tag + "_y_f64"], g[tag + "_y"]
        assert y.shape == ref64.shape and isinstance(stats, dict)
        err = float(np.abs(y - ref64).max())
        print(f"diffsvc step {tag} {precision}: max-abs vs reference fp64 {err:.3e} (|y|max {np.abs(ref64).max():.2f}, reference fp32 vs fp64 "
              f"{np.abs(ref32 - ref64).max():.1e}), SNR {snr_db(ref64, y):.1f} dB")
        if precision == "bf16":
            assert snr_db(ref64, y) > 35.0
        else:
            assert err < (2e-5 if precision == "fp32_simt" else 1e-4)
            assert float(np.abs(y - ref32).max()) < (2e-5 if precision == "fp32_simt" else 1e-4)


@pytest.mark.gpu
def test_denoiser_graph_steps_and_conditioner_cache(golden):
    """The captured CUDA graph follows the step index and the mel from call to call, equals the eager program bit for
    bit, and the cached conditioner projections are refreshed when the conditioner changes (new tensor, or the same
    tensor modified in place)."""
    g = golden("diffsvc.npz")
    m = _gpu_model("fp32")
    sd = {k: torch.from_numpy(v) for k, v in synth.synthetic_diffsvc_state_dict(MAPPER, seed=3).items()}
    mel, cond = torch.from_numpy(g["b2_mel"]).to(DEV), torch.from_numpy(g["b2_cond"]).to(DEV)
    assert m.use_cuda_graph

    def ref(mel_, cond_, steps):
        return DO.denoiser_forward(sd, MAPPER, mel_.cpu().double(), cond_.cpu().double(), torch.tensor(steps).unsqueeze(1)).numpy()

    for steps in ([17, 903], [0, 999], [500, 500]):
        y, _ = m(mel, cond, torch.tensor(steps, device=DEV).unsqueeze(1))
        assert np.abs(y.cpu().numpy() - ref(mel, cond, steps)).max() < 1e-4
    y_graph, _ = m(mel * 0.5, cond, torch.tensor([[3], [4]], device=DEV))
    m.use_cuda_graph = False
    y_eager, _ = m(mel * 0.5, cond, torch.tensor([[3], [4]], device=DEV))
    assert torch.equal(y_graph, y_eager)
    m.use_cuda_graph = True
    cond2 = cond * 0.7
    y2, _ = m(mel, cond2, torch.tensor([[17], [903]], device=DEV))
    assert np.abs(y2.cpu().numpy() - ref(mel, cond2, [17, 903])).max() < 1e-4
    cond2.mul_(0.5)  # in place: same object, new version
    y3, _ = m(mel, cond2, torch.tensor([[17], [903]], device=DEV))
    assert np.abs(y3.cpu().numpy() - ref(mel, cond2, [17, 903])).max() < 1e-4
    y4, _ = m(mel, cond2, torch.tensor([17, 903], device=DEV))  # [B] steps, cached conditioner
    assert torch.equal(y3, y4)
    # fractional steps: StepEncoder.lerp_embedding (modules/diffsvc.py:57-67)
    tf = torch.tensor([[17.25], [902.5]])
    y5, _ = m(mel, cond2, tf.to(DEV))
    ref5 = DO.denoiser_forward(sd, MAPPER, mel.cpu().double(), cond2.cpu().double(), tf.double()).numpy()
    assert np.abs(y5.cpu().numpy() - ref5).max() < 1e-4
    assert float((y5 - y4).abs().max()) > 1e-3
    assert m.launches_per_step(2, 61) == 2 + 1 + 20 * 4 + 3


def test_denoiser_loader_reads_the_mapper_checkpoint(tmp_path):
    """denoiser_model_loader takes the DiffSVC tensors (ModuleList index 1, reference utils/load_models.py:18-21) out of a
    mapper checkpoint saved the way svc_model_loader expects it ({"state_dict": ...}, optional module. prefix)."""
    import warnings

    from svc_inference_pipeline_b200.utils.load_models import denoiser_model_loader

    small = dict(MAPPER, residual_channels=16, conditioner_size=16, residual_layer_num=2, diffusion_fc_size=8)
    sd = synth.synthetic_diffsvc_state_dict(small, seed=1)
    ckpt = {"module.1." + k: torch.from_numpy(v) for k, v in sd.items()}
    ckpt["module.0.content_encoder.weight"] = torch.zeros(3, 3)              # the condition encoders: not ours
    ckpt["module.1.skip_projection.bias"] = torch.zeros(5)                   # wrong shape: dropped, reported
    path = str(tmp_path / "mapper.pt")
    torch.save({"state_dict": ckpt}, path)
    cfg = JsonHParams(device="cpu", svc_model_path=path, mapper=small)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        m = denoiser_model_loader(cfg)
    assert any("mismatched shape" in str(x.message) for x in w)
    assert not m.training
    rep = m.load_report
    assert rep["missing"] == ["skip_projection.bias"] and rep["unknown"] == [] and rep["other_modules"] == ["0"]
    assert [n for n, *_ in rep["wrong_shape"]] == ["skip_projection.bias"]
    got = m.state_dict()
    for k, v in sd.items():
        if k != "skip_projection.bias":
            np.testing.assert_array_equal(got[k].numpy(), v)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [dict(n_mel=80, residual_channels=512, conditioner_size=256, residual_layer_num=3, dilation_cycle_length=2, diffusion_fc_size=64),
                                   dict(n_mel=128, residual_channels=1032, conditioner_size=72, residual_layer_num=2, dilation_cycle_length=1, diffusion_fc_size=256,
                                        residual_kernel_size=5)])
@pytest.mark.parametrize("B,Ln", [(1, 77), (3, 1500)])
def test_denoiser_other_hyperparameters(shape, B, Ln):
    """Mapper configurations other than the reference's (channel counts that are not multiples of 64, 2C beyond the tile
    tables at the narrow cap, a 5-tap undilated kernel, row counts in every tile-width regime) against the fp64 oracle."""
    from svc_inference_pipeline_b200.modules.diffsvc import DiffSVC

    cfg = dict(MAPPER, **shape)
    sdn = synth.synthetic_diffsvc_state_dict(cfg, seed=11)
    sd = {k: torch.from_numpy(v) for k, v in sdn.items()}
    m = DiffSVC(JsonHParams(**cfg), precision="fp32")
    m.load_state_dict(sd)
    m = m.to(DEV).eval()
    g = torch.Generator().manual_seed(B * 1000 + Ln)
    mel, cond = torch.randn(B, Ln, cfg["n_mel"], generator=g), torch.randn(B, Ln, cfg["conditioner_size"], generator=g)
    t = torch.randint(0, 1000, (B, 1), generator=g)
    y, _ = m(mel.to(DEV), cond.to(DEV), t.to(DEV))
    ref = DO.denoiser_forward(sd, cfg, mel.double(), cond.double(), t).numpy()
    err = float(np.abs(y.cpu().numpy() - ref).max())
    print(f"diffsvc {shape['residual_channels']}ch B{B}x{Ln}: max-abs vs fp64 oracle {err:.3e} (|y|max {np.abs(ref).max():.2f})")
    assert err < 1e-4 * max(1.0, float(np.abs(ref).max()))
