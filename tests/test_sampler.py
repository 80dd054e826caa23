"""Diffusion sampler (the caller of the DiffSVC denoiser step; reference modules/diffsvcrepo_inference.py): oracle vs the
unmodified reference function (CPU), and the B200 ``svc_model_inference`` vs both (``-m gpu``).  Goldens:
tests/golden/sampler.npz, written by make_golden.py::golden_sampler."""
import numpy as np
import pytest
import torch

from oracle import diffsvc_oracle as DO
from oracle import sampler_oracle as SO
from svc_inference_pipeline_b200.utils import synth
from svc_inference_pipeline_b200.utils.util import JsonHParams

MAPPER = dict(noise_schedule_factors=[0.0001, 0.02, 1000], n_mel=100, residual_channels=64, diffusion_fc_size=128, conditioner_size=64,
              dilation_cycle_length=4, residual_kernel_size=3, residual_layer_num=4)
DEV = "cuda:0"


def _sd():
    return {k: torch.from_numpy(v) for k, v in synth.synthetic_diffsvc_state_dict(MAPPER, seed=5).items()}


def _oracle_run(g, tag, dtype, fast):
    sd = _sd()
    cond = torch.from_numpy(g[tag + "_cond"]).to(dtype)
    denoise = lambda x, c, t: DO.denoiser_forward(sd, MAPPER, x, c, t)
    return SO.svc_model_inference(denoise, cond, g["noise_schedule"].tolist(), torch.from_numpy(g[tag + "_x0"]).to(dtype),
                                  torch.from_numpy(g[tag + "_noise"]), fast_inference=fast, speedup=3).numpy()


@pytest.mark.parametrize("tag", ["n2", "n1"])
def test_oracle_vs_reference(golden, tag):
    g = golden("sampler.npz")
    for fast, key in ((False, "_y"), (True, "_fast")):
        y32, y64 = _oracle_run(g, tag, torch.float32, fast), _oracle_run(g, tag, torch.float64, fast)
        ref = g[tag + key]
        assert y32.shape == ref.shape
        assert np.abs(y32 - ref).max() < 2e-5, (tag, fast)
        assert np.abs(y64 - ref).max() < 1e-4, (tag, fast)
    assert g["n1_y"].shape == (100, 96) and g["n2_y"].shape == (100, 61, 2)  # the reference's .T of [N, T, n_mel]


def test_schedule_tables_match_the_oracle():
    from svc_inference_pipeline_b200.modules.diffsvc import SCHEDULE_ROWS
    from svc_inference_pipeline_b200.modules.diffsvcrepo_inference import schedule_tables

    sched = np.linspace(1e-4, 0.02, 1000).tolist()
    t, o = schedule_tables(sched), SO.schedule(sched)
    names = dict(zip(SCHEDULE_ROWS, ("sqrt_recip", "sqrt_recipm1", "coef1", "coef2", "logvar", "alphas_cumprod")))
    assert t.shape == (6, 1000) and t.dtype == np.float32
    for k, row in enumerate(SCHEDULE_ROWS):
        np.testing.assert_array_equal(t[k], o[names[row]].numpy())
    assert t[4, 0] == np.float32(np.log(1e-20))  # posterior variance 0 at the first step, clipped
    with pytest.raises(ValueError):
        schedule_tables([])


def test_sampler_refuses_a_foreign_denoiser():
    from svc_inference_pipeline_b200.modules.diffsvcrepo_inference import svc_model_inference

    cfg = JsonHParams(mapper=JsonHParams(noise_schedule=[0.1, 0.2]))
    with pytest.raises(TypeError):
        svc_model_inference([lambda b: b["cond"], torch.nn.Identity()], {"y": torch.zeros(1, 8, 100), "cond": torch.zeros(1, 8, 64)}, cfg)


def _gpu_model(precision):
    from svc_inference_pipeline_b200.modules.diffsvc import DiffSVC

    m = DiffSVC(JsonHParams(**MAPPER), precision=precision)
    m.load_state_dict(_sd())
    return m.to(DEV).eval()


def _run(m, g, tag, fast, noise=True):
    from svc_inference_pipeline_b200.modules.diffsvcrepo_inference import svc_model_inference

    cfg = JsonHParams(mapper=JsonHParams(noise_schedule=g["noise_schedule"].tolist()))
    cond = torch.from_numpy(g[tag + "_cond"]).to(DEV)
    batch = {"y": torch.zeros(*g[tag + "_x0"].shape, device=DEV), "cond": cond}
    nz = {"x0": torch.from_numpy(g[tag + "_x0"]), "steps": torch.from_numpy(g[tag + "_noise"]).to(DEV)} if noise else None
    return svc_model_inference([lambda b: b["cond"], m], batch, cfg, fast_inference=fast, speedup=3, noise=nz)


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32_simt", "fp32", "bf16"])
@pytest.mark.parametrize("tag", ["n2", "n1"])
def test_sampler_vs_reference(golden, precision, tag):
    g = golden("sampler.npz")
    m = _gpu_model(precision)
    for fast, key in ((False, "_y"), (True, "_fast")):
        y = _run(m, g, tag, fast).cpu().numpy()
        ref = g[tag + key]
        ref64 = _oracle_run(g, tag, torch.float64, fast)
        assert y.shape == ref.shape
        err, err64 = float(np.abs(y - ref).max()), float(np.abs(y - ref64).max())
        print(f"sampler {tag} {precision} {'plms' if fast else 'ddpm'}: max-abs vs reference fp32 {err:.3e}, vs fp64 oracle {err64:.3e} "
              f"(reference fp32 vs fp64 {np.abs(ref - ref64).max():.1e}, |y|max {np.abs(ref).max():.2f})")
        if precision == "bf16":
            assert err64 < 0.1 * np.abs(ref).max()
        else:
            tol = (5e-5 if precision == "fp32_simt" else 2e-4) * max(1.0, np.abs(ref).max())
            assert err < tol and err64 < tol


@pytest.mark.gpu
def test_sampler_draws_like_the_reference_and_graph_equals_eager(golden):
    """Without ``noise`` the function draws x0 and the per-step noise with the reference's own calls, in its order: a
    seeded run equals a run fed with the tensors those calls return.  The CUDA-graph path equals the launch list."""
    g = golden("sampler.npz")
    m = _gpu_model("fp32")
    N, T, n_mel = g["n2_x0"].shape
    steps = len(g["noise_schedule"])
    torch.manual_seed(77)
    x0 = torch.normal(0, 1 / 1.2, size=(N, T, n_mel), device=DEV)
    nz = torch.zeros(steps, N, 1, n_mel, T, device=DEV)
    for i in reversed(range(steps)):
        nz[i] = torch.randn(N, 1, n_mel, T, device=DEV)
    from svc_inference_pipeline_b200.modules.diffsvcrepo_inference import svc_model_inference

    cfg = JsonHParams(mapper=JsonHParams(noise_schedule=g["noise_schedule"].tolist()))
    batch = {"y": torch.zeros(N, T, n_mel, device=DEV), "cond": torch.from_numpy(g["n2_cond"]).to(DEV)}
    model = [lambda b: b["cond"], m]
    y_fed = svc_model_inference(model, batch, cfg, noise={"x0": x0, "steps": nz})
    torch.manual_seed(77)
    y_seeded = svc_model_inference(model, batch, cfg)
    assert torch.equal(y_fed, y_seeded)
    m.use_cuda_graph = False
    y_eager = svc_model_inference(model, batch, cfg, noise={"x0": x0, "steps": nz})
    assert torch.equal(y_fed, y_eager)
    for fast in (True,):
        y_e = svc_model_inference(model, batch, cfg, fast_inference=True, speedup=3, noise={"x0": x0})
        m.use_cuda_graph = True
        y_g = svc_model_inference(model, batch, cfg, fast_inference=True, speedup=3, noise={"x0": x0})
        assert torch.equal(y_e, y_g)
    # the denoiser's own call surface still works between sampler runs (shared program, restored state)
    out, _ = m(torch.zeros(N, T, n_mel, device=DEV), batch["cond"], torch.zeros(N, 1, dtype=torch.long, device=DEV))
    assert torch.isfinite(out).all()
    with pytest.raises(IndexError):
        m.sampler(batch["cond"], np.zeros((6, 4), np.float32)).ddpm_step(4)
    with pytest.raises(ValueError):
        m.sampler(batch["cond"], np.zeros((6, 1001), np.float32))
