"""Op-level parity of the CUDA kernels (through the C ABI) against the CPU oracle and the
reference-generated golden vectors.  Needs a B200: run with ``-m gpu``."""
import numpy as np
import pytest
import torch

from oracle import bigvgan_oracle as O
from util_cases import bf16_round

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    from svc_inference_pipeline_b200 import _lib as L
    from svc_inference_pipeline_b200 import ops as _ops

    L.check(L.lib().bvg_device_check(0), "device_check")
    return _ops, L


def cl(x):  # [B, C, L] numpy -> channels-last cuda tensor
    return torch.from_numpy(np.ascontiguousarray(np.transpose(x, (0, 2, 1)))).to(DEV)


def cf(t):  # channels-last cuda tensor -> [B, C, L] numpy
    return np.transpose(t.cpu().numpy(), (0, 2, 1))


def snake_params(alpha, beta, logscale):
    a = np.exp(alpha) if logscale else alpha
    b = a if beta is None else (np.exp(beta) if logscale else beta)
    return torch.from_numpy(a.astype(np.float32)).to(DEV), torch.from_numpy((1.0 / (b.astype(np.float32) + np.float32(1e-9))).astype(np.float32)).to(DEV)


# ------------------------------------------------------------------------------------------
# K-A fused Activation1d
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["snake", "snakebeta"])
@pytest.mark.parametrize("scale", ["lin", "log"])
@pytest.mark.parametrize("fast_sin", [False, True])
def test_amp_golden(ops, golden, name, scale, fast_sin):
    _ops, L = ops
    g = golden("activation1d.npz")
    f = golden("filters.npz")["aa12"]
    log = scale == "log"
    alpha = g["alpha"] if log else (1.0 + 0.3 * g["alpha"]).astype(np.float32)
    beta = None if name == "snake" else (g["beta"] if log else (1.0 + 0.3 * g["beta"]).astype(np.float32))
    a, invb = snake_params(alpha, beta, log)
    y = cf(_ops.activation1d(cl(g["x"]), a, invb, f, f, fast_sin=fast_sin))
    np.testing.assert_allclose(y, g[f"{name}_{scale}_a1d"], atol=8e-6, rtol=2e-6)
    np.testing.assert_allclose(y, g[f"{name}_{scale}_a1d_f64"], atol=8e-6, rtol=2e-6)


@pytest.mark.parametrize("ln", [1, 2, 3, 5, 6, 11, 12, 13])
def test_amp_edges(ops, golden, ln):
    _ops, L = ops
    g = golden("activation1d.npz")
    f = golden("filters.npz")["aa12"]
    a, invb = snake_params(g["alpha"][:3], g["beta"][:3], True)
    y = cf(_ops.activation1d(cl(g[f"edge{ln}_x"]), a, invb, f, f))
    np.testing.assert_allclose(y, g[f"edge{ln}_y"], atol=1e-5, rtol=2e-6)


def test_amp_large_argument(ops, golden):
    _ops, L = ops
    g = golden("activation1d.npz")
    f = golden("filters.npz")["aa12"]
    a, invb = snake_params(g["big_alpha"], g["big_beta"], True)
    ref64 = O.activation1d(g["big_x"].astype(np.float64), g["big_alpha"].astype(np.float64), g["big_beta"].astype(np.float64), True, f.astype(np.float64), f.astype(np.float64))
    np.testing.assert_allclose(ref64, g["big_y_f64"], atol=1e-12, rtol=1e-12)  # oracle == reference in fp64
    ref_noise = float(np.abs(g["big_y"] - g["big_y_f64"]).max())  # the reference's own fp32 vs its fp64: 1.7e-5 here
    for fast in (False, True):
        y = cf(_ops.activation1d(cl(g["big_x"]), a, invb, f, f, fast_sin=fast))
        err = float(np.abs(y - ref64).max())
        print(f"large-argument Activation1d (|a u| ~ 1e2, |y| <= 17): fast_sin={fast} max-abs vs fp64 {err:.3e}; reference fp32 vs fp64 {ref_noise:.3e}")
        # |a*u| reaches ~1e2: fp32 rounding of the product alone moves the phase by ~1e-5
        # measured on B200: 4.9e-5 with MUFU on the raw argument, 5.0e-5 with the exact reduction (3x the reference's
        # own fp32 noise on this input, the same for both: the raw-argument cosine costs no accuracy)
        assert err < 1e-4, (fast, err)
        assert np.abs(y - g["big_y"]).max() < 1e-4


@pytest.mark.parametrize("shape", [(2, 24, 1000), (1, 768, 301), (3, 6, 97), (2, 5, 64), (1, 48, 2500)])
@pytest.mark.parametrize("vec", [0, 1, 2, 4])
def test_amp_shapes_vs_oracle(ops, shape, vec):
    """Ragged lengths, every vector width, chunk boundaries (the time loop is chunked per thread)."""
    _ops, L = ops
    B, Ch, Ln = shape
    if vec and Ch % vec:
        pytest.skip("channel count not divisible")
    rng = np.random.default_rng(5)
    x = (rng.standard_normal(shape) * 1.5).astype(np.float32)
    alpha = (rng.standard_normal(Ch) * 0.3).astype(np.float32)
    beta = (rng.standard_normal(Ch) * 0.3).astype(np.float32)
    f = golden_taps()
    a, invb = snake_params(alpha, beta, True)
    ref = O.activation1d(x.astype(np.float64), alpha.astype(np.float64), beta.astype(np.float64), True, f.astype(np.float64), f.astype(np.float64))
    L.set_tuning("amp_vec", vec)
    try:
        for chunk in (0, 1, 2):
            L.set_tuning("amp_chunk", chunk)
            y = cf(_ops.activation1d(cl(x), a, invb, f, f))
            assert np.abs(y - ref).max() < 1e-5, (chunk, np.abs(y - ref).max())
    finally:
        L.set_tuning("amp_vec", 0)
        L.set_tuning("amp_chunk", 0)


def golden_taps():
    from svc_inference_pipeline_b200.utils import synth

    return synth.aa_filter_taps()


@pytest.mark.parametrize("in_dt,out_dt", [(0, 1), (0, 2), (1, 0), (1, 1), (1, 2)])
def test_amp_formats(ops, in_dt, out_dt):
    """bf16 / split element formats: the kernel computes in fp32 on the values it reads, and the
    stored value is the rounding of the fp32 result."""
    _ops, L = ops
    rng = np.random.default_rng(6)
    x = (rng.standard_normal((2, 48, 517)) * 1.5).astype(np.float32)
    alpha = (rng.standard_normal(48) * 0.3).astype(np.float32)
    beta = (rng.standard_normal(48) * 0.3).astype(np.float32)
    f = golden_taps()
    a, invb = snake_params(alpha, beta, True)
    x_seen = bf16_round(x) if in_dt == L.BF16 else x
    ref = O.activation1d(x_seen.astype(np.float64), alpha.astype(np.float64), beta.astype(np.float64), True, f.astype(np.float64), f.astype(np.float64))
    y = cf(_ops.activation1d(cl(x), a, invb, f, f, in_dtype=in_dt, out_dtype=out_dt, fast_sin=True))
    if out_dt == L.BF16:
        # BF16 -> BF16 runs on the tensor-core kernel, which also rounds the activated 2x-rate signal to
        # bf16 before the low-pass: up to ~1 bf16 ulp of the largest value instead of 1/2
        tol = 2**-7 if in_dt == L.BF16 else 2**-8
        assert np.abs(y - ref).max() <= np.abs(ref).max() * tol + 1e-5
        np.testing.assert_array_equal(y, bf16_round(y))
    elif out_dt == L.SPLIT:
        assert np.abs(y - ref).max() <= np.abs(ref).max() * 2**-15 + 1e-5
    else:
        assert np.abs(y - ref).max() < 2e-5


@pytest.mark.parametrize("in_dt,out_dt", [(0, 0), (0, 2), (1, 1), (1, 2), (0, 1)])
@pytest.mark.parametrize("fast_sin", [False, True])
def test_amp_packed_equals_scalar(ops, in_dt, out_dt, fast_sin):
    """The FFMA2 (fma.rn.f32x2) kernel performs the scalar kernel's fp32 operations two channels at a
    time: outputs must be bit-identical, including both replicate clamps, ragged chunk ends and the clamp-free
    fast path of interior chunks; the per-C instantiations (compile-time channel count) must equal the
    runtime-C kernel."""
    _ops, L = ops
    rng = np.random.default_rng(11)
    for shape in [(2, 24, 1000), (1, 768, 301), (3, 6, 97), (1, 48, 13), (1, 2, 5), (1, 96, 200), (2, 24, 95), (1, 24, 1), (1, 24, 2), (1, 24, 3), (1, 24, 7)]:
        Ch = shape[1]
        x = (rng.standard_normal(shape) * 1.5).astype(np.float32)
        a, invb = snake_params((rng.standard_normal(Ch) * 0.3).astype(np.float32), (rng.standard_normal(Ch) * 0.3).astype(np.float32), True)
        f = golden_taps()
        for chunk in (0, 1):
            L.set_tuning("amp_mma", 0)
            L.set_tuning("amp_stream", 0)
            L.set_tuning("amp_chunk", chunk)
            try:
                y_packed = _ops.activation1d(cl(x), a, invb, f, f, in_dtype=in_dt, out_dtype=out_dt, fast_sin=fast_sin).cpu().numpy()
                L.set_tuning("amp_ct", 0)  # runtime channel count instead of the per-C instantiation
                y_packed_rt = _ops.activation1d(cl(x), a, invb, f, f, in_dtype=in_dt, out_dtype=out_dt, fast_sin=fast_sin).cpu().numpy()
                L.set_tuning("amp_packed", 0)
                y_scalar = _ops.activation1d(cl(x), a, invb, f, f, in_dtype=in_dt, out_dtype=out_dt, fast_sin=fast_sin).cpu().numpy()
            finally:
                L.set_tuning("amp_packed", 1)
                L.set_tuning("amp_ct", 1)
                L.set_tuning("amp_mma", 1)
                L.set_tuning("amp_stream", 0)
                L.set_tuning("amp_chunk", 0)
            np.testing.assert_array_equal(y_packed, y_packed_rt)
            if fast_sin:
                np.testing.assert_array_equal(y_packed, y_scalar)
            else:
                # the scalar kernel's range reduction lets nvcc contract u*a + magic into one FFMA; the packed one
                # rounds the product first: the reduced phase can differ by one rounding (never a parity flip of sin^2)
                tol = {L.F32: 2e-6, L.SPLIT: 2**-14, L.BF16: 2**-7}[out_dt]  # one fp32 ulp can move a bf16 / split rounding
                np.testing.assert_allclose(y_packed, y_scalar, atol=tol * max(1.0, float(np.abs(y_scalar).max())), rtol=0)


AMP_MMA_SHAPES = [(2, 24, 1000), (1, 768, 301), (3, 8, 97), (2, 40, 64), (1, 48, 2500), (2, 96, 133), (1, 16, 4), (1, 64, 259), (2, 112, 515)]


@pytest.mark.parametrize("shape", AMP_MMA_SHAPES + [(1, 16, ln) for ln in (1, 2, 3, 5, 7, 8, 9, 11, 12, 13, 15, 16, 17, 19, 20, 21, 27, 28)])
@pytest.mark.parametrize("mode", ["f32_split", "bf16_bf16", "f32_split_stream", "bf16_bf16_stream"])
@pytest.mark.parametrize("fast_sin", [False, True])
def test_amp_mma_vs_oracle(ops, shape, mode, fast_sin):
    """Tensor-core Activation1d (amp_mma.cu: both FIRs as banded-Toeplitz MMAs) on the two operand
    formats the generator uses, ragged lengths (both replicate clamps, partial tiles), channel counts
    that leave 16-channel groups partly empty; checked against the fp64 oracle and against the FFMA
    kernel (amp_kernel.cu) on the same inputs."""
    _ops, L = ops
    B, Ch, Ln = shape
    rng = np.random.default_rng(7 + Ch + Ln)
    x = (rng.standard_normal(shape) * 1.5).astype(np.float32)
    alpha = (rng.standard_normal(Ch) * 0.3).astype(np.float32)
    beta = (rng.standard_normal(Ch) * 0.3).astype(np.float32)
    f = golden_taps()
    a, invb = snake_params(alpha, beta, True)
    stream = mode.endswith("_stream")  # amp_stream.cu: the per-warp cp.async ring variant of either format
    in_dt, out_dt = (L.F32, L.SPLIT) if mode.startswith("f32_split") else (L.BF16, L.BF16)
    x_seen = bf16_round(x) if in_dt == L.BF16 else x
    ref = O.activation1d(x_seen.astype(np.float64), alpha.astype(np.float64), beta.astype(np.float64), True, f.astype(np.float64), f.astype(np.float64))
    scale = np.abs(ref).max()
    try:
        L.set_tuning("amp_mma", 0 if stream else 2)  # force the tensor-core kernel under test for every supported shape
        L.set_tuning("amp_stream", 1 if stream else 0)
        L.set_tuning("amp_stream_bf16", 1 if stream else 0)
        y = cf(_ops.activation1d(cl(x), a, invb, f, f, in_dtype=in_dt, out_dtype=out_dt, fast_sin=fast_sin))
        L.set_tuning("amp_mma", 0)
        L.set_tuning("amp_stream", 0)
        L.set_tuning("amp_stream_bf16", 0)
        y_ffma = cf(_ops.activation1d(cl(x), a, invb, f, f, in_dtype=in_dt, out_dtype=out_dt, fast_sin=fast_sin))
    finally:
        L.set_tuning("amp_mma", 1)
        L.set_tuning("amp_stream", 0)
        L.set_tuning("amp_stream_bf16", 0)
    assert np.isfinite(y).all()
    if out_dt == L.SPLIT:
        # operands carry 16 mantissa bits (hi + lo), products accumulate in fp32
        assert np.abs(y - ref).max() <= scale * 2**-14 + 1e-5, np.abs(y - ref).max()
        assert np.abs(y - y_ffma).max() <= scale * 2**-14 + 1e-5
    else:
        # s enters the low-pass MMA as fp16 (11 significant bits; the FFMA kernel keeps it in fp32) and the result is bf16
        assert np.abs(y - ref).max() <= scale * 2**-7 + 1e-5, np.abs(y - ref).max()
        np.testing.assert_array_equal(y, bf16_round(y))
        rel = np.sqrt(((y - ref) ** 2).sum() / (ref**2).sum())
        rel_ffma = np.sqrt(((y_ffma - ref) ** 2).sum() / (ref**2).sum())
        assert rel < 2**-8.0 and rel < 1.6 * rel_ffma + 1e-4, (rel, rel_ffma)


# ------------------------------------------------------------------------------------------
# K-C / K-T convolutions
# ------------------------------------------------------------------------------------------
def _conv_case(golden, tag, transposed):
    g = golden("convs.npz")
    return {k: g[f"{tag}_{k}"] for k in ("x", "v", "g", "b", "y", "w", "args")}


def _run_conv(ops, c, transposed, backend, split, x_override=None):
    _ops, L = ops
    args = [int(a) for a in c["args"]]
    if transposed:
        cin, cout, k, u = args
        kw = dict(transposed=True, stride=u, padding=(k - u) // 2)
    else:
        ch, k, d = args
        cin = cout = ch
        kw = dict(dilation=d, padding=O.get_padding(k, d))
    v, gg, b = (torch.from_numpy(c[n]).to(DEV) for n in ("v", "g", "b"))
    pc = _ops.pack_conv(v, gg, b, backend=backend, split=split, **kw)
    x = c["x"] if x_override is None else x_override
    xc = cl(x)
    if xc.shape[-1] != pc.x_pitch:  # zero-pad channels up to the pitch the backend expects
        xc = torch.nn.functional.pad(xc, (0, pc.x_pitch - xc.shape[-1]))
    y = _ops.conv(xc.contiguous(), pc)
    B, Ln, n = y.shape
    if transposed:
        y = y.reshape(B, Ln * u, cout)
    return cf(y)


@pytest.mark.parametrize("tag", ["c8k3d1", "c8k7d3", "c6k11d5", "c4k11d5_short"])
def test_conv_simt_golden(ops, golden, tag):
    c = _conv_case(golden, tag, False)
    y = _run_conv(ops, c, False, 0, False)
    np.testing.assert_allclose(y, c["y"], atol=5e-6, rtol=1e-5)


@pytest.mark.parametrize("tag", ["t8to4k8u4", "t6to3k4u2", "t4to2k16u8", "t4to2k4u2_len1"])
def test_convT_simt_golden(ops, golden, tag):
    c = _conv_case(golden, tag, True)
    y = _run_conv(ops, c, True, 0, False)
    assert y.shape == c["y"].shape
    np.testing.assert_allclose(y, c["y"], atol=5e-6, rtol=1e-5)


def _oracle_conv(x, v, g, b, transposed, k, d=1, u=1):
    w = O.weight_norm_fold(v.astype(np.float64), g.astype(np.float64))
    if transposed:
        return O.conv_transpose1d(x.astype(np.float64), w, b.astype(np.float64), u, (k - u) // 2), w
    return O.conv1d(x.astype(np.float64), w, b.astype(np.float64), d, O.get_padding(k, d)), w


UMMA_CONV_CASES = [
    # (B, C, L, k, d)
    (2, 64, 300, 3, 1),
    (1, 128, 517, 7, 3),
    (2, 192, 260, 11, 5),
    (1, 96, 400, 11, 1),
    (2, 48, 333, 7, 5),
    (3, 24, 1000, 11, 5),
    (1, 24, 50, 3, 1),
    (1, 768, 140, 3, 1),
    (1, 384, 129, 7, 1),
    (2, 8, 77, 3, 3),
]


def _umma_conv_check(ops, B, Ch, Ln, k, d, split, mb):
    _ops, L = ops
    rng = np.random.default_rng(100 + Ch + k + d)
    x = rng.standard_normal((B, Ch, Ln)).astype(np.float32)
    v = (rng.standard_normal((Ch, Ch, k)) / np.sqrt(Ch * k)).astype(np.float32)
    g = (np.sqrt((v.astype(np.float64) ** 2).sum(axis=(1, 2), keepdims=True)) * rng.uniform(0.7, 1.4, (Ch, 1, 1))).astype(np.float32)
    b = (rng.standard_normal(Ch) * 0.1).astype(np.float32)
    L.set_tuning("umma_mb", mb)
    try:
        pc = _ops.pack_conv(*(torch.from_numpy(t).to(DEV) for t in (v, g, b)), dilation=d, padding=O.get_padding(k, d), backend=L.UMMA, split=split)
        if mb * ((pc.desc.n_tile + 31) // 32 * 32) > 512:
            pytest.skip("accumulators of this many M blocks do not fit in TMEM")
        y = cf(_ops.conv(cl(x), pc))
    finally:
        L.set_tuning("umma_mb", 0)
    ref, w = _oracle_conv(x, v, g, b, False, k, d=d)
    if split:
        err = np.abs(y - ref).max()
        assert err < 3e-5 * max(1.0, np.abs(ref).max()), f"split err {err}"
    else:
        # exact emulation: bf16 operands, wide accumulation
        wq = bf16_round(w.astype(np.float32)).astype(np.float64)
        refq = O.conv1d(bf16_round(x).astype(np.float64), wq, b.astype(np.float64), d, O.get_padding(k, d))
        err = np.abs(y - refq).max()
        assert err < 2e-3, f"bf16 err vs bf16-emulated oracle {err}"
        assert np.abs(y - ref).max() < 0.05 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("case", UMMA_CONV_CASES)
@pytest.mark.parametrize("mb", [0, 1, 2, 4])
def test_conv_umma_bf16(ops, case, mb):
    """bf16 operands; mb = M blocks per tile (0 = the host heuristic).  Every tap reads one shared A
    halo tile through row-shifted UMMA descriptors."""
    _umma_conv_check(ops, *case, split=False, mb=mb)


@pytest.mark.parametrize("case", UMMA_CONV_CASES[:7])
@pytest.mark.parametrize("mb", [0, 1, 2])
def test_conv_umma_split(ops, case, mb):
    _umma_conv_check(ops, *case, split=True, mb=mb)


FOLD_CASES = [
    # (B, C, L, k, d, fold)
    (2, 24, 1000, 11, 1, 4),
    (1, 24, 516, 7, 1, 4),
    (3, 24, 64, 3, 1, 4),
    (1, 24, 2048, 11, 3, 4),
    (2, 24, 400, 7, 5, 4),
    (2, 48, 334, 11, 1, 2),
    (1, 48, 600, 7, 1, 2),
    (1, 8, 96, 11, 1, 8),
    (1, 24, 4, 11, 1, 4),
    (1, 24, 8, 7, 1, 4),
]


@pytest.mark.parametrize("case", FOLD_CASES)
@pytest.mark.parametrize("split", [False, True])
def test_conv_umma_time_fold(ops, case, split):
    """bvg_conv_geom.fold: P rows read as one row of P*C channels and the taps rearranged block-Toeplitz give the
    same Conv1d -- against the fp64 oracle on the unfolded layer (zero padding at both ends, batch items kept apart,
    dilations whose folded form has all-zero blocks) and against the unfolded tensor-core layer itself."""
    _ops, L = ops
    B, Ch, Ln, k, d, fold = case
    rng = np.random.default_rng(300 + Ch + k + d + Ln)
    x = rng.standard_normal((B, Ch, Ln)).astype(np.float32)
    v = (rng.standard_normal((Ch, Ch, k)) / np.sqrt(Ch * k)).astype(np.float32)
    g = (np.sqrt((v.astype(np.float64) ** 2).sum(axis=(1, 2), keepdims=True)) * rng.uniform(0.7, 1.4, (Ch, 1, 1))).astype(np.float32)
    b = (rng.standard_normal(Ch) * 0.1).astype(np.float32)
    wts = [torch.from_numpy(t).to(DEV) for t in (v, g, b)]
    pc_f = _ops.pack_conv(*wts, dilation=d, padding=O.get_padding(k, d), backend=L.UMMA, split=split, fold=fold)
    pc_u = _ops.pack_conv(*wts, dilation=d, padding=O.get_padding(k, d), backend=L.UMMA, split=split)
    assert pc_f.x_pitch == fold * Ch and pc_f.n_total == fold * Ch
    xl = cl(x)                                             # [B, L, C]
    y_f = _ops.conv(xl.reshape(B, Ln // fold, fold * Ch), pc_f).reshape(B, Ln, Ch)
    y_u = _ops.conv(xl, pc_u)
    ref, _ = _oracle_conv(x, v, g, b, False, k, d=d)
    yf, yu = cf(y_f), cf(y_u)
    scale = max(1.0, np.abs(ref).max())
    if split:
        assert np.abs(yf - ref).max() < 3e-5 * scale
        assert np.abs(yf - yu).max() < 3e-5 * scale
    else:
        assert np.abs(yf - ref).max() < 0.05 * scale
        assert np.abs(yf - yu).max() < 2e-3  # same bf16 operands, different accumulation order


FUSED_CASES = [
    # (B, C, L, k, d)
    (2, 48, 1000, 11, 1),
    (1, 48, 517, 7, 3),
    (3, 48, 260, 3, 5),
    (2, 24, 1500, 11, 5),
    (1, 24, 333, 3, 1),
    (1, 64, 700, 7, 1),
    (1, 8, 200, 7, 5),
    (2, 40, 129, 11, 3),
    (1, 48, 5, 7, 1),
    (1, 24, 1, 3, 1),
    (4, 48, 4100, 11, 1),
]


@pytest.mark.parametrize("case", FUSED_CASES)
@pytest.mark.parametrize("epi", ["plain", "res", "res_acc"])
@pytest.mark.parametrize("fast_sin", [True, False])
def test_conv_fused_activation1d(ops, case, epi, fast_sin):
    """bvg_conv_desc.pre_amp: the convolution computes its (hi, lo) operand from the Activation1d's fp32 input inside
    the kernel.  Same arithmetic as amp_kernel_p2 followed by the unfused convolution, so the results must agree
    to rounding -- sequence ends (both replicate clamps next to the conv's zero padding), ragged tiles, several
    tiles per CTA, channel counts that leave pad columns (C = 24, 40, 8), every resblock epilogue."""
    _ops, L = ops
    B, Ch, Ln, k, d = case
    rng = np.random.default_rng(500 + Ch + k + d + Ln)
    x = (rng.standard_normal((B, Ch, Ln)) * 1.5).astype(np.float32)
    v = (rng.standard_normal((Ch, Ch, k)) / np.sqrt(Ch * k)).astype(np.float32)
    g = (np.sqrt((v.astype(np.float64) ** 2).sum(axis=(1, 2), keepdims=True)) * rng.uniform(0.7, 1.4, (Ch, 1, 1))).astype(np.float32)
    b = (rng.standard_normal(Ch) * 0.1).astype(np.float32)
    a, invb = snake_params((rng.standard_normal(Ch) * 0.3).astype(np.float32), (rng.standard_normal(Ch) * 0.3).astype(np.float32), True)
    f = golden_taps()
    pc = _ops.pack_conv(*(torch.from_numpy(t).to(DEV) for t in (v, g, b)), dilation=d, padding=O.get_padding(k, d), backend=L.UMMA, split=True)
    res = cl(rng.standard_normal((B, Ch, Ln)).astype(np.float32)) if epi != "plain" else None
    acc = cl(rng.standard_normal((B, Ch, Ln)).astype(np.float32)) if epi == "res_acc" else None
    kw = dict(res=res, acc=acc, div=3.0 if epi == "res_acc" else 1.0)
    xl = cl(x)
    L.set_tuning("amp_mma", 0)
    try:
        z = _ops.activation1d(xl, a, invb, f, f, in_dtype=L.F32, out_dtype=L.SPLIT, fast_sin=fast_sin)
    finally:
        L.set_tuning("amp_mma", 1)
    y_ref = _ops.conv(z, pc, **kw)
    y_fused = _ops.conv(xl, pc, pre_amp=(a, invb, f, f, fast_sin), **kw)
    assert torch.isfinite(y_fused).all()
    # (the unfused reference goes through a float32 copy of the (hi, lo) planes, whose sum can round away the last
    # bits of a tiny lo term: agreement to fp32 rounding instead of bit equality)
    err = float((y_fused - y_ref).abs().max())
    assert err <= 1e-5 * max(1.0, float(y_ref.abs().max())), err


def test_conv_umma_large_rows(ops):
    """Many tiles per CTA (pipeline wrap-around of every barrier ring) and ragged last tiles."""
    _umma_conv_check(ops, 4, 48, 20011, 7, 3, split=False, mb=0)
    _umma_conv_check(ops, 2, 192, 9001, 3, 1, split=True, mb=0)


@pytest.mark.parametrize("cin,cout,k,u,Ln", [(64, 32, 8, 4, 50), (128, 64, 4, 2, 333), (48, 24, 4, 2, 200), (256, 128, 16, 8, 40), (1536, 768, 8, 4, 20)])
@pytest.mark.parametrize("split", [False, True])
def test_convT_umma(ops, cin, cout, k, u, Ln, split):
    _ops, L = ops
    rng = np.random.default_rng(cin + k)
    x = rng.standard_normal((2, cin, Ln)).astype(np.float32)
    v = (rng.standard_normal((cin, cout, k)) / np.sqrt(cin * k / u)).astype(np.float32)
    g = (np.sqrt((v.astype(np.float64) ** 2).sum(axis=(1, 2), keepdims=True)) * rng.uniform(0.7, 1.4, (cin, 1, 1))).astype(np.float32)
    b = (rng.standard_normal(cout) * 0.1).astype(np.float32)
    pc = _ops.pack_conv(*(torch.from_numpy(t).to(DEV) for t in (v, g, b)), transposed=True, stride=u, padding=(k - u) // 2, backend=L.UMMA, split=split)
    y = _ops.conv(cl(x), pc)
    y = cf(y.reshape(2, Ln * u, cout))
    ref, w = _oracle_conv(x, v, g, b, True, k, u=u)
    assert y.shape == ref.shape
    if split:
        assert np.abs(y - ref).max() < 3e-5 * max(1.0, np.abs(ref).max())
    else:
        wq = bf16_round(w.astype(np.float32)).astype(np.float64)
        refq = O.conv_transpose1d(bf16_round(x).astype(np.float64), wq, b.astype(np.float64), u, (k - u) // 2)
        assert np.abs(y - refq).max() < 2e-3


def test_conv_epilogue_variants(ops):
    """residual add, running-sum add, division and every output format, on both backends."""
    _ops, L = ops
    rng = np.random.default_rng(9)
    B, Ch, Ln, k, d = 2, 64, 203, 7, 3
    x = rng.standard_normal((B, Ch, Ln)).astype(np.float32)
    res = rng.standard_normal((B, Ch, Ln)).astype(np.float32)
    acc = rng.standard_normal((B, Ch, Ln)).astype(np.float32)
    v = (rng.standard_normal((Ch, Ch, k)) / np.sqrt(Ch * k)).astype(np.float32)
    g = np.sqrt((v.astype(np.float64) ** 2).sum(axis=(1, 2), keepdims=True)).astype(np.float32)
    b = (rng.standard_normal(Ch) * 0.1).astype(np.float32)
    ref, _ = _oracle_conv(x, v, g, b, False, k, d=d)
    want = (ref + res + acc) / 3.0
    for backend, split, tol in ((L.SIMT, False, 1e-5), (L.UMMA, True, 5e-5)):
        pc = _ops.pack_conv(*(torch.from_numpy(t).to(DEV) for t in (v, g, b)), dilation=d, padding=O.get_padding(k, d), backend=backend, split=split)
        for out_dt, otol in ((L.F32, 0.0), (L.BF16, 2**-8), (L.SPLIT, 2**-15)):
            y = cf(_ops.conv(cl(x), pc, out_dtype=out_dt, res=cl(res), acc=cl(acc), div=3.0))
            assert np.abs(y - want).max() < tol + otol * np.abs(want).max(), (backend, out_dt, np.abs(y - want).max())
        y = cf(_ops.conv(cl(x), pc, res=cl(res), res_dtype=L.BF16))
        assert np.abs(y - (ref + bf16_round(res))).max() < tol


# ------------------------------------------------------------------------------------------
# head / tail
# ------------------------------------------------------------------------------------------
def test_pack_mel(ops):
    _ops, L = ops
    rng = np.random.default_rng(3)
    mel = rng.standard_normal((3, 100, 77)).astype(np.float32)
    for dt, tol in ((L.F32, 0), (L.BF16, 2**-8), (L.SPLIT, 2**-15)):
        y = _ops.pack_mel(torch.from_numpy(mel).to(DEV), 104, dt).cpu().numpy()
        assert y.shape == (3, 77, 104)
        assert np.abs(y[:, :, :100] - np.transpose(mel, (0, 2, 1))).max() <= tol * np.abs(mel).max()
        assert np.all(y[:, :, 100:] == 0)


def test_post(ops):
    _ops, L = ops
    rng = np.random.default_rng(4)
    x = rng.standard_normal((2, 24, 999)).astype(np.float32)
    v = (rng.standard_normal((1, 24, 7)) * 0.1).astype(np.float32)
    g = np.array([[[1.7]]], dtype=np.float32)
    b = np.array([0.05], dtype=np.float32)
    w = O.weight_norm_fold(v.astype(np.float64), g.astype(np.float64))
    ref = np.tanh(O.conv1d(x.astype(np.float64), w, b.astype(np.float64), 1, 3))[:, 0]
    y = _ops.post(cl(x), *(torch.from_numpy(t).to(DEV) for t in (v, g, b))).cpu().numpy()
    assert np.abs(y - ref).max() < 2e-6


# ------------------------------------------------------------------------------------------
# guard-band checks (compute-sanitizer is closed on this pool): every kernel that masks partial
# tiles writes into a buffer surrounded by sentinels, which must survive
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kernel", ["mma", "stream", "packed", "scalar"])
@pytest.mark.parametrize("mode", ["f32_split", "bf16_bf16", "f32_f32"])
def test_amp_guard_bands(ops, kernel, mode):
    import ctypes as C

    _ops, L = ops
    if kernel == "mma" and mode == "f32_f32":
        pytest.skip("the tensor-core kernel does not take F32 -> F32")
    if kernel == "stream" and mode == "f32_f32":
        pytest.skip("the streaming kernel does not take F32 -> F32")
    in_dt, out_dt = {"f32_split": (L.F32, L.SPLIT), "bf16_bf16": (L.BF16, L.BF16), "f32_f32": (L.F32, L.F32)}[mode]
    G = 4096
    f = golden_taps()
    rng = np.random.default_rng(9)
    for (B, Ch, Ln) in [(2, 24, 61), (1, 40, 130), (3, 8, 7), (1, 112, 257), (2, 48, 64)]:
        n = B * Ln * Ch
        x = torch.from_numpy((rng.standard_normal(n) * 1.5).astype(np.float32)).to(DEV)
        xin = x if in_dt == L.F32 else x.to(torch.bfloat16)
        odt = torch.float32 if out_dt == L.F32 else torch.bfloat16
        sent = 12345.0
        planes = [torch.full((n + 2 * G,), sent, dtype=odt, device=DEV) for _ in range(2 if out_dt == L.SPLIT else 1)]
        a, invb = snake_params((rng.standard_normal(Ch) * 0.3).astype(np.float32), (rng.standard_normal(Ch) * 0.3).astype(np.float32), True)
        d = L.AmpDesc()
        d.x = L.Tensor(xin.data_ptr(), None, in_dt, 0)
        esz = planes[0].element_size()
        d.y = L.Tensor(planes[0].data_ptr() + G * esz, (planes[1].data_ptr() + G * esz) if len(planes) > 1 else None, out_dt, 0)
        d.d_a, d.d_invb = a.data_ptr(), invb.data_ptr()
        d.taps_up = (C.c_float * 12)(*f.tolist())
        d.taps_down = (C.c_float * 12)(*f.tolist())
        d.B, d.L, d.C, d.fast_sin = B, Ln, Ch, 1
        try:
            L.set_tuning("amp_mma", 2 if kernel == "mma" else 0)
            L.set_tuning("amp_stream", 1 if kernel == "stream" else 0)
            L.set_tuning("amp_stream_bf16", 1 if kernel == "stream" else 0)
            L.set_tuning("amp_packed", 0 if kernel == "scalar" else 1)
            tp = L.tuning_ptr()  # the knobs travel with the descriptor (bvg_amp_desc.tune)
            if tp is not None:
                d.tune = tp
            L.check(L.lib().bvg_amp_fwd(C.byref(d), torch.cuda.current_stream().cuda_stream), "amp_fwd")
            torch.cuda.synchronize()
        finally:
            L.set_tuning("amp_mma", 1)
            L.set_tuning("amp_stream", 0)
            L.set_tuning("amp_stream_bf16", 0)
            L.set_tuning("amp_packed", 1)
        for pl in planes:
            assert bool((pl[:G] == sent).all()) and bool((pl[-G:] == sent).all()), (kernel, mode, B, Ch, Ln)
            assert bool((pl[G:-G] != sent).all())  # and every element inside was written


def test_tail_and_pack_guard_bands(ops):
    import ctypes as C

    _ops, L = ops
    G = 1024
    rng = np.random.default_rng(10)
    B, Ln, sil, fade = 3, 5000, 37, 640
    wave = torch.from_numpy((rng.standard_normal(B * Ln) * 0.1).astype(np.float32)).to(DEV)
    n = B * (Ln + 2 * sil)
    pcm = torch.full((n + 2 * G,), 777, dtype=torch.int16, device=DEV)
    peak = torch.full((B + 2,), -1.0, dtype=torch.float32, device=DEV)
    d = L.TailDesc()
    d.d_wave, d.d_pcm, d.d_peak = wave.data_ptr(), pcm.data_ptr() + 2 * G, peak.data_ptr() + 4
    d.B, d.L, d.fade_len, d.silence, d.volume_peak = B, Ln, fade, sil, 0.9
    L.check(L.lib().bvg_tail_fwd(C.byref(d), torch.cuda.current_stream().cuda_stream), "tail_fwd")
    torch.cuda.synchronize()
    assert bool((pcm[:G] == 777).all()) and bool((pcm[-G:] == 777).all())
    assert float(peak[0]) == -1.0 and float(peak[-1]) == -1.0 and bool((peak[1:-1] > 0).all())
    got = pcm[G:-G].cpu().numpy().reshape(B, -1)
    ref = np.stack([O.synthesis_pcm16(w, fade // 20, sil * 20) for w in wave.cpu().numpy().reshape(B, Ln)])
    np.testing.assert_array_equal(got, ref)
    # head: [B, C, T] -> [B, T, c_pad] with zero-filled pad channels
    Bm, Cm, Tm, cpad = 2, 100, 45, 104
    mel = torch.from_numpy(rng.standard_normal((Bm, Cm, Tm)).astype(np.float32)).to(DEV)
    out = torch.full((Bm * Tm * cpad + 2 * G,), 5.0, dtype=torch.float32, device=DEV)
    pd = L.PackDesc()
    pd.d_mel = mel.data_ptr()
    pd.out = L.Tensor(out.data_ptr() + 4 * G, None, L.F32, 0)
    pd.B, pd.C, pd.T, pd.c_pad = Bm, Cm, Tm, cpad
    L.check(L.lib().bvg_pack_mel(C.byref(pd), torch.cuda.current_stream().cuda_stream), "pack_mel")
    torch.cuda.synchronize()
    assert bool((out[:G] == 5.0).all()) and bool((out[-G:] == 5.0).all())
    body = out[G:-G].reshape(Bm, Tm, cpad)
    assert torch.equal(body[:, :, :Cm], mel.permute(0, 2, 1)) and bool((body[:, :, Cm:] == 0).all())


@pytest.mark.parametrize("backend", ["umma_split", "umma_bf16", "simt"])
def test_conv_guard_bands(ops, backend):
    """Convolution epilogues (partial M tiles, N tiles wider than Cout, transposed convs) never write
    outside [B, L, n_total]."""
    import ctypes as C

    _ops, L = ops
    G = 8192
    rng = np.random.default_rng(12)
    be = L.SIMT if backend == "simt" else L.UMMA
    split = backend == "umma_split"
    x_dt = L.F32 if backend == "simt" else (L.SPLIT if split else L.BF16)
    for (cin, cout, k, dil, stride, tr, B, Ln) in [(24, 24, 11, 5, 1, False, 2, 301), (48, 48, 7, 3, 1, False, 1, 130), (96, 48, 4, 1, 2, True, 2, 77),
                                                  (200, 200, 3, 1, 1, False, 1, 257), (16, 8, 8, 1, 4, True, 3, 33),
                                                  # CTA-pair kernel (N tile >= 128): a second 256-row pair tile with one valid row (the
                                                  # peer CTA wholly past the end), 128-column SPLIT tiles, the single-tile C = 192 form,
                                                  # a transposed conv with several N tiles
                                                  (256, 256, 3, 1, 1, False, 1, 257), (384, 384, 3, 3, 1, False, 2, 131), (192, 192, 7, 5, 1, False, 1, 300),
                                                  (384, 192, 4, 1, 2, True, 2, 129)]:
        shape = (cin, cout, k) if tr else (cout, cin, k)
        v = torch.from_numpy(rng.standard_normal(shape).astype(np.float32) * 0.1).to(DEV)
        g = v.flatten(1).norm(dim=1).reshape(-1, 1, 1) * 1.1
        bias = torch.from_numpy(rng.standard_normal(cout).astype(np.float32)).to(DEV)
        pad = (k - stride) // 2 if tr else (k * dil - dil) // 2
        pc = _ops.pack_conv(v, g, bias, transposed=tr, dilation=dil, stride=stride, padding=pad, backend=be, split=split)
        x = torch.from_numpy(rng.standard_normal((B, Ln, pc.x_pitch)).astype(np.float32)).to(DEV)
        if pc.x_pitch > cin:
            x[:, :, cin:] = 0
        xb = _ops.to_buf(x, x_dt)
        n = B * Ln * pc.n_total
        for out_dt in ((L.F32, L.SPLIT) if backend != "umma_bf16" else (L.BF16,)):
            odt = torch.float32 if out_dt == L.F32 else torch.bfloat16
            planes = [torch.full((n + 2 * G,), 321.0, dtype=odt, device=DEV) for _ in range(2 if out_dt == L.SPLIT else 1)]
            esz = planes[0].element_size()
            d = L.ConvDesc()
            d.x = xb.tensor()
            d.out = L.Tensor(planes[0].data_ptr() + G * esz, (planes[1].data_ptr() + G * esz) if len(planes) > 1 else None, out_dt, 0)
            d.res = d.acc_in = L.Tensor(None, None, L.F32, 0)
            d.div, d.B, d.L = 1.0, B, Ln
            d.w = C.pointer(pc.desc)
            L.check(L.lib().bvg_conv_fwd(C.byref(d), torch.cuda.current_stream().cuda_stream), "conv_fwd")
            torch.cuda.synchronize()
            for pl in planes:
                assert bool((pl[:G] == 321.0).all()) and bool((pl[-G:] == 321.0).all()), (backend, cin, cout, k, out_dt)


PAIR_CASES = [
    # (B, Cin, Cout, L, k, d, transposed, stride)
    (2, 256, 256, 300, 3, 1, False, 1),    # two pair tiles per item, the second partial
    (1, 384, 384, 1000, 7, 3, False, 1),   # SPLIT: 3 x 128-column tiles, two accumulator stages
    (3, 192, 192, 129, 11, 5, False, 1),   # single tile: one accumulator for the three products
    (1, 768, 768, 260, 3, 1, False, 1),    # 12 K slices
    (2, 384, 192, 77, 4, 1, True, 2),      # transposed conv, per-tile tap tables
    (1, 100, 256, 515, 7, 1, False, 1),    # conv_pre-like: Cin padded to 104, last K slice with 3 K steps
]


@pytest.mark.parametrize("case", PAIR_CASES)
@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("epi", ["plain", "res", "res_acc_div", "acc", "relu"])
def test_conv_pair_vs_single_cta(ops, case, split, epi):
    """conv_pair_kernel (tcgen05.mma.cta_group::2, the default for N tiles >= 128) against the single-CTA kernel on the
    same layer (bvg_tuning.umma_pair = 0) and against the fp64 oracle: ragged lengths, several batch items, every
    epilogue the pair kernel takes."""
    _ops, L = ops
    B, cin, cout, Ln, k, d, tr, u = case
    rng = np.random.default_rng(cin + cout + Ln + k)
    x = rng.standard_normal((B, cin, Ln)).astype(np.float32)
    shape = (cin, cout, k) if tr else (cout, cin, k)
    v = (rng.standard_normal(shape) / np.sqrt(cin * k)).astype(np.float32)
    g = (np.sqrt((v.astype(np.float64) ** 2).sum(axis=(1, 2), keepdims=True)) * rng.uniform(0.7, 1.4, (shape[0], 1, 1))).astype(np.float32)
    b = (rng.standard_normal(cout) * 0.1).astype(np.float32)
    pad = (k - u) // 2 if tr else O.get_padding(k, d)
    wts = [torch.from_numpy(t).to(DEV) for t in (v, g, b)]
    w64 = O.weight_norm_fold(v.astype(np.float64), g.astype(np.float64))
    ref = O.conv_transpose1d(x.astype(np.float64), w64, b.astype(np.float64), u, pad) if tr else O.conv1d(x.astype(np.float64), w64, b.astype(np.float64), d, pad)
    Lo = ref.shape[-1]
    extra = rng.standard_normal((2, B, cout, Lo)).astype(np.float32)
    kw, want = {}, ref
    if epi in ("res", "res_acc_div"):
        kw["res"] = cl(extra[0]).reshape(B, Ln, -1)
        want = want + extra[0]
    if epi in ("res_acc_div", "acc"):
        kw["acc"] = cl(extra[1]).reshape(B, Ln, -1)
        want = want + extra[1]
    if epi == "res_acc_div":
        kw["div"] = 3.0
        want = want / 3.0
    if epi == "relu":
        want = np.maximum(want, 0.0)

    def run():
        pc = _ops.pack_conv(*wts, transposed=tr, dilation=d, stride=u, padding=pad, backend=L.UMMA, split=split)
        xl = cl(x)
        if pc.x_pitch > cin:
            xl = torch.nn.functional.pad(xl, (0, pc.x_pitch - cin))
        y = _ops.conv(xl, pc, relu=(epi == "relu"), **kw)
        return cf(y.reshape(B, Lo, cout)), pc

    y_pair, pc = run()
    if pc.desc.n_tile < 128:  # (a transposed conv whose per-phase width packs as 96-column tiles: single-CTA kernel)
        assert tr and split
        pytest.skip("this layer packs below the pair kernel's tile width")
    assert pc.desc.split != 2
    try:
        L.set_tuning("umma_pair", 0)
        y_single, _ = run()
    finally:
        L.reset_tuning()
    scale = max(1.0, float(np.abs(want).max()))
    tol = 3e-5 if split else 2e-2
    assert np.abs(y_pair - want).max() < tol * scale, np.abs(y_pair - want).max()
    assert np.abs(y_pair - y_single).max() < tol * scale
