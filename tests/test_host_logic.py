"""Host-side logic that needs no GPU: config reader, checkpoint filter, state_dict grammar,
C-ABI symbol/struct checks, conv geometry tables."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from util_cases import REPO, TINY, V2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_json5_subset_reader(tmp_path):
    from svc_inference_pipeline_b200.utils.util import JsonHParams, load_config, loads_json5_subset

    text = '{\n // comment\n "a": 1, /* block */ "s": "http://x//y", "l": [1, 2, 3,], "d": {"k": true,},\n}'
    d = loads_json5_subset(text)
    assert d == {"a": 1, "s": "http://x//y", "l": [1, 2, 3], "d": {"k": True}}
    cfg = load_config(os.path.join(ROOT, "svc_inference_pipeline_b200", "config", "config.json"))
    assert isinstance(cfg, JsonHParams) and cfg.hop_length == 256 and cfg.vocoder.upsample_rates == [4, 4, 2, 2, 2, 2]
    assert "vocoder" in cfg and cfg["vocoder"].activation == "snakebeta" and len(cfg.vocoder) == 9
    assert int(np.prod(cfg.vocoder.upsample_rates)) == cfg.hop_length


def test_basic_config_inheritance(tmp_path, monkeypatch):
    from svc_inference_pipeline_b200.utils.util import load_config

    (tmp_path / "base.json").write_text('{"a": 1, "v": {"x": 1, "y": 2}}')
    (tmp_path / "child.json").write_text('{"basic_config": "base.json", "v": {"y": 3}, "b": 2,}')
    monkeypatch.setenv("WORD_DIR", str(tmp_path))
    cfg = load_config(str(tmp_path / "child.json"))
    assert cfg.a == 1 and cfg.b == 2 and cfg.v.x == 1 and cfg.v.y == 3


def test_generator_structure_matches_reference_grammar():
    from svc_inference_pipeline_b200.modules.bigvgan import Generator
    from svc_inference_pipeline_b200.utils import synth
    from svc_inference_pipeline_b200.utils.util import JsonHParams

    for cfgd, n_params in ((TINY, None), (REPO, 112_446_290), (V2, 122_184_530)):
        model = Generator(JsonHParams(**cfgd))
        spec = synth.state_dict_spec(cfgd)
        sd = model.state_dict()
        assert list(sd.keys()) == list(spec.keys())
        assert all(tuple(sd[k].shape) == spec[k][0] for k in sd)
        if n_params:
            assert sum(p.numel() for p in model.parameters()) == n_params
            assert len(sd) == 784
    with pytest.raises(NotImplementedError):
        Generator(JsonHParams(**dict(TINY, activation="relu")))
    with pytest.raises(RuntimeError, match="no CPU path"):
        Generator(JsonHParams(**TINY))(torch.zeros(1, 10, 8))


def test_default_init_matches_weight_norm_identity():
    from svc_inference_pipeline_b200.modules.bigvgan import Generator
    from svc_inference_pipeline_b200.utils.util import JsonHParams

    m = Generator(JsonHParams(**TINY))
    v, g = m.ups[0][0].weight_v, m.ups[0][0].weight_g
    assert g.shape == (32, 1, 1)  # ConvTranspose1d: dim 0 is Cin
    torch.testing.assert_close(g.flatten(), v.flatten(1).norm(dim=1))
    assert float(m.resblocks[0].activations[0].act.alpha.abs().max()) == 0.0  # log-scale init
    # remove_weight_norm keeps the folded weight
    w_before = v * (g / v.flatten(1).norm(dim=1).reshape(-1, 1, 1))
    m.remove_weight_norm()
    torch.testing.assert_close(m.ups[0][0].weight_v, w_before)


def test_checkpoint_filter_and_loader_cpu(tmp_path):
    from svc_inference_pipeline_b200.utils import synth
    from svc_inference_pipeline_b200.utils.load_models import filter_state_dict, vocoder_model_loader
    from svc_inference_pipeline_b200.utils.util import JsonHParams

    sd = synth.synthetic_state_dict(TINY, seed=3)
    ckpt = {"module." + k: torch.from_numpy(v) for k, v in sd.items()}
    ckpt["module.conv_pre.bias"] = torch.zeros(7)  # wrong shape -> dropped, keeps init
    ckpt["module.not_a_key"] = torch.zeros(1)
    del ckpt["module.conv_post.bias"]
    path = str(tmp_path / "v.pt")
    torch.save({"generator_state_dict": ckpt}, path)
    cfg = JsonHParams(device="cpu", vocoder_model_path=path, hop_length=8, vocoder=dict(TINY))
    with pytest.warns(UserWarning):
        model = vocoder_model_loader(cfg)
    rep = model.load_report
    assert rep["unknown"] == ["not_a_key"]
    assert [w[0] for w in rep["wrong_shape"]] == ["conv_pre.bias"]
    assert set(rep["missing"]) == {"conv_pre.bias", "conv_post.bias"}
    assert not model.training and next(model.parameters()).device.type == "cpu"
    np.testing.assert_array_equal(model.state_dict()["ups.0.0.weight_v"].numpy(), sd["ups.0.0.weight_v"])
    kept, _ = filter_state_dict({"module.conv_pre.weight_g": torch.zeros(32, 1, 1)}, model.state_dict())
    assert list(kept) == ["conv_pre.weight_g"]
    # a folded ".weight" checkpoint (saved after remove_weight_norm) is accepted too
    folded = {k: torch.from_numpy(v) for k, v in sd.items() if not k.startswith("conv_pre.weight")}
    v, g = sd["conv_pre.weight_v"], sd["conv_pre.weight_g"]
    w = v * (g / np.sqrt((v**2).sum(axis=(1, 2), keepdims=True)))
    folded["conv_pre.weight"] = torch.from_numpy(w)
    model.load_state_dict(folded)
    torch.testing.assert_close(model.conv_pre.weight_v.detach(), torch.from_numpy(w))


def header_functions():
    text = open(os.path.join(ROOT, "include", "bvg_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bvg_[a-z0-9_]+)\s*\(", text)))


def test_capi_exports_every_declared_symbol():
    from svc_inference_pipeline_b200 import _lib as L

    lib = L.lib()  # raises if the .so is missing: there is no fallback
    names = header_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/bvg_b200.h but not exported"
    assert sorted(L.EXPORTS) == names
    assert lib.bvg_abi_version() == L.ABI_VERSION
    assert lib.bvg_sizeof_op() == C.sizeof(L.Op)
    assert lib.bvg_sizeof_conv_weights() == C.sizeof(L.ConvWeights)


def test_conv_geometry_tables():
    from svc_inference_pipeline_b200 import _lib as L

    lib = L.lib()

    def geom(*a):
        g, w = L.ConvGeom(*a), L.ConvWeights()
        L.check(lib.bvg_conv_geometry(C.byref(g), C.byref(w)))
        return w

    w = geom(0, 768, 768, 11, 5, 1, 25, L.UMMA, 0, 0)
    assert (w.n_total, w.n_tile, w.n_tiles, w.cin_pad, w.x_pitch, w.tap_stride, w.split) == (768, 256, 3, 768, 768, 11, 0)
    assert list(w.shift[2])[:11] == [5 * (j - 5) for j in range(11)]
    # split operands: wide layers keep two weight planes for the CTA-pair kernel (conv_pair.cu, correction products in their
    # own accumulator columns); without it (bvg_tuning.umma_pair = 0) they take 128-column
    # tiles with both planes stacked (split == 2) for the same reason; narrow layers are always stacked
    w = geom(0, 768, 768, 11, 5, 1, 25, L.UMMA, 1, 0)   # pair kernel: 128 columns, two planes, main + correction x 2 stages
    assert (w.n_total, w.n_tile, w.n_tiles, w.split) == (768, 128, 6, 1)
    w = geom(0, 384, 384, 7, 1, 1, 3, L.UMMA, 1, 0)
    assert (w.n_tile, w.n_tiles, w.split) == (128, 3, 1)
    w = geom(0, 192, 192, 7, 1, 1, 3, L.UMMA, 1, 0)      # single tile: one accumulator for the three products, two stages
    assert (w.n_tile, w.n_tiles, w.split) == (192, 1, 1)
    w = geom(0, 96, 96, 7, 1, 1, 3, L.UMMA, 1, 0)
    assert (w.n_tile, w.n_tiles, w.split) == (96, 1, 2)
    try:
        L.set_tuning("umma_pair", 0)

        def geom0(*a):
            g, w = L.ConvGeom(*a), L.ConvWeights()
            g.tune = L.tuning_ptr()
            L.check(lib.bvg_conv_geometry(C.byref(g), C.byref(w)))
            return w

        w = geom0(0, 768, 768, 11, 5, 1, 25, L.UMMA, 1, 0)
        assert (w.n_total, w.n_tile, w.n_tiles, w.split) == (768, 128, 6, 2)
        assert list(w.shift[5])[:11] == [5 * (j - 5) for j in range(11)]
        w = geom0(0, 384, 384, 7, 1, 1, 3, L.UMMA, 1, 0)
        assert (w.n_tile, w.n_tiles, w.split) == (128, 3, 2)
        w = geom0(0, 192, 192, 7, 1, 1, 3, L.UMMA, 1, 0)
        assert (w.n_tile, w.n_tiles, w.split) == (192, 1, 1)
        w = geom0(1, 1536, 768, 16, 1, 8, 4, L.UMMA, 1, 0)  # v2 ups.0: 6144 columns would be 48 tiles of 128 -> stays at 256
        assert (w.n_total, w.n_tile, w.n_tiles, w.split) == (6144, 256, 24, 1)
    finally:
        L.reset_tuning()
    w = geom(1, 1536, 768, 8, 1, 4, 2, L.UMMA, 0, 0)  # ups.0: phase r uses taps {-1,0} (r<2) or {0,1}
    assert (w.n_total, w.n_tile, w.n_tiles, w.tap_stride) == (3072, 256, 12, 2)
    assert [list(w.shift[t])[:2] for t in (0, 5, 6, 11)] == [[-1, 0], [-1, 0], [0, 1], [0, 1]]
    w = geom(1, 48, 24, 4, 1, 2, 1, L.UMMA, 0, 0)  # narrow: one tile spans both phases -> 3 taps
    assert (w.n_total, w.n_tile, w.n_tiles, w.n_taps[0]) == (48, 48, 1, 3)
    w = geom(0, 100, 1536, 7, 1, 1, 3, L.UMMA, 0, 0)
    assert (w.cin_pad, w.x_pitch, w.n_tiles) == (128, 104, 6)
    w = geom(0, 100, 1536, 7, 1, 1, 3, L.SIMT, 0, 0)
    assert (w.cin_pad, w.x_pitch, w.n_tile, w.n_tiles) == (100, 100, 64, 24)
    w = geom(1, 768, 384, 16, 1, 8, 4, L.SIMT, 0, 0)
    assert w.n_taps[0] == 3 and list(w.shift[0])[:3] == [-1, 0, 1]
    g = L.ConvGeom(0, 8, 8, 40, 1, 1, 0, L.SIMT, 0, 0)
    assert lib.bvg_conv_geometry(C.byref(g), C.byref(L.ConvWeights())) != 0
    assert b"BVG_MAX_TAPS" in lib.bvg_last_error()


def test_conv_geometry_time_fold():
    """bvg_conv_geom.fold (host part): the folded layer's sizes and tap table, and -- in numpy, with the oracle's
    conv1d -- that the block-Toeplitz arrangement pack.cu::folded_weight implements is the same convolution."""
    from oracle import bigvgan_oracle as O
    from svc_inference_pipeline_b200 import _lib as L

    lib = L.lib()

    def geom(*a):
        g, w = L.ConvGeom(*a), L.ConvWeights()
        rc = lib.bvg_conv_geometry(C.byref(g), C.byref(w))
        return rc, w

    rc, w = geom(0, 24, 24, 11, 1, 1, 5, L.UMMA, 1, 0, 4)
    assert rc == 0 and (w.cin, w.n_total, w.x_pitch, w.cin_pad, w.n_tiles, w.n_taps[0]) == (96, 96, 96, 128, 1, 5)
    assert list(w.shift[0])[:5] == [-2, -1, 0, 1, 2]
    rc, w = geom(0, 24, 24, 11, 5, 1, 25, L.UMMA, 0, 0, 4)
    assert rc == 0 and w.n_taps[0] == 15 and list(w.shift[0])[:15] == list(range(-7, 8))
    rc, w = geom(0, 24, 24, 3, 1, 1, 1, L.UMMA, 0, 0, 4)
    assert rc == 0 and list(w.shift[0])[:3] == [-1, 0, 1]
    assert geom(1, 48, 24, 4, 1, 2, 1, L.UMMA, 0, 0, 2)[0] != 0   # transposed convs do not fold
    assert geom(0, 24, 24, 11, 1, 1, 5, L.SIMT, 0, 0, 4)[0] != 0  # nor does the SIMT backend
    rc, w1 = geom(0, 24, 24, 11, 1, 1, 5, L.UMMA, 0, 0, 1)
    rc, w0 = geom(0, 24, 24, 11, 1, 1, 5, L.UMMA, 0, 0, 0)
    assert (w1.cin, w1.n_total, w1.n_taps[0]) == (w0.cin, w0.n_total, w0.n_taps[0]) == (24, 24, 11)

    rng = np.random.default_rng(5)
    for (ch, k, d, P, Ln) in [(6, 11, 1, 4, 40), (4, 7, 3, 4, 24), (4, 3, 5, 2, 18), (2, 11, 5, 4, 64)]:
        pad = O.get_padding(k, d)
        wt = rng.standard_normal((ch, ch, k))
        bias = rng.standard_normal(ch)
        x = rng.standard_normal((2, ch, Ln))
        ref = O.conv1d(x, wt, bias, d, pad)
        s_lo, s_hi = -((pad + P - 1) // P), ((k - 1) * d - pad + P - 1) // P
        wf = np.zeros((P * ch, P * ch, s_hi - s_lo + 1))
        for si, sh in enumerate(range(s_lo, s_hi + 1)):
            for po in range(P):
                for pi in range(P):
                    num = P * sh + pi - po + pad
                    if num >= 0 and num % d == 0 and num // d < k:
                        wf[po * ch:(po + 1) * ch, pi * ch:(pi + 1) * ch, si] = wt[:, :, num // d]
        # [B, C, L] -> folded [B, P*C, L/P] (row q holds times P q .. P q + P - 1, phase-major channels)
        xf = x.reshape(2, ch, Ln // P, P).transpose(0, 3, 1, 2).reshape(2, P * ch, Ln // P)
        yf = O.conv1d(xf, wf, np.tile(bias, P), 1, -s_lo)
        assert s_hi == -s_lo  # odd kernels with "same" padding fold symmetrically
        y = yf.reshape(2, P, ch, Ln // P).transpose(0, 2, 3, 1).reshape(2, ch, Ln)
        np.testing.assert_allclose(y, ref, atol=1e-12)


def test_no_product_import_of_oracle():
    """The oracle is test infrastructure: nothing under the package may import it."""
    pkg = os.path.join(ROOT, "svc_inference_pipeline_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_amp_mma_index_math():
    """Lane-level emulation of the tensor-core Activation1d kernel (amp_mma.cu): Toeplitz fragments,
    fragment chaining, tile offsets and both replicate clamps reproduce the oracle exactly."""
    import amp_mma_emulation as E

    assert E.main(lengths=(1, 2, 3, 4, 7, 8, 9, 12, 13, 16, 20, 31, 32, 33, 67), verbose=False) < 1e-12


def test_load_mel_min_max_honours_reference_config_keys(tmp_path):
    """Reference contract (utils/acoustic_feature_extraction.py:66-72): the statistics come from the pickles named by
    cfg.min_mel_file / cfg.max_mel_file; the shipped copy is only a fallback, and a config without them warns."""
    import pickle
    import warnings

    import torch

    from svc_inference_pipeline_b200.utils.acoustic_feature_extraction import denormalize_mel_channel, load_mel_min_max
    from svc_inference_pipeline_b200.utils.util import JsonHParams

    rng = np.random.default_rng(1)
    lo = rng.uniform(-12, -10, 100).astype(np.float32)
    hi = rng.uniform(-5, 1, 100).astype(np.float32)
    for name, arr in (("mel_min.pkl", lo), ("mel_max.pkl", hi)):
        with open(tmp_path / name, "wb") as f:
            pickle.dump(arr, f)
    cfg = JsonHParams(min_mel_file=str(tmp_path / "mel_min.pkl"), max_mel_file=str(tmp_path / "mel_max.pkl"))
    got_lo, got_hi = load_mel_min_max(cfg)
    np.testing.assert_array_equal(got_lo, lo)
    np.testing.assert_array_equal(got_hi, hi)
    mel = rng.uniform(-1, 1, (100, 9)).astype(np.float32)
    ref = (mel + 1) / 2 * (np.expand_dims(hi, -1) - np.expand_dims(lo, -1) + 1e-12) + np.expand_dims(lo, -1)
    np.testing.assert_array_equal(denormalize_mel_channel(torch.from_numpy(mel), cfg).numpy(), ref)
    np.savez(tmp_path / "range.npz", mel_min=lo + 1, mel_max=hi + 1)
    np.testing.assert_array_equal(load_mel_min_max(JsonHParams(mel_range_path=str(tmp_path / "range.npz")))[0], lo + 1)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        d_lo, _ = load_mel_min_max(JsonHParams(hop_length=256))
        assert any("shipped" in str(x.message) for x in w)
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        np.testing.assert_array_equal(load_mel_min_max()[0], d_lo)  # no config: the shipped copy, silently


def test_tuning_travels_with_descriptors():
    """The library keeps no global knobs (include/bvg_b200.h, bvg_tuning): the binding's Tuning object is attached to
    descriptors only when a knob differs from bvg_tuning_defaults(), and geometry reads it from bvg_conv_geom.tune."""
    from svc_inference_pipeline_b200 import _lib as L

    lib = L.lib()
    assert not hasattr(lib, "bvg_set_tuning")
    L.reset_tuning()
    assert L.tuning_ptr() is None
    t = L.Tuning()
    lib.bvg_tuning_defaults(C.byref(t))
    assert (t.amp_mma, t.amp_packed, t.amp_ct, t.umma_ntile_cap, t.umma_stack, t.umma_pair, t.amp_vec, t.umma_mb) == (1, 1, 1, 256, 128, 1, 0, 0)
    g, w = L.ConvGeom(0, 768, 768, 3, 1, 1, 1, L.UMMA, 0, 0), L.ConvWeights()
    L.check(lib.bvg_conv_geometry(C.byref(g), C.byref(w)))
    assert w.n_tile == 256
    try:
        L.set_tuning("umma_ntile_cap", 128)
        assert L.tuning_ptr() is not None
        g.tune = L.tuning_ptr()
        L.check(lib.bvg_conv_geometry(C.byref(g), C.byref(w)))
        assert w.n_tile == 128
        with pytest.raises(L.BvgError):
            L.set_tuning("no_such_knob", 1)
    finally:
        L.reset_tuning()
    assert L.tuning_ptr() is None


def test_product_mel_filterbank_is_the_pinned_one(golden):
    """utils/mel.py::mel_filterbank (product, feeds bvg_logmel_fwd) == the basis make_golden.py committed after checking
    it against transformers.audio_utils.mel_filter_bank(norm="slaney", mel_scale="slaney")."""
    import torch  # noqa: F401  (utils.mel imports torch)

    from svc_inference_pipeline_b200.utils.mel import mel_filterbank

    np.testing.assert_array_equal(mel_filterbank(24000, 1024, 100, 0, 12000), golden("logmel.npz")["basis"])
    fb = mel_filterbank(44100, 2048, 128, 0, 22050)
    assert fb.shape == (128, 1025) and (fb >= 0).all() and (fb.sum(axis=1) > 0).all()


def test_new_tuning_is_a_private_copy():
    """L.new_tuning: the binding's knobs with overrides on top, independent of later set_tuning calls (DiffSVC packs its
    dense layers with a row-count dependent tile cap without touching the process-wide object)."""
    from svc_inference_pipeline_b200 import _lib as L
    from svc_inference_pipeline_b200.modules.diffsvc import DiffSVC

    try:
        L.reset_tuning()
        own = L.new_tuning(umma_ntile_cap=32)
        assert own.umma_ntile_cap == 32 and own.umma_pair == 1 and L.tuning_ptr() is None
        L.set_tuning("umma_pair", 0)
        assert own.umma_pair == 1 and L.new_tuning().umma_pair == 0
        with pytest.raises(L.BvgError):
            L.new_tuning(no_such_knob=1)
    finally:
        L.reset_tuning()
    assert [DiffSVC.tile_cap(r) for r in (379, 512, 513, 4096, 4097, 15008)] == [32, 32, 64, 64, 128, 128]


def test_descriptor_validation_without_a_gpu():
    """The C ABI rejects malformed descriptors before any CUDA call (runs on a box without a GPU): sampler update, row
    operations, per-channel divisor of the convolution epilogue, program flags."""
    from svc_inference_pipeline_b200 import _lib as L

    lib = L.lib()
    d = L.SampleDesc()
    assert lib.bvg_sample_fwd(C.byref(d), None) == -1 and b"null pointer" in lib.bvg_last_error()
    buf = (C.c_float * 64)()
    p = C.addressof(buf)
    d.d_x = d.d_x_out = d.d_eps = d.d_step = p
    d.B, d.L, d.n_mel, d.n_steps = 1, 4, 4, 8
    d.mode = L.SAMPLE_DDPM
    assert lib.bvg_sample_fwd(C.byref(d), None) == -1 and b"schedule tables" in lib.bvg_last_error()
    d.mode = L.SAMPLE_PLMS
    assert lib.bvg_sample_fwd(C.byref(d), None) == -1 and b"alphas_cumprod" in lib.bvg_last_error()
    d.d_alphas_cumprod, d.interval, d.combine = p, 10, 5
    assert lib.bvg_sample_fwd(C.byref(d), None) == -1 and b"unknown combination" in lib.bvg_last_error()
    d.combine = 3  # two earlier predictions needed, none given
    assert lib.bvg_sample_fwd(C.byref(d), None) == -1 and b"earlier predictions" in lib.bvg_last_error()
    d.mode = 7
    assert lib.bvg_sample_fwd(C.byref(d), None) == -1 and b"unknown mode" in lib.bvg_last_error()
    r = L.RowopDesc()
    r.kind, r.d_x, r.B, r.L, r.C, r.x_pitch, r.out_pitch = L.ROW_GATE, p, 1, 4, 8, 8, 8   # GATE reads 2C channels
    r.out = L.Tensor(p, None, L.F32, 0)
    assert lib.bvg_rowop_fwd(C.byref(r), None) == -1 and b"row pitch" in lib.bvg_last_error()
    assert lib.bvg_program_set_pdl(None, 1) == -1


def test_tile_cap_is_a_preference_bounded_by_the_tap_tables():
    """bvg_tuning.umma_ntile_cap narrows the N tiles (DiffSVC packs 32-column tiles for one utterance), but a layer too
    wide for BVG_MAX_NTILES tiles of that width gets wider tiles instead of an error."""
    from svc_inference_pipeline_b200 import _lib as L

    def geom(cap, cin, cout, k, transposed=0, stride=1, pad=0):
        t = L.new_tuning(umma_ntile_cap=cap)
        g = L.ConvGeom(transposed, cin, cout, k, 1, stride, pad, L.UMMA, 1, 0, 1)
        g.tune = C.pointer(t)
        w = L.ConvWeights()
        L.check(L.lib().bvg_conv_geometry(C.byref(g), C.byref(w)), "geometry")
        return w.n_tile, w.n_tiles

    assert geom(32, 384, 768, 3, pad=1) == (32, 24)
    assert geom(32, 1024, 2048, 3, pad=1) == (64, 32)              # 64 tiles of 32 columns would not fit the tables
    assert geom(64, 1536, 768, 8, 1, 4, 2) == (128, 24)            # transposed: 4 phases x 768 columns
    assert geom(32, 1536, 768, 16, 1, 8, 4) == (256, 24)           # the v2 generator's first up-conv: 8 x 768 columns
