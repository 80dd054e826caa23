"""Shared constants/helpers of the test-suite (no GPU needed to import)."""
import numpy as np

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
TINY_VARIANTS = {
    "b1_snakebeta_log": dict(),
    "b2_snake_lin": dict(resblock="2", activation="snake", snake_logscale=False, resblock_dilation_sizes=[[1, 3], [1, 3]]),
    "b1_snake_log": dict(activation="snake"),
    "b2_snakebeta_lin": dict(resblock="2", snake_logscale=False, resblock_dilation_sizes=[[1, 3], [1, 3]]),
}
REPO = dict(
    TINY,
    resblock_kernel_sizes=[3, 7, 11],
    upsample_rates=[4, 4, 2, 2, 2, 2],
    input_dim=100,
    upsample_initial_channel=1536,
    upsample_kernel_sizes=[8, 8, 4, 4, 4, 4],
    resblock_dilation_sizes=[[1, 3, 5]] * 3,
)
V2 = dict(REPO, input_dim=128, upsample_rates=[8, 4, 2, 2, 2, 2], upsample_kernel_sizes=[16, 8, 4, 4, 4, 4])


def tiny_cfg_sd(tag):
    from svc_inference_pipeline_b200.utils import synth

    cfg = dict(TINY, **TINY_VARIANTS[tag])
    sd = synth.synthetic_state_dict(cfg, seed=7)
    if not cfg["snake_logscale"]:
        for k in sd:
            if k.endswith(".alpha") or k.endswith(".beta"):
                sd[k] = (1.0 + 0.6 * sd[k]).astype(np.float32)
    return cfg, sd


def snr_db(ref, y):
    ref = np.asarray(ref, np.float64)
    y = np.asarray(y, np.float64)
    return 10 * np.log10((ref**2).sum() / max(((y - ref) ** 2).sum(), 1e-300))


def bf16_round(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32)
