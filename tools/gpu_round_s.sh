#!/bin/bash
# sampler (svc_model_inference) parity + timing
set -u
OUT=gpurun_out
python -m pytest tests/test_sampler.py tests/test_diffsvc.py -m gpu -q -s > $OUT/r02_pytest_sampler.log 2>&1; echo "pytest rc=$?"; grep -E "^sampler|passed|failed|Error|error" $OUT/r02_pytest_sampler.log | head -40
python - <<'PY'
import time, numpy as np, torch
from svc_inference_pipeline_b200.modules.diffsvc import DiffSVC
from svc_inference_pipeline_b200.modules.diffsvcrepo_inference import svc_model_inference
from svc_inference_pipeline_b200.utils import synth
from svc_inference_pipeline_b200.utils.util import JsonHParams
dev = "cuda:0"
mcfg = dict(noise_schedule_factors=[0.0001, 0.02, 1000], n_mel=100, residual_channels=384, diffusion_fc_size=128, conditioner_size=384,
            dilation_cycle_length=4, residual_kernel_size=3, residual_layer_num=20)
dm = DiffSVC(JsonHParams(**mcfg), precision="fp32")
dm.load_state_dict({k: torch.from_numpy(v) for k, v in synth.synthetic_diffsvc_state_dict(mcfg, seed=3).items()})
dm = dm.to(dev).eval()
sched = np.linspace(1e-4, 0.02, 1000).tolist()
batch = {"y": torch.zeros(1, 379, 100, device=dev), "cond": torch.randn(1, 379, 384, device=dev)}
model = [lambda b: b["cond"], dm]
for prec in ("fp32", "bf16"):
    dm.set_precision(prec)
    for fast in (False, True):
        svc_model_inference(model, batch, JsonHParams(mapper=JsonHParams(noise_schedule=sched[:20])), fast_inference=fast)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        y = svc_model_inference(model, batch, JsonHParams(mapper=JsonHParams(noise_schedule=sched)), fast_inference=fast).cpu()
        print(prec, "plms100" if fast else "ddpm1000", f"{time.perf_counter() - t0:.3f} s", tuple(y.shape), bool(torch.isfinite(y).all()), float(y.abs().max()))
PY
