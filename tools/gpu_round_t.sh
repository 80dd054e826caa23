#!/bin/bash
# programmatic dependent launch: parity (whole GPU suite), sampler and vocoder timings on / off
set -u
OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/r02_pytest_gpu_t.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/r02_pytest_gpu_t.log
timeout 300 python - <<'PY'
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
outs = {}
for pdl in (0, 1, 0, 1):
    dm.set_pdl(bool(pdl))
    for prec in ("fp32", "bf16"):
        dm.set_precision(prec)
        svc_model_inference(model, batch, JsonHParams(mapper=JsonHParams(noise_schedule=sched[:20])))
        torch.cuda.synchronize()
        torch.manual_seed(5)
        t0 = time.perf_counter()
        y = svc_model_inference(model, batch, JsonHParams(mapper=JsonHParams(noise_schedule=sched))).cpu()
        dt = time.perf_counter() - t0
        same = torch.equal(outs.setdefault(prec, y), y)
        print(f"sampler pdl={pdl} {prec} ddpm1000 {dt:.3f} s identical={same}", flush=True)
PY
for pdl in 0 1; do for graph in 0 1; do
  timeout 300 python tools/time_forward.py --batch 1 --frames 379 --reps 20 --parts 0 --pdl $pdl --graph $graph 2>&1 | grep "ms/step"
done; done
for pdl in 0 1 0 1; do
  timeout 300 python tools/time_forward.py --pdl $pdl --precisions fp32 --parts 0,2 2>&1 | grep "ms/step"
  timeout 300 python tools/time_forward.py --pdl $pdl --precisions bf16 --parts 0 2>&1 | grep "ms/step"
done
