"""Short program for ncu: the DiffSVC denoiser (reference mapper hyper-parameters) and exactly two steps of the sampler
(first = warm-up), launch list instead of the CUDA graph: 86 launches of the step program + the sampler update.
    python tools/ncu_diffsvc_target.py [fp32|bf16] [B] [L]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from svc_inference_pipeline_b200.modules.diffsvc import DiffSVC
from svc_inference_pipeline_b200.modules.diffsvcrepo_inference import schedule_tables
from svc_inference_pipeline_b200.utils import synth
from svc_inference_pipeline_b200.utils.util import JsonHParams

precision = sys.argv[1] if len(sys.argv) > 1 else "fp32"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
Ln = int(sys.argv[3]) if len(sys.argv) > 3 else 379
mcfg = dict(noise_schedule_factors=[0.0001, 0.02, 1000], n_mel=100, residual_channels=384, diffusion_fc_size=128, conditioner_size=384,
            dilation_cycle_length=4, residual_kernel_size=3, residual_layer_num=20)
dm = DiffSVC(JsonHParams(**mcfg), precision=precision)
dm.load_state_dict({k: torch.from_numpy(v) for k, v in synth.synthetic_diffsvc_state_dict(mcfg, seed=3).items()})
dm = dm.cuda().eval()
dm.use_cuda_graph = False
state = dm.sampler(torch.randn(B, Ln, 384, device="cuda"), schedule_tables(np.linspace(1e-4, 0.02, 1000).tolist()))
state.x.normal_()
for step in (999, 998):
    state.noise.normal_()
    state.ddpm_step(step)
torch.cuda.synchronize()
print("ok", float(state.x.abs().max()))
