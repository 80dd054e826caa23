"""fp32-path error budget on the GPU: which part of the 1e-4 gate each approximation uses.
For each checkpoint recipe (tests/golden/recipes.npz, repo_generator.npz): max-abs error vs the reference's fp64 waveform of
  fp32 (split operands on tcgen05) x {MUFU on the raw argument, exact range reduction}, fp32_simt (exact-fp32 FFMA convs) x the same.
    python tools/precision_diag.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from util_cases import REPO
from svc_inference_pipeline_b200.modules.bigvgan import Generator
from svc_inference_pipeline_b200.utils import synth
from svc_inference_pipeline_b200.utils.util import JsonHParams
g = np.load(os.path.join(ROOT, "tests/golden/recipes.npz"))
g0 = np.load(os.path.join(ROOT, "tests/golden/repo_generator.npz"))
cases = [("repo", g0["logmel_mel"], g0["logmel_y_f64"], g0["logmel_y"]), ("survey", g["mel"], g["survey_y_f64"], g["survey_y"]), ("large_alpha", g["mel"], g["large_alpha_y_f64"], g["large_alpha_y"])]
for recipe, mel, ref64, ref32 in cases:
    sd = synth.synthetic_state_dict(REPO, 0, recipe=recipe)
    x = torch.from_numpy(mel).cuda()
    row = [f"{recipe:12s} ref fp32-vs-fp64 {np.abs(ref32 - ref64).max():.2e} |y|max {np.abs(ref64).max():.2f}:"]
    for prec in ("fp32", "fp32_simt", "bf16"):
        for precise in (False, True):
            m = Generator(JsonHParams(**REPO), precision=prec, precise_sin=precise)
            m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
            m = m.cuda().eval()
            y = m(x).cpu().numpy()
            err = np.abs(y - ref64).max()
            snr = 10 * np.log10((ref64**2).sum() / ((y - ref64) ** 2).sum())
            row.append(f"{prec}/{'exact' if precise else 'mufu'} {err:.2e} ({snr:.0f} dB)")
            del m
    print("  ".join(row), flush=True)
