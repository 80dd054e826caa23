"""Per-launch device time of one DiffSVC denoiser step (CUDA events between the launches of the program) and the
PyTorch-eager-on-this-GPU time of the same step (oracle/diffsvc_oracle.py run on the device: measurement baseline only).
    python tools/profile_diffsvc.py [--batch 1] [--frames 379] [--precision fp32]"""
import argparse, ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from svc_inference_pipeline_b200 import _lib as L
from svc_inference_pipeline_b200.modules.diffsvc import DiffSVC
from svc_inference_pipeline_b200.utils import synth
from svc_inference_pipeline_b200.utils.util import JsonHParams
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--frames", type=int, default=379)
ap.add_argument("--precision", default="fp32")
ap.add_argument("--eager", type=int, default=1)
ap.add_argument("--tune", action="append", default=[], help="name=value tuning knob (bvg_tuning)")
ap.add_argument("--brief", type=int, default=0)
a = ap.parse_args()
for kv in a.tune:
    k, v = kv.split("=")
    L.set_tuning(k, int(v))
dev = "cuda:0"
mcfg = dict(noise_schedule_factors=[0.0001, 0.02, 1000], n_mel=100, residual_channels=384, diffusion_fc_size=128, conditioner_size=384,
            dilation_cycle_length=4, residual_kernel_size=3, residual_layer_num=20)
sd = {k: torch.from_numpy(v) for k, v in synth.synthetic_diffsvc_state_dict(mcfg, seed=3).items()}
dm = DiffSVC(JsonHParams(**mcfg), precision=a.precision)
dm.load_state_dict(sd)
dm = dm.to(dev).eval()
B, Ln = a.batch, a.frames
mel, cond = torch.randn(B, Ln, 100, device=dev), torch.randn(B, Ln, 384, device=dev)
t = torch.full((B, 1), 500, dtype=torch.long, device=dev)
for _ in range(3):
    dm(mel, cond, t)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    dm(mel, cond, t)
e1.record()
torch.cuda.synchronize()
print(f"{a.precision} tune={a.tune} B{B}x{Ln}: forward (graph replay + copies) {e0.elapsed_time(e1) / 50:.3f} ms/step, {dm.launches_per_step(B, Ln)} launches")
prog = dm._program(B, Ln)
n = prog.launches
per = (C.c_float * n)()
ms = (C.c_float * L.N_OP_KINDS)()
cnt = (C.c_int32 * L.N_OP_KINDS)()
acc = [0.0] * n
stream = torch.cuda.current_stream().cuda_stream
for _ in range(5):
    L.check(L.lib().bvg_program_run_timed(prog.handle, stream, ms, cnt, per), "timed")
    for i in range(n):
        acc[i] += per[i] / 5
names = ["diffembed", "rowop mel", "conv pre"] + sum([[f"L{i} addvec", f"L{i} dilated", f"L{i} gate", f"L{i} outproj"] for i in range(20)], []) + ["scale", "skipproj", "out"]
assert len(names) == n
if not a.brief:
    for i in list(range(3)) + list(range(3, 3 + 16)) + list(range(n - 3, n)):
        print(f"  {names[i]:14s} {acc[i] * 1e3:7.1f} us")
kinds = {}
for nm, v in zip(names, acc):
    k = nm.split()[-1] if nm.startswith("L") else nm
    kinds[k] = kinds.get(k, 0.0) + v
print("  by kind (ms):", {k: round(v, 3) for k, v in kinds.items()}, "total", round(sum(acc), 3))
if a.eager:
    from oracle import diffsvc_oracle as DO
    sdg = {k: v.to(dev) for k, v in sd.items()}
    for _ in range(3):
        DO.denoiser_forward(sdg, mcfg, mel, cond, t)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        DO.denoiser_forward(sdg, mcfg, mel, cond, t)
    e1.record()
    torch.cuda.synchronize()
    print(f"PyTorch eager on this GPU (fp32, TF32 off): {e0.elapsed_time(e1) / 20:.3f} ms/step")
