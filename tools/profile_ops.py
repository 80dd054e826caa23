"""Per-launch device-time table of one Generator.forward (CUDA events between launches).
    python tools/profile_ops.py [--precision fp32|bf16] [--batch 16] [--frames 938] [--top 30]"""
import argparse, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from svc_inference_pipeline_b200.modules.bigvgan import Generator
from svc_inference_pipeline_b200.utils import synth
from svc_inference_pipeline_b200.utils.util import load_config

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="fp32")
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--frames", type=int, default=938)
ap.add_argument("--out", default=None)
ap.add_argument("--tune", action="append", default=[], help="name=value tuning knob (bvg_tuning, attached to every descriptor built afterwards)")
ap.add_argument("--no-fold", action="store_true", help="narrow convolutions without time folding")
ap.add_argument("--fuse", action="store_true", help="Activation1d fused into the narrow convolutions (Generator.fuse_amp, off by default)")
a = ap.parse_args()
from svc_inference_pipeline_b200 import _lib as _L
for kv in a.tune:
    k, v = kv.split("=")
    _L.set_tuning(k, int(v))
cfg = load_config(os.path.join(ROOT, "svc_inference_pipeline_b200", "config", "config.json"))
m = Generator(cfg.vocoder, precision=a.precision)
m.time_fold = not a.no_fold
m.fuse_amp = a.fuse
m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.synthetic_state_dict({k: cfg.vocoder[k] for k in cfg.vocoder.keys()}, 0).items()})
m = m.cuda().eval()
mel = torch.from_numpy(synth.synthetic_mel(a.batch, 100, a.frames, 1235)).cuda()
for _ in range(2):
    m(mel)
rows = m.profile_ops(a.batch, a.frames, reps=2)
tot = sum(r[2] for r in rows)
lines = [f"# precision={a.precision} batch={a.batch} frames={a.frames} tune={a.tune} total={tot:.2f} ms"]
agg = {}
for lab, kind, ms, work in rows:
    rate = (work / (ms * 1e-3) / 1e12 if kind == "conv" else work / (ms * 1e-3) / 1e9) if ms > 0 else 0
    unit = "TFLOP/s" if kind == "conv" else "GB/s"
    lines.append(f"{kind:5s} {ms:8.3f} ms {100 * ms / tot:5.1f}%  {rate:9.1f} {unit:8s} {lab}")
    key = (kind, lab.split(" L")[-1] if kind != "pack" else "")
    g = agg.setdefault((kind, lab.split()[-1]), [0.0, 0.0, 0])
    g[0] += ms; g[1] += work; g[2] += 1
lines.append("# by (kind, L):")
for (kind, ln), (ms, work, n) in sorted(agg.items()):
    rate = (work / (ms * 1e-3) / 1e12 if kind == "conv" else work / (ms * 1e-3) / 1e9) if ms > 0 else 0
    lines.append(f"#  {kind:5s} {ln:10s} n={n:3d} {ms:8.2f} ms {100 * ms / tot:5.1f}%  {rate:9.1f} {'TFLOP/s' if kind == 'conv' else 'GB/s'}")
text = "\n".join(lines)
print(text)
if a.out:
    open(a.out, "w").write(text + "\n")
