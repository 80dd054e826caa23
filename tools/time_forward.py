"""Device time of Generator.forward at the bench shape with / without the two-stream overlap.
    python tools/time_forward.py [--batch 16] [--frames 938]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from svc_inference_pipeline_b200.modules.bigvgan import Generator
from svc_inference_pipeline_b200.utils import synth
from svc_inference_pipeline_b200.utils.util import load_config
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--frames", type=int, default=938)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--precisions", default="fp32,bf16")
a = ap.parse_args()
cfg = load_config(os.path.join(ROOT, "svc_inference_pipeline_b200", "config", "config.json"))
m = Generator(cfg.vocoder)
m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.synthetic_state_dict({k: cfg.vocoder[k] for k in cfg.vocoder.keys()}, 0).items()})
m = m.cuda().eval()
mel = torch.from_numpy(synth.synthetic_mel(a.batch, 100, a.frames, 1235)).cuda()
for prec in a.precisions.split(","):
    m.set_precision(prec)
    outs = {}
    for ov in (False, 2, 3, 4, False, 2, 3, 4):
        m.overlap_streams = bool(ov)
        m.overlap_parts = ov or 2
        for _ in range(3):
            y = m(mel)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            y = m(mel)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        outs[ov] = y
        print(f"{prec} overlap={ov}: {ms:.2f} ms/step, {a.batch * a.frames * 256 / 24000 / (ms / 1e3):.0f} audio-s/s", flush=True)
    print(f"{prec} max |overlap - plain| = {max(float((outs[k] - outs[False]).abs().max()) for k in (2, 3, 4)):.3e}")
