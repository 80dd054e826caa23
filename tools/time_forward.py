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
ap.add_argument("--v2", action="store_true", help="BASELINE configs[4]: 512x generator (input_dim 128, rates 8,4,2,2,2,2), hop 512 at 44.1 kHz; "
                                                  "one rank's share of B64 x 30 s on 8 GPUs is --batch 8 --frames 2584")
ap.add_argument("--tune", action="append", default=[], help="name=value tuning knob (bvg_tuning, attached to every descriptor built afterwards)")
ap.add_argument("--parts", default="0,2", help="overlap settings to time (0 = one stream, n = n batch parts on n streams)")
ap.add_argument("--pdl", type=int, default=0, help="programmatic dependent launch between the kernels (Generator.set_pdl)")
ap.add_argument("--graph", type=int, default=0, help="replay the forward as a CUDA graph")
a = ap.parse_args()
from svc_inference_pipeline_b200 import _lib as _L
for kv in a.tune:
    k, v = kv.split("=")
    _L.set_tuning(k, int(v))
cfg = load_config(os.path.join(ROOT, "svc_inference_pipeline_b200", "config", "config.json"))
vc = {k: cfg.vocoder[k] for k in cfg.vocoder.keys()}
hop, fs = 256, 24000
if a.v2:
    vc.update(input_dim=128, upsample_rates=[8, 4, 2, 2, 2, 2], upsample_kernel_sizes=[16, 8, 4, 4, 4, 4])
    hop, fs = 512, 44100
from svc_inference_pipeline_b200.utils.util import JsonHParams
m = Generator(JsonHParams(**vc))
m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.synthetic_state_dict(vc, 0).items()})
m = m.cuda().eval()
m.set_pdl(bool(a.pdl))
m.use_cuda_graph = bool(a.graph)
mel = torch.from_numpy(synth.synthetic_mel(a.batch, vc["input_dim"], a.frames, 1235)).cuda()
for prec in a.precisions.split(","):
    m.set_precision(prec)
    outs = {}
    for ov in [int(t) for t in a.parts.split(",")] * 2:
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
        print(f"{prec} pdl={a.pdl} graph={a.graph} overlap={ov}: {ms:.3f} ms/step, {a.batch * a.frames * hop / fs / (ms / 1e3):.0f} audio-s/s", flush=True)
    if len(outs) > 1:
        print(f"{prec} max |overlap - plain| = {max(float((outs[k] - outs[0]).abs().max()) for k in outs if k):.3e}")
