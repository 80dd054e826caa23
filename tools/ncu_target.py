"""Short program for ncu: builds the repo generator and runs exactly two forwards at the bench shape
(first = warm-up).  One forward = 226 launches: pack, conv_pre, 6 x (up-conv + 54 layer launches) ..."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from svc_inference_pipeline_b200.modules.bigvgan import Generator
from svc_inference_pipeline_b200.utils import synth
from svc_inference_pipeline_b200.utils.util import load_config

precision = sys.argv[1] if len(sys.argv) > 1 else "fp32"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
T = int(sys.argv[3]) if len(sys.argv) > 3 else 938
cfg = load_config(os.path.join(ROOT, "svc_inference_pipeline_b200", "config", "config.json"))
m = Generator(cfg.vocoder, precision=precision)
m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.synthetic_state_dict({k: cfg.vocoder[k] for k in cfg.vocoder.keys()}, 0).items()})
m = m.cuda().eval()
m.overlap_streams = False  # one program of the whole batch on one stream: 226 launches per forward, in program order
mel = torch.from_numpy(synth.synthetic_mel(B, 100, T, 1235)).cuda()
for _ in range(2):
    y = m(mel)
torch.cuda.synchronize()
print("ok", float(y.abs().max()))
