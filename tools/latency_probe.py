"""Where the wall-clock time of one utterance through synthesis_audios goes (configs[0]: 379 frames): H2D, forward as
launch list / CUDA graph / with programmatic dependent launch, D2H, host tail.
    python tools/latency_probe.py"""
import os, sys, time
sys.path.insert(0, "/root/repo")
import torch, numpy as np
from svc_inference_pipeline_b200.modules.bigvgan import Generator
from svc_inference_pipeline_b200.modules.bigvgan_inference import synthesis_audios, vocoder_inference
from svc_inference_pipeline_b200.utils import synth
from svc_inference_pipeline_b200.utils.util import load_config
cfg = load_config("/root/repo/svc_inference_pipeline_b200/config/config.json")
m = Generator(cfg.vocoder)
m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.synthetic_state_dict({k: cfg.vocoder[k] for k in cfg.vocoder.keys()}, 0).items()})
m = m.cuda().eval()
m.use_cuda_graph = True
mel = torch.from_numpy(synth.synthetic_mel(1, 100, 379, 1))[0]
for _ in range(3): synthesis_audios(m, mel, cfg)
def T(f, n=20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): r = f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("synthesis_audios      %.3f ms" % T(lambda: synthesis_audios(m, mel, cfg)))
dev = torch.device("cuda:0")
print("vocoder_inference     %.3f ms" % T(lambda: vocoder_inference(cfg, m, mel.unsqueeze(0), dev)))
md = mel.unsqueeze(0).cuda()
print("H2D mel               %.3f ms" % T(lambda: mel.unsqueeze(0).to(dev)))
print("forward_borrowed      %.3f ms" % T(lambda: m.forward_borrowed(md)))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): m.forward_borrowed(md)
e1.record(); torch.cuda.synchronize()
print("forward device time   %.3f ms" % (e0.elapsed_time(e1) / 20))
y = m.forward_borrowed(md).squeeze(1)
def d2h():
    h = torch.empty(y.shape, dtype=y.dtype, pin_memory=True); h.copy_(y, non_blocking=True); torch.cuda.current_stream().synchronize(); return h
print("D2H pinned            %.3f ms" % T(d2h))
print("D2H .cpu()            %.3f ms" % T(lambda: y.cpu()))
a = y.cpu()[0]
def tail():
    f = torch.linspace(1, 0, steps=20 * 256); b = a.clone()[: 379 * 256]; b[-5120:] *= f; return b.numpy()
print("fade tail (host)      %.3f ms" % T(tail))
m.use_cuda_graph = False
print("forward eager launches %.3f ms" % T(lambda: m.forward_borrowed(md)))
m.set_pdl(True)
print("forward eager + pdl   %.3f ms" % T(lambda: m.forward_borrowed(md)))
m.use_cuda_graph = True
print("forward graph + pdl   %.3f ms" % T(lambda: m.forward_borrowed(md)))
