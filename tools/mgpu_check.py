"""Multi-GPU check under torchrun (one rank per GPU, NCCL): the time-sharded long-form path and the
batch-sharded path against a single-GPU whole forward, plus the device time of the 1-hour configuration
(BASELINE.json configs[3]) when --hour is given.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/mgpu_check.py [--hour]"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from svc_inference_pipeline_b200 import sharding as S
from svc_inference_pipeline_b200.modules.bigvgan import Generator
from svc_inference_pipeline_b200.utils import synth
from svc_inference_pipeline_b200.utils.util import load_config

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
cfg = load_config(os.path.join(ROOT, "svc_inference_pipeline_b200", "config", "config.json"))
prec = "bf16" if "--bf16" in sys.argv else "fp32"
m = Generator(cfg.vocoder, precision=prec)
m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.synthetic_state_dict({k: cfg.vocoder[k] for k in cfg.vocoder.keys()}, 0).items()})
m = m.to(dev).eval()
res = {"world": world, "precision": prec}
# 1. long-form: 3000 frames (32 s) cut into 256-frame chunks, sharded along time
mel = torch.from_numpy(synth.synthetic_mel(1, 100, 3000, 77))[0].to(dev)
full = m(mel[None])[0, 0]
out = S.vocode_long_distributed(m, mel, 256, chunk_frames=256, batch_chunks=4)
res["long_max_abs_vs_full"] = float((out - full).abs().max())
# 2. batch sharding: 5 utterances over `world` ranks (unequal shares)
mels = torch.from_numpy(synth.synthetic_mel(5, 100, 120, 78)).to(dev)
yb = S.vocode_batch_distributed(m, mels, 256)
res["batch_max_abs_vs_full"] = float((yb - m(mels)[:, 0]).abs().max())
if "--hour" in sys.argv:
    T = 337500  # one hour at 93.75 frames/s
    mel_h = torch.from_numpy(synth.synthetic_mel(1, 100, T, 79))[0].to(dev)
    for _ in range(2):
        S.vocode_long_distributed(m, mel_h, 256, chunk_frames=4096, batch_chunks=16)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    y = S.vocode_long_distributed(m, mel_h, 256, chunk_frames=4096, batch_chunks=16)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    res["hour_ms"] = float(ms.item())
    res["hour_audio_s_per_s"] = 3600.0 / (float(ms.item()) / 1e3)
    res["hour_samples"] = int(y.numel())
if rank == 0:
    print(json.dumps(res), flush=True)
if world > 1:
    dist.destroy_process_group()
