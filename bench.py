#!/usr/bin/env python
"""Headline benchmark: BigVGAN vocoder audio-seconds per wall-second on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A *step* is one Generator.forward over one batch of synthetic mels.  Workload at N=1 =
BASELINE.json configs[1]: repo generator (config/config.json), fp32 path, batch 16 x 938 frames
(10.005 s each); for N>1 every rank runs that same batch (weak scaling: configs[2]'s batch 128 is
8 x 16) and the waveforms are gathered to rank 0 inside the step.  One JSON line on stdout:

* ``value``  : whole-job audio-s/s with the mel batch already resident in HBM;
* ``e2e``    : the same through the reference-facing call ``vocoder_inference`` with pinned HOST
               buffers (H2D of the mels + D2H of the waveform inside the timed region);
* ``roofline``: tensor-pipe roofline of the conv (tcgen05 tap-GEMM) kernel class, the dominant one;
               ``roofline_amp``: HBM roofline of the fused Activation1d kernel class;
* ``cpu_baseline``: the PyTorch-CPU port of the reference path (oracle/bigvgan_torch_cpu.py) timed on
               this box's host cores on a bounded sample;
* ``bf16``   : the same batch on the bf16 path (configs[2] precision), device-resident.

``--impl reference`` times only the CPU port (the reference has no GPU kernels of its own and its
Python cannot travel to the GPU box) and prints the same line with impl=reference.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAMES_PER_ITEM = 938
BATCH_PER_GPU = 16
FS, HOP = 24000, 256
GFLOP_PER_FRAME = 1.8041      # dense-conv FLOPs per mel frame, 3-pass split counted once
AMP_ELEMS_PER_FRAME = 614_400  # Activation1d elements per mel frame over the 109 calls


def load_traffic(precision, kernel):
    """DRAM bytes per launch of a kernel class from the committed ncu launch list (None if absent)."""
    path = os.path.join(ROOT, "profiles", "r01_traffic_v18.json")
    try:
        return json.load(open(path))[precision][kernel]["dram_bytes_per_launch"]
    except Exception:
        return None


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tensor=p.get("bf16_tflops_sustained", p["bf16_tflops"]), src="measured (MEASURED_PEAKS.json; sustained bf16, copy GB/s)")
    return dict(hbm=6650.0, tensor=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(self.rows)}


def cpu_port_rate(steps=3, warmup=1, budget_s=40.0):
    """Audio-s/s of the PyTorch-CPU port of the reference path on one batch item of the workload
    ([1,100,938] = 10.005 s of audio), all host threads.  `steps` timed runs after `warmup` untimed
    ones, cut short when `budget_s` of wall clock is spent."""
    import torch

    from oracle import bigvgan_torch_cpu as port
    from svc_inference_pipeline_b200.utils import synth
    from svc_inference_pipeline_b200.utils.util import load_config

    cfg = load_config(os.path.join(ROOT, "svc_inference_pipeline_b200", "config", "config.json"))
    vc = {k: cfg.vocoder[k] for k in cfg.vocoder.keys()}
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = {k: torch.from_numpy(v) for k, v in synth.synthetic_state_dict(vc, seed=0).items()}
    t_begin = time.perf_counter()
    mel_small = torch.from_numpy(synth.synthetic_mel(1, 100, 128, seed=1236))
    t0 = time.perf_counter()
    port.vocoder_inference(sd, vc, mel_small)  # first touch: oneDNN primitive caches, page-in
    t_small = time.perf_counter() - t0
    frames = FRAMES_PER_ITEM if t_small * (FRAMES_PER_ITEM / 128) * 2 < budget_s else 256
    mel = torch.from_numpy(synth.synthetic_mel(1, 100, frames, seed=1236))
    times = []
    for i in range(max(0, warmup - 1) + max(1, steps)):
        t0 = time.perf_counter()
        port.vocoder_inference(sd, vc, mel)
        dt = time.perf_counter() - t0
        if i >= max(0, warmup - 1):
            times.append(dt)
        if time.perf_counter() - t_begin + dt > budget_s and times:
            break
    audio_s = frames * HOP / FS
    mean = sum(times) / len(times)
    return {"value": audio_s / mean, "unit": "audio_s_per_s", "cores": cores, "kind": "port", "steps_run": len(times), "ms_per_step": mean * 1e3,
            "sample": f"1 item [1,100,{frames}] ({audio_s:.2f} s audio) of the workload per step, PyTorch-CPU port of the reference path "
                      f"(oracle/bigvgan_torch_cpu.py), {len(times)} timed steps, mean {mean:.2f} s, best {min(times):.2f} s"}


def run_reference(args, rank, world):
    """--impl reference: the reference's own implementation of this path is PyTorch on CPU (it has no
    GPU kernels, and its Python tree cannot travel to the GPU box), so this arm times the CPU port on
    the host cores; rank 0 only."""
    if rank != 0:
        return
    base = cpu_port_rate(steps=max(1, args.steps), warmup=max(1, args.warmup), budget_s=150.0)
    line = {
        "impl": "reference", "metric": "bigvgan_audio_seconds_per_second", "value": base["value"], "unit": "audio_s_per_s",
        "n_gpus": args.gpus, "steps": base["steps_run"], "warmup": max(1, args.warmup), "ms_per_step": base["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "reference path = PyTorch on CPU (the reference ships no GPU kernels); each step is a bounded sample: one batch item"},
        "cpu_baseline": base, "e2e": {"value": base["value"], "unit": "audio_s_per_s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


WORKLOAD = f"configs[1]: repo BigVGAN generator (112.4M params), fp32 path, batch {BATCH_PER_GPU} x {FRAMES_PER_ITEM} frames (10.005 s @24 kHz) per GPU"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16", "fp32_simt"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU)
    ap.add_argument("--frames", type=int, default=FRAMES_PER_ITEM)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-bf16", action="store_true")
    ap.add_argument("--torch-gpu-baseline", action="store_true", help="also time the PyTorch port on the GPU (eager cuDNN)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        return run_reference(args, rank, world)

    import numpy as np
    import torch
    import torch.distributed as dist

    from svc_inference_pipeline_b200.modules.bigvgan import Generator
    from svc_inference_pipeline_b200.modules.bigvgan_inference import vocoder_inference
    from svc_inference_pipeline_b200.utils import synth
    from svc_inference_pipeline_b200.utils.util import load_config

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    cfg = load_config(os.path.join(ROOT, "svc_inference_pipeline_b200", "config", "config.json"))
    B, T = args.batch, args.frames
    steps, warmup = max(1, args.steps), max(3, args.warmup)
    peaks = load_peaks()

    model = Generator(cfg.vocoder, precision=args.precision)
    sd = synth.synthetic_state_dict({k: cfg.vocoder[k] for k in cfg.vocoder.keys()}, seed=0)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    del sd
    model = model.to(dev).eval()

    mel_host = torch.from_numpy(synth.synthetic_mel(B, 100, T, seed=1235 + rank)).pin_memory()
    mel_dev = mel_host.to(dev)
    audio_s_per_step = B * T * HOP / FS
    gathered = torch.empty(world * B, 1, T * HOP, dtype=torch.float32, device=dev) if world > 1 else None

    def step():
        y = model(mel_dev)
        if world > 1:
            dist.all_gather_into_tensor(gathered, y)  # waveforms back to the caller (rank 0 reads `gathered`)
        return y

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(warmup):
        step()
    barrier()

    # ---- timed region: device-resident inputs -------------------------------------------------
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        barrier()
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    value = world * audio_s_per_step * steps / (ms / 1e3)
    # the overlapped forward issues two half-batch programs (two streams) per step
    launches = model.launches_per_forward(B, T) * (2 if (model.overlap_streams and B >= 2) else 1) * steps

    # ---- e2e: the reference-facing call with host buffers ------------------------------------------
    def e2e_step():
        return vocoder_inference(cfg, model, mel_host, dev)  # .to(device) + forward + .cpu()

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, min(steps, 5))
    for _ in range(e2e_steps):
        out_host = e2e_step()
    torch.cuda.synchronize(dev)
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    e2e = {"value": world * audio_s_per_step * e2e_steps / (e2e_ms / 1e3), "unit": "audio_s_per_s",
           "h2d_bytes_per_step": mel_host.numel() * 4, "d2h_bytes_per_step": out_host.numel() * 4,
           "api": "modules.bigvgan_inference.vocoder_inference(cfg, model, mels_cpu_pinned, device)"}

    # ---- per-kernel-class device time (CUDA events around every launch of one program run) ------------
    prof = model.profile_classes(B, T, reps=2)
    frames = B * T
    conv_flops = GFLOP_PER_FRAME * 1e9 * frames
    amp_bytes = AMP_ELEMS_PER_FRAME * frames * (8 if args.precision != "bf16" else 4)
    conv_tflops = conv_flops / (prof["conv_ms"] / 1e3) / 1e12
    amp_gbs = amp_bytes / (prof["amp_ms"] / 1e3) / 1e9
    roofline = {"bound": "tensor", "kernel": "conv_umma_kernel (116 launches/step)" if args.precision != "fp32_simt" else "conv_simt_kernel",
                "achieved": conv_tflops, "peak": peaks["tensor"], "unit": "TFLOP/s", "frac": conv_tflops / peaks["tensor"],
                "traffic": load_traffic("bf16" if args.precision == "bf16" else "fp32", "conv_umma_kernel"),
                "issued_tflops": conv_tflops * (3 if args.precision == "fp32" else 1), "issued_frac": conv_tflops * (3 if args.precision == "fp32" else 1) / peaks["tensor"],
                "peak_source": peaks["src"], "avg_launch_ms": prof["conv_ms"] / max(1, prof["conv_n"]), "share_of_step": prof["conv_ms"] / prof["total_ms"],
                "note": "algorithmic FLOPs 2*Cin*Cout*K*L per conv (1.8041 GFLOP/frame); the fp32 path issues 3 bf16 MMAs per product (hi*hi + lo*hi + hi*lo), "
                        "counted once in achieved/frac and three times in issued_*; ncu: 97 % / 86 % tensor-pipe active on the C=768 / C=384 layers (profiles/r01_ncu_summary_v14.md); "
                        "traffic = mean DRAM bytes per launch from the committed ncu launch list"}
    roofline_amp = {"bound": "hbm", "kernel": "amp_kernel_p2 / amp_mma_kernel (109 launches/step)", "achieved": amp_gbs, "peak": peaks["hbm"], "unit": "GB/s",
                    "frac": amp_gbs / peaks["hbm"], "traffic": load_traffic("bf16" if args.precision == "bf16" else "fp32", "amp_kernel"), "avg_launch_ms": prof["amp_ms"] / max(1, prof["amp_n"]),
                    "share_of_step": prof["amp_ms"] / prof["total_ms"], "bytes_per_elem": 8 if args.precision != "bf16" else 4}

    line = {
        "metric": "bigvgan_audio_seconds_per_second", "value": value, "unit": "audio_s_per_s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"fp32": "f32 (bf16x3 split operands on tcgen05, fp32 accumulate, fp32 activations)", "bf16": "bf16", "fp32_simt": "f32"}[args.precision],
        "data": "synthetic (log-mel range of the reference's mel_min/max; random-init checkpoint, seed 0)",
        "config": {"workload": WORKLOAD if (B, T) == (BATCH_PER_GPU, FRAMES_PER_ITEM) else f"custom batch {B} x {T} frames", "batch_per_gpu": B, "frames": T,
                   "precision": args.precision, "l2": "activation working set (GBs) far exceeds the 126 MB L2; no flush needed",
                   "streams": "two half-batches on two CUDA streams per rank, ops launched alternately" if (model.overlap_streams and B >= 2) else "one stream",
                   "parallelism": f"dp{world}: independent utterances per rank" + (", all_gather of waveforms inside the step" if world > 1 else "")},
        "e2e": e2e, "gpu_launches": launches, "clocks": clk.summary(), "roofline": roofline, "roofline_amp": roofline_amp,
        "class_ms_per_step": {k: prof[k] for k in ("conv_ms", "amp_ms", "other_ms", "total_ms")},
    }

    if rank == 0 and not args.no_bf16 and args.precision == "fp32":
        model.set_precision("bf16")
        for _ in range(2):
            model(mel_dev)
        torch.cuda.synchronize(dev)
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        nb = max(2, min(steps, 5))
        for _ in range(nb):
            model(mel_dev)
        b1.record()
        torch.cuda.synchronize(dev)
        bms = b0.elapsed_time(b1) / nb
        bprof = model.profile_classes(B, T, reps=2)
        line["bf16"] = {"value_per_gpu": audio_s_per_step / (bms / 1e3), "ms_per_step": bms,
                        "conv_tflops": conv_flops / (bprof["conv_ms"] / 1e3) / 1e12, "conv_frac": conv_flops / (bprof["conv_ms"] / 1e3) / 1e12 / peaks["tensor"],
                        "amp_gbs": AMP_ELEMS_PER_FRAME * frames * 4 / (bprof["amp_ms"] / 1e3) / 1e9,
                        "amp_frac": AMP_ELEMS_PER_FRAME * frames * 4 / (bprof["amp_ms"] / 1e3) / 1e9 / peaks["hbm"],
                        "class_ms_per_step": {k: bprof[k] for k in ("conv_ms", "amp_ms", "other_ms", "total_ms")}}
        model.set_precision(args.precision)

    if rank == 0:
        # configs[0] shape: one utterance of 379 frames (4.04 s) through synthesis_audios, eager and as a CUDA graph
        from svc_inference_pipeline_b200.modules.bigvgan_inference import synthesis_audios
        m1 = torch.from_numpy(synth.synthetic_mel(1, 100, 379, seed=1234))[0]
        lat = {}
        for graph in (False, True):
            model.use_cuda_graph = graph
            for _ in range(3):
                synthesis_audios(model, m1, cfg)
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            for _ in range(5):
                synthesis_audios(model, m1, cfg)
            lat["cuda_graph_ms" if graph else "eager_ms"] = (time.perf_counter() - t0) / 5 * 1e3
        model.use_cuda_graph = False
        lat["audio_s"] = 379 * HOP / FS
        lat["api"] = "synthesis_audios(model, mel[100,379], cfg): H2D + forward + D2H + fade, wall clock"
        line["single_utterance_latency"] = lat

    if rank == 0 and args.torch_gpu_baseline:
        from oracle import bigvgan_torch_cpu as port
        tsd = {k: v.detach() for k, v in model.state_dict().items()}
        m1 = mel_dev[:4].contiguous()
        port.generator_forward(tsd, cfg.vocoder, m1)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        port.generator_forward(tsd, cfg.vocoder, m1)
        torch.cuda.synchronize(dev)
        line["torch_eager_gpu"] = {"value": 4 * T * HOP / FS / (time.perf_counter() - t0), "unit": "audio_s_per_s", "sample": f"PyTorch eager (cuDNN) port, fp32, batch 4 x {T}"}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_port_rate(steps=2, warmup=1, budget_s=30.0)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
