#!/usr/bin/env python
"""Headline benchmark: BigVGAN vocoder audio-seconds per wall-second on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A *step* is one Generator.forward over one batch of synthetic mels.  Workload at every N =
BASELINE.json configs[1] per GPU: repo generator (config/config.json), fp32 path, batch 16 x 938 frames
(10.005 s each); for N>1 every rank runs its own such batch (weak scaling) and the waveforms travel to
rank 0 inside the step (NCCL gather on a side stream, double-buffered, so step i's gather runs under step
i+1's forward; everything is drained before the closing event).  One JSON line on stdout:

* ``value``  : whole-job audio-s/s with the mel batch already resident in HBM;
* ``e2e``    : the same through the reference-facing call ``vocoder_inference`` with pinned HOST
               buffers (H2D of the mels + D2H of the waveform inside the timed region);
* ``roofline``: tensor-pipe roofline of the conv (tcgen05 tap-GEMM) kernel class, the dominant one;
               ``roofline_amp``: HBM roofline of the fused Activation1d kernel class;
* ``parity`` : item 7 of rank 0's batch against the unmodified reference's waveform for that mel
               (tests/golden/bench_item.npz): the gates of north_star measured on the benchmarked run;
* ``cpu_baseline``: the PyTorch-CPU port of the reference path (oracle/bigvgan_torch_cpu.py) timed on
               this box's host cores on a bounded sample;  ``torch_eager_gpu``: the same port run by
               PyTorch eager (cuDNN) on this GPU -- the "same box" baseline -- with ``vs_eager``;
* ``bf16``   : the same batch on the bf16 path (configs[2] precision), device-resident;
* N > 1 only: ``bf16_b128`` (configs[2]: 128 items in total, data-parallel), ``hour`` (configs[3]: one
               1-hour mel time-sharded with halo + cross-fade, gathered to rank 0), ``v2`` (configs[4]:
               512x generator, 64 x 30 s in total).

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
TRAFFIC_FILES = ("r02_traffic.json", "r01_traffic_v18.json")


def load_traffic(precision, kernel):
    """DRAM bytes per launch of a kernel class from the committed ncu launch list (None if absent)."""
    for name in TRAFFIC_FILES:
        try:
            return json.load(open(os.path.join(ROOT, "profiles", name)))[precision][kernel]["dram_bytes_per_launch"]
        except Exception:
            continue
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
            "rate_is": "per-item CPU rate: one batch item per step, not the batch",
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
        "config": {"workload": WORKLOAD, "note": "reference path = PyTorch on CPU (the reference ships no GPU kernels); per-item CPU rate: each step is a bounded sample, "
                                                 "one batch item of the workload, and the rate is audio-seconds per second so it compares with the batch rate"},
        "cpu_baseline": base, "e2e": {"value": base["value"], "unit": "audio_s_per_s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


WORKLOAD = f"configs[1]: repo BigVGAN generator (112.4M params), fp32 path, batch {BATCH_PER_GPU} x {FRAMES_PER_ITEM} frames (10.005 s @24 kHz) per GPU"


def snr_db(ref, y):
    import numpy as np

    ref, y = np.asarray(ref, np.float64), np.asarray(y, np.float64)
    return float(10 * np.log10((ref**2).sum() / max(((y - ref) ** 2).sum(), 1e-300)))


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
    ap.add_argument("--no-eager", action="store_true", help="skip the PyTorch-eager-on-this-GPU baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="N > 1: skip the configs[2..4] legs (bf16 B128, 1-hour, v2)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        return run_reference(args, rank, world)

    import numpy as np
    import torch
    import torch.distributed as dist

    from svc_inference_pipeline_b200 import sharding as S
    from svc_inference_pipeline_b200.modules.bigvgan import Generator
    from svc_inference_pipeline_b200.modules.bigvgan_inference import vocoder_inference
    from svc_inference_pipeline_b200.utils import synth
    from svc_inference_pipeline_b200.utils.util import JsonHParams, load_config

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    cfg = load_config(os.path.join(ROOT, "svc_inference_pipeline_b200", "config", "config.json"))
    vcd = {k: cfg.vocoder[k] for k in cfg.vocoder.keys()}
    B, T = args.batch, args.frames
    steps, warmup = max(1, args.steps), max(3, args.warmup)
    peaks = load_peaks()

    model = Generator(cfg.vocoder, precision=args.precision)
    sd = synth.synthetic_state_dict(vcd, seed=0)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    del sd
    model = model.to(dev).eval()

    mel_host = torch.from_numpy(synth.synthetic_mel(B, 100, T, seed=1235 + rank)).pin_memory()
    mel_dev = mel_host.to(dev)
    audio_s_per_step = B * T * HOP / FS

    # ---- waveforms back to the caller: gather to rank 0 on a side stream, double-buffered ------------------
    side = torch.cuda.Stream(device=dev) if world > 1 else None
    send = [torch.empty(B, 1, T * HOP, dtype=torch.float32, device=dev) for _ in range(2)] if world > 1 else None
    recv = [[torch.empty(B, 1, T * HOP, dtype=torch.float32, device=dev) for _ in range(world)] for _ in range(2)] if (world > 1 and rank == 0) else None
    copied = [None, None]
    counter = [0]

    def step():
        i = counter[0] & 1
        counter[0] += 1
        cur = torch.cuda.current_stream(dev)
        if world > 1 and copied[i ^ 1] is not None:
            cur.wait_event(copied[i ^ 1])  # the previous step's waveform has left the program's output buffer
        y = model.forward_borrowed(mel_dev)  # the program's own output buffer: consumed below / by the caller at once
        if world > 1:
            done = cur.record_event()
            with torch.cuda.stream(side):
                side.wait_event(done)
                send[i].copy_(y, non_blocking=True)
                copied[i] = side.record_event()
                dist.gather(send[i], recv[i] if rank == 0 else None, dst=0)  # runs under the next step's forward
        return y

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def over_ranks(ms):
        """max over ranks + every rank's own figure"""
        if world == 1:
            return ms, [ms]
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        allt = torch.empty(world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allt, t)
        v = [float(x) for x in allt.cpu()]
        return max(v), v

    def timed(fn, n, warm=1):
        """device time of n calls of fn after `warm` untimed ones; max over ranks (ms per call)"""
        for _ in range(warm):
            fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        barrier()
        return over_ranks(a.elapsed_time(b) / n)[0]

    for _ in range(warmup):
        step()
    barrier()

    # ---- timed region: device-resident inputs -------------------------------------------------
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        barrier()
        e0.record()
        for _ in range(steps):
            y_last = step()
        if world > 1:
            torch.cuda.current_stream(dev).wait_stream(side)  # the closing event sees every gather
        e1.record()
        barrier()
    ms, per_rank = over_ranks(e0.elapsed_time(e1))
    value = world * audio_s_per_step * steps / (ms / 1e3)
    # the overlapped forward issues two half-batch programs (two streams) per step
    launches = model.launches_per_forward(B, T) * (2 if model.overlaps(B) else 1) * steps
    y_item7 = y_last[7:8].cpu().numpy() if (rank == 0 and B > 7) else None

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
    e2e_ms, e2e_per_rank = over_ranks((time.perf_counter() - t0) * 1e3)
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
    passes = 3 if args.precision == "fp32" else 1
    roofline = {"bound": "tensor", "kernel": f"tcgen05 tap-GEMM convolutions: conv_pair_kernel (cta_group::2, wide layers) + conv_umma_kernel (narrow layers), {prof['conv_n']} launches/step" if args.precision != "fp32_simt" else "conv_simt_kernel",
                "achieved": conv_tflops, "peak": peaks["tensor"], "unit": "TFLOP/s", "frac": conv_tflops / peaks["tensor"],
                "traffic": load_traffic("bf16" if args.precision == "bf16" else "fp32", "conv_umma_kernel"),
                "issued_tflops": conv_tflops * passes, "issued_frac": conv_tflops * passes / peaks["tensor"],
                "peak_source": peaks["src"], "avg_launch_ms": prof["conv_ms"] / max(1, prof["conv_n"]), "share_of_step": prof["conv_ms"] / prof["total_ms"],
                "note": "algorithmic FLOPs 2*Cin*Cout*K*L per conv (1.8041 GFLOP/frame); the fp32 path issues 3 bf16 MMAs per product (hi*hi + lo*hi + hi*lo), "
                        "counted once in achieved/frac and three times in issued_*; traffic = mean DRAM bytes per launch from the committed ncu launch list"}
    roofline_amp = {"bound": "hbm", "kernel": f"amp_kernel_p2 / amp_mma_kernel ({prof['amp_n']} launches/step)", "achieved": amp_gbs, "peak": peaks["hbm"], "unit": "GB/s",
                    "frac": amp_gbs / peaks["hbm"], "traffic": load_traffic("bf16" if args.precision == "bf16" else "fp32", "amp_kernel"), "avg_launch_ms": prof["amp_ms"] / max(1, prof["amp_n"]),
                    "share_of_step": prof["amp_ms"] / prof["total_ms"], "bytes_per_elem": 8 if args.precision != "bf16" else 4}

    line = {
        "metric": "bigvgan_audio_seconds_per_second", "value": value, "unit": "audio_s_per_s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"fp32": "f32 (bf16x3 split operands on tcgen05: 3 MMA passes per product, fp32 accumulate, fp32 activations)", "bf16": "bf16", "fp32_simt": "f32"}[args.precision],
        "data": "synthetic (log-mel range of the reference's mel_min/max; random-init checkpoint, seed 0)",
        "config": {"workload": WORKLOAD if (B, T) == (BATCH_PER_GPU, FRAMES_PER_ITEM) else f"custom batch {B} x {T} frames", "batch_per_gpu": B, "frames": T,
                   "precision": args.precision, "l2": "activation working set (GBs) far exceeds the 126 MB L2; no flush needed",
                   "streams": "two half-batches on two CUDA streams per rank, ops launched alternately" if model.overlaps(B) else "one stream",
                   "parallelism": f"dp{world}: independent utterances per rank" + (", NCCL gather of the waveforms to rank 0 on a side stream (double-buffered) inside the step" if world > 1 else "")},
        "e2e": e2e, "gpu_launches": launches, "clocks": clk.summary(), "roofline": roofline, "roofline_amp": roofline_amp,
        "class_ms_per_step": {k: prof[k] for k in ("conv_ms", "amp_ms", "other_ms", "total_ms")},
    }
    if world > 1:
        srt = sorted(per_rank)
        line["per_rank_ms"] = {"min": srt[0] / steps, "median": srt[len(srt) // 2] / steps, "max": srt[-1] / steps, "all": [v / steps for v in per_rank],
                               "e2e_all": [v / e2e_steps for v in e2e_per_rank]}

    # ---- parity of the benchmarked run against the unmodified reference (rank 0's item 7) -----------------
    gold_path = os.path.join(ROOT, "tests", "golden", "bench_item.npz")
    gold = np.load(gold_path) if (rank == 0 and os.path.exists(gold_path) and (B, T) == (BATCH_PER_GPU, FRAMES_PER_ITEM) and args.precision == "fp32") else None
    if gold is not None and y_item7 is not None:
        line["parity"] = {"what": "item 7 of rank 0's timed batch vs the unmodified reference (tests/golden/bench_item.npz: reference fp32 and fp64 waveforms of that mel)",
                          "fp32_path_max_abs_vs_ref_fp64": float(np.abs(y_item7 - gold["y_f64"]).max()), "fp32_path_max_abs_vs_ref_fp32": float(np.abs(y_item7 - gold["y"]).max()),
                          "gate_fp32_max_abs": 1e-4, "reference_fp32_vs_its_fp64": float(gold["ref_fp32_vs_fp64"])}

    if rank == 0 and not args.no_bf16 and args.precision == "fp32":
        model.set_precision("bf16")
        for _ in range(2):
            yb = model.forward_borrowed(mel_dev)
        torch.cuda.synchronize(dev)
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        nb = max(2, min(steps, 5))
        for _ in range(nb):
            yb = model.forward_borrowed(mel_dev)
        b1.record()
        torch.cuda.synchronize(dev)
        bms = b0.elapsed_time(b1) / nb
        bprof = model.profile_classes(B, T, reps=2)
        line["bf16"] = {"value_per_gpu": audio_s_per_step / (bms / 1e3), "ms_per_step": bms,
                        "conv_tflops": conv_flops / (bprof["conv_ms"] / 1e3) / 1e12, "conv_frac": conv_flops / (bprof["conv_ms"] / 1e3) / 1e12 / peaks["tensor"],
                        "amp_gbs": AMP_ELEMS_PER_FRAME * frames * 4 / (bprof["amp_ms"] / 1e3) / 1e9,
                        "amp_frac": AMP_ELEMS_PER_FRAME * frames * 4 / (bprof["amp_ms"] / 1e3) / 1e9 / peaks["hbm"],
                        "class_ms_per_step": {k: bprof[k] for k in ("conv_ms", "amp_ms", "other_ms", "total_ms")}}
        if gold is not None:
            from svc_inference_pipeline_b200.utils.mel import log_mel_l1  # the reference's analysis (utils/mel.py:130-174) on the device

            ref_dev = torch.from_numpy(gold["y_f64"].astype(np.float32)).to(dev)
            line["parity"].update({"bf16_path_snr_db": snr_db(gold["y_f64"], yb[7:8].cpu().numpy()), "gate_bf16_snr_db": 35.0,
                                   "bf16_path_log_mel_l1": log_mel_l1(ref_dev, yb[7:8]), "gate_bf16_log_mel_l1": 1e-2})
        model.set_precision(args.precision)

    # ---- the two other checkpoint recipes of utils/synth.py against the reference on them (tests/golden/recipes.npz) ---
    rec_path = os.path.join(ROOT, "tests", "golden", "recipes.npz")
    if rank == 0 and world == 1 and "parity" in line and os.path.exists(rec_path) and not args.no_extra:
        from svc_inference_pipeline_b200.utils.mel import log_mel_l1

        rec = np.load(rec_path)
        xr = torch.from_numpy(rec["mel"]).to(dev)
        line["parity"]["recipes"] = {"what": "repo generator with the 'survey' (SURVEY 8d as written) and 'large_alpha' checkpoint recipes, 96 log-mel frames, vs the unmodified "
                                             "reference's fp64 waveform; the fp32 gate (1e-4) must hold, the bf16 figures are reported (both nets amplify rounding: DESIGN.md section 5)"}
        for recipe in ("survey", "large_alpha"):
            mr = Generator(cfg.vocoder, precision="fp32")
            mr.load_state_dict({k: torch.from_numpy(v) for k, v in synth.synthetic_state_dict(vcd, seed=0, recipe=recipe).items()})
            mr = mr.to(dev).eval()
            ref64 = rec[recipe + "_y_f64"]
            y32 = mr(xr).cpu().numpy()
            mr.set_precision("bf16")
            yb = mr(xr)
            line["parity"]["recipes"][recipe] = {
                "fp32_path_max_abs_vs_ref_fp64": float(np.abs(y32 - ref64).max()), "reference_fp32_vs_its_fp64": float(np.abs(rec[recipe + "_y"] - ref64).max()),
                "bf16_path_snr_db": snr_db(ref64, yb.cpu().numpy()),
                "bf16_path_log_mel_l1": log_mel_l1(torch.from_numpy(ref64.astype(np.float32)).to(dev), yb)}
            del mr
        torch.cuda.empty_cache()

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

    if rank == 0 and world == 1 and not args.no_extra:
        # one DiffSVC denoiser step (the function infer.py's sampler calls 1000 times per utterance),
        # reference mapper hyper-parameters, CUDA-graph replay, device-resident inputs
        from svc_inference_pipeline_b200.modules.diffsvc import DiffSVC

        mcfg = dict(noise_schedule_factors=[0.0001, 0.02, 1000], n_mel=100, residual_channels=384, diffusion_fc_size=128, conditioner_size=384,
                    dilation_cycle_length=4, residual_kernel_size=3, residual_layer_num=20)
        dm = DiffSVC(JsonHParams(**mcfg), precision="fp32")
        dm.load_state_dict({k: torch.from_numpy(v) for k, v in synth.synthetic_diffsvc_state_dict(mcfg, seed=3).items()})
        dm = dm.to(dev).eval()
        dstep = {"what": "DiffSVC.forward(mel[B,L,100], cond[B,L,384], t[B,1]) on libbvg_b200 (20 dilated layers, C=384), one CUDA-graph replay per step; ms per step",
                 "launches_per_step": dm.launches_per_step(1, 379)}
        for prec in ("fp32", "bf16"):
            dm.set_precision(prec)
            for (b_, l_) in ((1, 379), (16, 938)):
                xm = torch.randn(b_, l_, 100, device=dev)
                xc = torch.randn(b_, l_, 384, device=dev)
                tt = torch.full((b_, 1), 500, dtype=torch.long, device=dev)
                for _ in range(3):
                    dm(xm, xc, tt)
                torch.cuda.synchronize(dev)
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                for _ in range(20):
                    dm(xm, xc, tt)
                a1.record()
                torch.cuda.synchronize(dev)
                dstep[f"{prec}_b{b_}x{l_}_ms"] = a0.elapsed_time(a1) / 20
        if not args.no_eager:
            # same-box baseline: the functional port of the reference's DiffSVC.forward (oracle/diffsvc_oracle.py, measurement
            # only) run by PyTorch eager on this GPU, fp32 with TF32 off
            from oracle import diffsvc_oracle as DO

            sdg = {k: torch.from_numpy(v).to(dev) for k, v in synth.synthetic_diffsvc_state_dict(mcfg, seed=3).items()}
            for (b_, l_) in ((1, 379), (16, 938)):
                xm, xc = torch.randn(b_, l_, 100, device=dev), torch.randn(b_, l_, 384, device=dev)
                tt = torch.full((b_, 1), 500, dtype=torch.long, device=dev)
                for _ in range(3):
                    DO.denoiser_forward(sdg, mcfg, xm, xc, tt)
                torch.cuda.synchronize(dev)
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                for _ in range(10):
                    DO.denoiser_forward(sdg, mcfg, xm, xc, tt)
                a1.record()
                torch.cuda.synchronize(dev)
                dstep[f"torch_eager_gpu_fp32_b{b_}x{l_}_ms"] = a0.elapsed_time(a1) / 10
            del sdg
        # the sampler that drives it (modules/diffsvcrepo_inference.py::svc_model_inference, the call infer.py:79 makes): the
        # reference's 1000-step schedule on one utterance of config 1's length, host wall clock, result copied to the host
        from svc_inference_pipeline_b200.modules.diffsvcrepo_inference import svc_model_inference

        sched = np.linspace(1e-4, 0.02, 1000).tolist()
        scfg = JsonHParams(mapper=JsonHParams(noise_schedule=sched))
        cond_fn = lambda batch: batch["cond"]
        batch = {"y": torch.zeros(1, 379, 100, device=dev), "cond": torch.randn(1, 379, 384, device=dev)}
        samp = {"what": "svc_model_inference([cond, DiffSVC], batch, cfg) on one utterance of 379 frames, 1000 p_sample steps (default) / 100 PLMS steps "
                        "(fast_inference, speedup 10): seconds per utterance, host wall clock incl. the copy of the mel to the host"}
        for prec in ("fp32", "bf16"):
            dm.set_precision(prec)
            for fast in (False, True):
                svc_model_inference([cond_fn, dm], batch, JsonHParams(mapper=JsonHParams(noise_schedule=sched[:20])), fast_inference=fast, speedup=10)  # graphs
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                y = svc_model_inference([cond_fn, dm], batch, scfg, fast_inference=fast, speedup=10).cpu()
                samp[f"{prec}_{'plms100' if fast else 'ddpm1000'}_s"] = time.perf_counter() - t0
                samp[f"{prec}_{'plms100' if fast else 'ddpm1000'}_finite"] = bool(torch.isfinite(y).all())
        dstep["sampler"] = samp
        line["diffsvc_step"] = dstep
        del dm
        torch.cuda.empty_cache()

    # ---- N > 1: the other BASELINE configs as extra keys (all ranks take part) -----------------------------
    if world > 1 and not args.no_extra and args.precision == "fp32":
        extra_steps = 2
        # configs[2]: bf16 path, 128 items x 938 frames in total, data-parallel, waveforms gathered to rank 0
        per = max(1, 128 // world)
        model.set_precision("bf16")
        melb = torch.from_numpy(synth.synthetic_mel(per, 100, T, seed=2235 + rank)).to(dev)
        sendb = torch.empty(per, 1, T * HOP, dtype=torch.float32, device=dev)
        recvb = [torch.empty_like(sendb) for _ in range(world)] if rank == 0 else None

        def bf16_step():
            sendb.copy_(model.forward_borrowed(melb))
            dist.gather(sendb, recvb, dst=0)

        t_ms = timed(bf16_step, extra_steps)
        line["bf16_b128"] = {"config": f"configs[2]: bf16 path, {per * world} x {T} frames in total ({per} per GPU), gather to rank 0 inside the step",
                             "value": per * world * T * HOP / FS / (t_ms / 1e3), "unit": "audio_s_per_s", "ms_per_step": t_ms}
        del melb, sendb, recvb
        model.set_precision("fp32")
        model._invalidate()
        torch.cuda.empty_cache()
        # configs[3]: one 1-hour mel, time-sharded with a 48-frame halo, cross-faded on the device, gathered to rank 0
        T_hour = 337500
        mel_h = torch.from_numpy(synth.synthetic_mel(1, 100, T_hour, 79))[0].to(dev)
        hour = {"config": "configs[3]: 1-hour mel [100, 337500], balanced 4096-frame-class chunks + 48-frame halo, bvg_stitch_fwd cross-fade, NCCL gather to rank 0"}
        for prec in ("fp32", "bf16"):
            model.set_precision(prec)
            t_ms = timed(lambda: S.vocode_long_distributed(model, mel_h, HOP, chunk_frames=4096, batch_chunks=16, gather="root"), extra_steps)
            hour[prec] = {"value": 3600.0 / (t_ms / 1e3), "unit": "audio_s_per_s", "ms": t_ms}
            model._invalidate()
            torch.cuda.empty_cache()
        model.set_precision("fp32")
        plan = S.balanced_plan(T_hour, 4096, world=world)
        hour["chunks"], hour["chunk_frames"], hour["halo_overhead"] = len(plan), plan[0].end - plan[0].start, (plan[0].in_hi - plan[0].in_lo) / (plan[0].end - plan[0].start) - 1.0
        line["hour"] = hour
        del mel_h
        # configs[4]: 512x v2-style generator (input_dim 128, rates 8,4,2,2,2,2; hop 512 @ 44.1 kHz), 64 x 30 s in total
        v2cfg = dict(vcd, input_dim=128, upsample_rates=[8, 4, 2, 2, 2, 2], upsample_kernel_sizes=[16, 8, 4, 4, 4, 4])
        m2 = Generator(JsonHParams(**v2cfg), precision="fp32")
        m2.load_state_dict({k: torch.from_numpy(v) for k, v in synth.synthetic_state_dict(v2cfg, seed=0).items()})
        m2 = m2.to(dev).eval()
        per2, T2 = max(1, 64 // world), 2584
        sub = min(8, per2)
        mel2 = torch.from_numpy(synth.synthetic_mel(per2, 128, T2, seed=3235 + rank)).to(dev)
        send2 = torch.empty(per2, 1, T2 * 512, dtype=torch.float32, device=dev)
        recv2 = [torch.empty_like(send2) for _ in range(world)] if rank == 0 else None

        def v2_step():
            for a in range(0, per2, sub):
                send2[a : a + sub].copy_(m2.forward_borrowed(mel2[a : a + sub].contiguous()))
            dist.gather(send2, recv2, dst=0)

        v2 = {"config": f"configs[4]: 512x v2 generator (122.2M params), {per2 * world} x {T2} frames (30 s @44.1 kHz) in total, {per2} per GPU in sub-batches of {sub}, gather to rank 0"}
        for prec in ("fp32", "bf16"):
            m2.set_precision(prec)
            t_ms = timed(v2_step, extra_steps)
            v2[prec] = {"value": per2 * world * T2 * 512 / 44100 / (t_ms / 1e3), "unit": "audio_s_per_s", "ms_per_step": t_ms}
        line["v2"] = v2
        del m2, mel2, send2, recv2
        torch.cuda.empty_cache()

    # ---- baselines on the same box (rank 0, N = 1): PyTorch eager on this GPU, PyTorch on the host cores ------
    if rank == 0 and world == 1 and not args.no_eager:
        from oracle import bigvgan_torch_cpu as port  # baseline leg only: the port is what gets timed, never the product

        tsd = {k: v.detach().float() for k, v in model.state_dict().items()}
        eager = {"what": "PyTorch eager (cuDNN / ATen kernels) running the functional port of the reference path (oracle/bigvgan_torch_cpu.py: the same library "
                         f"calls per layer as modules/bigvgan.py, weight-norm recomputed per forward) on this GPU, batch {B} x {T} frames, device-resident, 1 warm-up + 2 timed",
                 "unit": "audio_s_per_s"}

        def eager_rate(fn):
            fn()
            torch.cuda.synchronize(dev)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(2):
                out = fn()
            b.record()
            torch.cuda.synchronize(dev)
            return audio_s_per_step / (a.elapsed_time(b) / 2 / 1e3), out

        tf32_conv, tf32_mm = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
        try:
            torch.backends.cudnn.allow_tf32 = False
            torch.backends.cuda.matmul.allow_tf32 = False
            eager["fp32"], y_e = eager_rate(lambda: port.generator_forward(tsd, cfg.vocoder, mel_dev))
            if gold is not None:
                eager["fp32_max_abs_vs_ref_fp64"] = float(np.abs(y_e[7:8].cpu().numpy() - gold["y_f64"]).max())
            torch.backends.cudnn.allow_tf32 = True
            eager["tf32_convs"], y_e = eager_rate(lambda: port.generator_forward(tsd, cfg.vocoder, mel_dev))
            if gold is not None:
                eager["tf32_max_abs_vs_ref_fp64"] = float(np.abs(y_e[7:8].cpu().numpy() - gold["y_f64"]).max())

            def autocast_fwd():
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    return port.generator_forward(tsd, cfg.vocoder, mel_dev)

            eager["autocast_bf16"], y_e = eager_rate(autocast_fwd)
            if gold is not None:
                eager["autocast_bf16_snr_db"] = snr_db(gold["y_f64"], y_e[7:8].float().cpu().numpy())
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32_conv, tf32_mm
        del tsd, y_e
        torch.cuda.empty_cache()
        line["torch_eager_gpu"] = eager
        line["vs_eager"] = {"fp32_path_vs_eager_fp32": value / eager["fp32"], "fp32_path_vs_eager_tf32": value / eager["tf32_convs"]}
        if "bf16" in line:
            line["vs_eager"]["bf16_path_vs_eager_autocast_bf16"] = line["bf16"]["value_per_gpu"] / eager["autocast_bf16"]

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_port_rate(steps=2, warmup=1, budget_s=30.0)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
