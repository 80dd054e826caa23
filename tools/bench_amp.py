"""Micro-benchmark of the fused Activation1d kernel over the generator's stage shapes and tuning knobs.
    python tools/bench_amp.py [--batch 16] [--frames 938]"""
import argparse, os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from svc_inference_pipeline_b200 import _lib as L
from svc_inference_pipeline_b200.modules.bigvgan import _Buf
from svc_inference_pipeline_b200.utils import synth

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--frames", type=int, default=938)
a = ap.parse_args()
dev = torch.device("cuda:0")
lib = L.lib()
taps = synth.aa_filter_taps()
shapes = [(768, 4), (384, 16), (96, 64), (24, 256)]
modes = [("f32->split precise", L.F32, L.SPLIT, 0), ("f32->f32 precise", L.F32, L.F32, 0), ("bf16->bf16 fast", L.BF16, L.BF16, 1), ("f32->split fast", L.F32, L.SPLIT, 1)]
esz = {L.F32: 4, L.BF16: 2, L.SPLIT: 4}
for ch, up in shapes:
    Ln = a.frames * up
    n = a.batch * Ln * ch
    a_par = torch.rand(ch, device=dev) + 0.5
    invb = torch.rand(ch, device=dev) + 0.5
    for name, idt, odt, fast in modes:
        x = _Buf(idt, n, dev); y = _Buf(odt, n, dev)
        x.hi.copy_(torch.randn(n, device=dev).to(x.hi.dtype))
        d = L.AmpDesc()
        d.x, d.y = x.tensor(), y.tensor()
        d.d_a, d.d_invb = a_par.data_ptr(), invb.data_ptr()
        d.taps_up = (C.c_float * 12)(*taps.tolist()); d.taps_down = (C.c_float * 12)(*taps.tolist())
        d.B, d.L, d.C, d.fast_sin = a.batch, Ln, ch, fast
        for vec in (4, 2):
            for chunk in (0, 8, 4, 2):
                L.set_tuning("amp_vec", vec); L.set_tuning("amp_chunk", chunk)
                st = torch.cuda.current_stream().cuda_stream
                for _ in range(3):
                    L.check(lib.bvg_amp_fwd(C.byref(d), st))
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    L.check(lib.bvg_amp_fwd(C.byref(d), st))
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 10
                gbs = n * (esz[idt] + esz[odt]) / ms / 1e6
                print(f"C{ch:4d} L{Ln:6d} {name:20s} vec{vec} chunk{chunk}: {ms*1e3:8.1f} us {gbs:7.0f} GB/s {n/ms/1e6:6.1f} Gelem/s", flush=True)
L.set_tuning("amp_vec", 0); L.set_tuning("amp_chunk", 0)
