"""Emulate the tensor-core operand formats on the CPU oracle to choose the fp32-parity scheme.
(Experiment script; results recorded in DESIGN.md.)"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import bigvgan_oracle as O
from svc_inference_pipeline_b200.utils import synth

def bf16(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32)

def tf32(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    r = ((u + 0xFFF + ((u >> 13) & 1)) >> 13) << 13
    return r.astype(np.uint32).view(np.float32)

MODE = None
_conv1d, _convT = O.conv1d, O.conv_transpose1d

def split(a, n):
    parts, r = [], a.astype(np.float32)
    for _ in range(n):
        p = bf16(r); parts.append(p); r = (r - p).astype(np.float32)
    return parts

def mm_conv(fn, x, w, b, *args):
    if MODE == "fp32":
        return fn(x, w, b, *args)
    if MODE == "tf32":
        return fn(tf32(x).astype(np.float64), tf32(w).astype(np.float64), b, *args).astype(np.float32)
    if MODE == "bf16":
        return fn(bf16(x).astype(np.float64), bf16(w).astype(np.float64), b, *args).astype(np.float32)
    n, terms = {"bf16x3": (2, [(0,0),(0,1),(1,0)]), "bf16x4": (2, [(0,0),(0,1),(1,0),(1,1)]),
                "bf16x6": (3, [(0,0),(0,1),(1,0),(1,1),(0,2),(2,0)])}[MODE]
    xs, ws = split(x, n), split(w, n)
    acc = None
    for i, j in terms:
        y = fn(xs[i].astype(np.float64), ws[j].astype(np.float64), None, *args)
        acc = y if acc is None else acc + y
    acc = acc.astype(np.float32)
    if b is not None:
        acc = acc + b[None, :, None]
    return acc

O.conv1d = lambda x, w, b, dilation=1, padding=0: mm_conv(_conv1d, x, w, b, dilation, padding)
O.conv_transpose1d = lambda x, w, b, stride, padding: mm_conv(_convT, x, w, b, stride, padding)

if __name__ == "__main__":
    sys.path.insert(0, "tests")
    from test_oracle_golden import REPO
    g = np.load("tests/golden/repo_generator.npz")
    sd = synth.synthetic_state_dict(REPO, 0)
    for tag in ("logmel", "randn"):
        ref = g[tag + "_y_f64"]
        for MODE in sys.argv[1:] or ["fp32", "bf16x3", "bf16x4", "tf32", "bf16"]:
            globals()["MODE"] = MODE
            y = O.generator_forward(sd, REPO, g[tag + "_mel"])
            err = np.abs(y - ref).max()
            snr = 10 * np.log10((ref ** 2).sum() / ((y - ref) ** 2).sum())
            print(f"{tag:7s} {MODE:7s} max-abs {err:.3e}  SNR {snr:.1f} dB", flush=True)
