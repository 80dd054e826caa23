"""Where the error of the 3-pass split convolution comes from: operand representation (hi + lo of 8 bits each) or the
tensor core's fp32 accumulation.  One dense layer (C -> C, k taps), random data:
  total  = kernel vs exact fp64 conv of the fp32 operands
  accum  = kernel vs fp64 conv of the SAME split operands (hi*hi + lo*hi + hi*lo): what the accumulator loses
  repr   = fp64 conv of the split operands vs exact: what the operand format loses
    python tools/conv_precision_diag.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from svc_inference_pipeline_b200 import _lib as L, ops
from util_cases import bf16_round

def conv64(x, w, k, d):  # x [C, L] f64, w [Co, Ci, k] f64, "same" zero padding
    C, Ln = x.shape
    pad = (k - 1) * d // 2
    xp = np.pad(x, ((0, 0), (pad, pad)))
    y = np.zeros((w.shape[0], Ln))
    for j in range(k):
        y += w[:, :, j] @ xp[:, j * d : j * d + Ln]
    return y

for C, k, Ln in [(768, 11, 512), (384, 11, 512), (96, 7, 1024), (24, 11, 2048)]:
    rng = np.random.default_rng(C + k)
    x = rng.standard_normal((1, C, Ln)).astype(np.float32)
    v = (rng.uniform(-1, 1, (C, C, k)) / np.sqrt(C * k)).astype(np.float32)
    g = np.sqrt((v.astype(np.float64) ** 2).sum(axis=(1, 2), keepdims=True)).astype(np.float32)
    b = np.zeros(C, np.float32)
    w = (g.astype(np.float64) * v / np.sqrt((v.astype(np.float64) ** 2).sum(axis=(1, 2), keepdims=True))).astype(np.float32)  # fold in fp64 -> fp32, as pack.cu does up to rounding
    pc = ops.pack_conv(*[torch.from_numpy(t).cuda() for t in (v, g, b)], dilation=1, padding=(k - 1) // 2, backend=L.UMMA, split=True)
    xl = torch.from_numpy(np.ascontiguousarray(x.transpose(0, 2, 1))).cuda()
    y = ops.conv(xl, pc).cpu().numpy()[0].T.astype(np.float64)
    exact = conv64(x[0].astype(np.float64), w.astype(np.float64), k, 1)
    xh = bf16_round(x[0]); xlo = bf16_round((x[0] - xh).astype(np.float32))
    wh = bf16_round(w); wlo = bf16_round((w - wh).astype(np.float32))
    rep = conv64(xh.astype(np.float64), wh.astype(np.float64), k, 1) + conv64(xlo.astype(np.float64), wh.astype(np.float64), k, 1) + conv64(xh.astype(np.float64), wlo.astype(np.float64), k, 1)
    rms = np.sqrt((exact ** 2).mean())
    f = lambda a: f"max {np.abs(a).max() / rms:.2e} rms {np.sqrt((a ** 2).mean()) / rms:.2e} mean {a.mean() / rms:+.1e}"
    print(f"C={C} k={k} K={C * k}: total [{f(y - exact)}]  accum [{f(y - rep)}]  repr [{f(rep - exact)}]  (relative to output rms {rms:.3f}); "
          f"fp32 FFMA-order reference: {np.abs(conv64(x[0].astype(np.float64), w.astype(np.float64), k, 1).astype(np.float32) - exact).max() / rms:.1e}", flush=True)
