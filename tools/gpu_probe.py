"""First-contact probe for a GPU box: runs each kernel class in a subprocess (a trapping kernel
kills its CUDA context) and prints error metrics instead of asserting.  Usage:
    python tools/gpu_probe.py [case ...]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = ["amp", "simt", "umma_mb1", "umma_mb2", "umma_mb4", "umma_split", "umma_convT", "tiny_gen"]


def run_case(name):
    import numpy as np
    import torch

    from oracle import bigvgan_oracle as O
    from svc_inference_pipeline_b200 import _lib as L
    from svc_inference_pipeline_b200 import ops
    from svc_inference_pipeline_b200.utils import synth
    from util_cases import bf16_round

    dev = "cuda:0"
    cl = lambda x: torch.from_numpy(np.ascontiguousarray(np.transpose(x, (0, 2, 1)))).to(dev)
    cf = lambda t: np.transpose(t.cpu().numpy(), (0, 2, 1))
    rng = np.random.default_rng(0)
    f = synth.aa_filter_taps()
    if name == "amp":
        for shape in [(2, 24, 1000), (1, 768, 301), (2, 5, 64)]:
            x = (rng.standard_normal(shape) * 1.5).astype(np.float32)
            al = (rng.standard_normal(shape[1]) * 0.3).astype(np.float32)
            be = (rng.standard_normal(shape[1]) * 0.3).astype(np.float32)
            ref = O.activation1d(x.astype(np.float64), al.astype(np.float64), be.astype(np.float64), True, f.astype(np.float64), f.astype(np.float64))
            a = torch.from_numpy(np.exp(al)).to(dev)
            ib = torch.from_numpy(1.0 / (np.exp(be) + np.float32(1e-9))).to(dev)
            for fast in (0, 1):
                y = cf(ops.activation1d(cl(x), a, ib, f, f, fast_sin=fast))
                print(f"amp {shape} fast_sin={fast}: max err {np.abs(y - ref).max():.3e}")
        return
    if name == "tiny_gen":
        import __graft_entry__ as g
        g.smoke()
        return

    def conv_case(B, Ch, Ln, k, d, backend, split, mb=0):
        x = rng.standard_normal((B, Ch, Ln)).astype(np.float32)
        v = (rng.standard_normal((Ch, Ch, k)) / np.sqrt(Ch * k)).astype(np.float32)
        g = (np.sqrt((v.astype(np.float64) ** 2).sum(axis=(1, 2), keepdims=True)) * rng.uniform(0.7, 1.4, (Ch, 1, 1))).astype(np.float32)
        b = (rng.standard_normal(Ch) * 0.1).astype(np.float32)
        L.set_tuning("umma_mb", mb)
        pc = ops.pack_conv(*(torch.from_numpy(t).to(dev) for t in (v, g, b)), dilation=d, padding=O.get_padding(k, d), backend=backend, split=split)
        y = cf(ops.conv(cl(x), pc))
        torch.cuda.synchronize()
        w = O.weight_norm_fold(v.astype(np.float64), g.astype(np.float64))
        ref = O.conv1d(x.astype(np.float64), w, b.astype(np.float64), d, O.get_padding(k, d))
        refq = O.conv1d(bf16_round(x).astype(np.float64), bf16_round(w.astype(np.float32)).astype(np.float64), b.astype(np.float64), d, O.get_padding(k, d))
        print(f"{name} B{B} C{Ch} L{Ln} k{k} d{d}: err vs fp64 {np.abs(y - ref).max():.3e}, vs bf16-emulated {np.abs(y - refq).max():.3e}, |ref|max {np.abs(ref).max():.2f}", flush=True)

    shapes = [(2, 64, 300, 3, 1), (1, 128, 517, 7, 3), (2, 192, 260, 11, 5), (3, 24, 1000, 11, 5), (1, 768, 140, 3, 1)]
    if name == "simt":
        for s in shapes[:3]:
            conv_case(*s, backend=L.SIMT, split=False)
    elif name in ("umma_mb1", "umma_mb2", "umma_mb4"):
        for s in shapes:
            if int(name[-1]) * (256 if s[1] >= 256 else 128) <= 512:
                conv_case(*s, backend=L.UMMA, split=False, mb=int(name[-1]))
    elif name == "umma_split":
        for s in shapes:
            conv_case(*s, backend=L.UMMA, split=True)
    elif name == "umma_convT":
        for cin, cout, k, u, Ln in [(64, 32, 8, 4, 50), (48, 24, 4, 2, 200)]:
            x = rng.standard_normal((2, cin, Ln)).astype(np.float32)
            v = (rng.standard_normal((cin, cout, k)) / np.sqrt(cin * k / u)).astype(np.float32)
            g = np.sqrt((v.astype(np.float64) ** 2).sum(axis=(1, 2), keepdims=True)).astype(np.float32)
            b = (rng.standard_normal(cout) * 0.1).astype(np.float32)
            pc = ops.pack_conv(*(torch.from_numpy(t).to(dev) for t in (v, g, b)), transposed=True, stride=u, padding=(k - u) // 2, backend=L.UMMA, split=False)
            y = cf(ops.conv(cl(x), pc).reshape(2, Ln * u, cout))
            w = O.weight_norm_fold(v.astype(np.float64), g.astype(np.float64))
            refq = O.conv_transpose1d(bf16_round(x).astype(np.float64), bf16_round(w.astype(np.float32)).astype(np.float64), b.astype(np.float64), u, (k - u) // 2)
            print(f"convT {cin}->{cout} k{k} u{u}: err vs bf16-emulated {np.abs(y - refq).max():.3e}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        run_case(sys.argv[2])
        sys.exit(0)
    for c in sys.argv[1:] or CASES:
        print(f"=== {c}", flush=True)
        try:
            r = subprocess.run([sys.executable, __file__, "--one", c], capture_output=True, text=True, timeout=300)
            print(r.stdout[-3000:], flush=True)
            if r.returncode != 0:
                print(f"[{c}] exit {r.returncode}\n{r.stderr[-2500:]}", flush=True)
        except subprocess.TimeoutExpired:
            print(f"[{c}] TIMEOUT", flush=True)
