"""Single-op entry points over the C ABI, on channels-last torch tensors.

These wrap one ``bvg_*_fwd`` call each (``include/bvg_b200.h``) for op-level parity tests, profiling
and debugging; the generator itself runs pre-built programs (``modules/bigvgan.py``).  Inputs and
outputs are float32 torch tensors ``[B, L, C]`` on the GPU; ``*_dtype`` selects the element format the
kernel actually reads / writes (F32, BF16, SPLIT), with conversion done by ``bvg_convert``.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L
from .modules.bigvgan import _NULL, _Buf, _PackedConv, _WNConv


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def to_buf(x: torch.Tensor, dtype: int) -> _Buf:
    """float32 tensor -> buffer in the given element format (device-side conversion)."""
    x = x.contiguous().float()
    buf = _Buf(dtype, x.numel(), x.device)
    src = L.Tensor(x.data_ptr(), None, L.F32, 0)
    dst = buf.tensor()
    with torch.cuda.device(x.device):
        L.check(L.lib().bvg_convert(C.byref(src), C.byref(dst), x.numel(), _stream(x.device)), "convert")
    buf._src = x
    return buf


def from_buf(buf: _Buf, shape) -> torch.Tensor:
    out = torch.empty(shape, dtype=torch.float32, device=buf.hi.device)
    src = buf.tensor()
    dst = L.Tensor(out.data_ptr(), None, L.F32, 0)
    with torch.cuda.device(out.device):
        L.check(L.lib().bvg_convert(C.byref(src), C.byref(dst), out.numel(), _stream(out.device)), "convert")
    return out


def activation1d(x, a, invb, taps_up, taps_down, in_dtype=L.F32, out_dtype=L.F32, fast_sin=False):
    """Fused Activation1d on ``x [B, L, C]``; ``a`` / ``invb`` are the per-channel snake parameters
    after exponentiation (``a = exp(alpha)``, ``invb = 1 / (exp(beta) + 1e-9)``)."""
    B, Ln, Ch = x.shape
    xb = to_buf(x, in_dtype)
    yb = _Buf(out_dtype, x.numel(), x.device)
    a = a.contiguous().float()
    invb = invb.contiguous().float()
    d = L.AmpDesc()
    d.x, d.y = xb.tensor(), yb.tensor()
    d.d_a, d.d_invb = a.data_ptr(), invb.data_ptr()
    d.taps_up = (C.c_float * 12)(*[float(t) for t in taps_up])
    d.taps_down = (C.c_float * 12)(*[float(t) for t in taps_down])
    d.B, d.L, d.C, d.fast_sin = B, Ln, Ch, int(fast_sin)
    tune = L.tuning_ptr()
    if tune is not None:
        d.tune = tune
    with torch.cuda.device(x.device):
        L.check(L.lib().bvg_amp_fwd(C.byref(d), _stream(x.device)), "amp_fwd")
    return from_buf(yb, x.shape)


def pack_conv(v, g, bias, *, transposed=False, dilation=1, stride=1, padding=0, backend=L.UMMA, split=False, fold=1) -> _PackedConv:
    """Fold + pack a weight-normed conv given reference-layout ``weight_v`` / ``weight_g`` / ``bias``."""
    if transposed:
        cin, cout, k = v.shape
    else:
        cout, cin, k = v.shape
    holder = _WNConv(cin, cout, k, dilation=dilation, stride=stride, padding=padding, transposed=transposed)
    with torch.no_grad():
        holder.weight_v.copy_(v)
        holder.weight_g.copy_(g.reshape(holder.weight_g.shape))
        holder.bias.copy_(bias)
    holder = holder.to(v.device)
    with torch.cuda.device(v.device):
        pc = _PackedConv(holder, backend, split, _stream(v.device), fold=fold)
    pc._holder = holder
    return pc


def conv(x, pc: _PackedConv, *, x_dtype=None, out_dtype=L.F32, res=None, res_dtype=L.F32, acc=None, acc_dtype=L.F32, div=1.0, pre_amp=None, relu=False):
    """Tap-GEMM convolution of ``x [B, L, x_pitch]`` with packed weights; returns ``[B, L, n_total]``
    (for a transposed conv reshape to ``[B, L*u, Cout]``).  ``pre_amp = (a, invb, taps_up, taps_down, fast_sin)``
    fuses that Activation1d in front: ``x`` is then its fp32 input (``bvg_conv_desc.pre_amp``)."""
    B, Ln, Cx = x.shape
    assert Cx == pc.x_pitch, f"x has {Cx} channels per row, packed weights expect pitch {pc.x_pitch}"
    if pre_amp is not None:
        x_dtype = L.F32
    if x_dtype is None:
        x_dtype = L.F32 if pc.desc.backend == L.SIMT else (L.SPLIT if pc.desc.split else L.BF16)
    xb = to_buf(x, x_dtype)
    n = pc.n_total
    ob = _Buf(out_dtype, B * Ln * n, x.device)
    d = L.ConvDesc()
    d.x, d.out = xb.tensor(), ob.tensor()
    rb = to_buf(res, res_dtype) if res is not None else None
    ab = to_buf(acc, acc_dtype) if acc is not None else None
    d.res = rb.tensor() if rb is not None else _NULL
    d.acc_in = ab.tensor() if ab is not None else _NULL
    d.div, d.B, d.L = float(div), B, Ln
    d.w = C.pointer(pc.desc)
    d.relu = int(relu)
    tune = L.tuning_ptr()
    if tune is not None:
        d.tune = tune
    if pre_amp is not None:
        a, invb, taps_up, taps_down, fast_sin = pre_amp
        a = a.contiguous().float()
        invb = invb.contiguous().float()
        ad = L.AmpDesc()
        ad.x = xb.tensor()
        ad.d_a, ad.d_invb = a.data_ptr(), invb.data_ptr()
        ad.taps_up = (C.c_float * 12)(*[float(t) for t in taps_up])
        ad.taps_down = (C.c_float * 12)(*[float(t) for t in taps_down])
        ad.B, ad.L, ad.C, ad.fast_sin = B, Ln, Cx, int(fast_sin)
        if tune is not None:
            ad.tune = tune
        d.pre_amp = C.cast(C.pointer(ad), C.c_void_p)
    with torch.cuda.device(x.device):
        L.check(L.lib().bvg_conv_fwd(C.byref(d), _stream(x.device)), "conv_fwd")
    return from_buf(ob, (B, Ln, n))


def post(x, v, g, bias, in_dtype=L.F32):
    """conv_post + tanh: ``x [B, L, C]`` -> ``[B, L]``; ``v [1, C, K]``, ``g [1,1,1]``, ``bias [1]``."""
    B, Ln, Ch = x.shape
    k = v.shape[-1]
    dev = x.device
    xb = to_buf(x, in_dtype)
    w = torch.empty(Ch * k, dtype=torch.float32, device=dev)
    scratch = torch.empty(1, dtype=torch.float32, device=dev)
    v = v.contiguous().float()
    g = g.contiguous().float()
    out = torch.empty(B, Ln, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        L.check(L.lib().bvg_pack_post_weights(v.data_ptr(), g.data_ptr(), Ch, k, w.data_ptr(), scratch.data_ptr(), _stream(dev)), "pack_post")
        d = L.PostDesc()
        d.x = xb.tensor()
        d.d_w = w.data_ptr()
        d.bias = float(bias.reshape(-1)[0])
        d.d_out = out.data_ptr()
        d.B, d.L, d.C, d.ksize = B, Ln, Ch, k
        L.check(L.lib().bvg_post_fwd(C.byref(d), _stream(dev)), "post_fwd")
    return out


def pack_mel(mel, c_pad, dtype=L.F32):
    """``mel [B, C, T]`` -> channels-last ``[B, T, c_pad]`` (zero-padded channels), as float32."""
    B, Ch, T = mel.shape
    mel = mel.contiguous().float()
    ob = _Buf(dtype, B * T * c_pad, mel.device)
    d = L.PackDesc()
    d.d_mel = mel.data_ptr()
    d.out = ob.tensor()
    d.B, d.C, d.T, d.c_pad = B, Ch, T, c_pad
    with torch.cuda.device(mel.device):
        L.check(L.lib().bvg_pack_mel(C.byref(d), _stream(mel.device)), "pack_mel")
    return from_buf(ob, (B, T, c_pad))
