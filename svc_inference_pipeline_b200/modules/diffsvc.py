"""B200-native DiffSVC denoiser with the reference's call surface.

Drop-in for the denoiser of the reference's ``modules/diffsvc.py`` -- ``DiffSVC(cfg.mapper)`` (``:235-282``) and its
``forward(mel_spec[B, L, n_mel], conditioner[B, L, cond], diffusion_step[B, 1]) -> (noise[B, L, n_mel], stats)``
(``:284-321``), the function the sampler calls once per diffusion step, 1000 times per utterance
(``modules/diffsvcrepo_inference.py:61, :234-235``):

* same constructor fields (``n_mel``, ``residual_channels``, ``diffusion_fc_size``, ``conditioner_size``,
  ``dilation_cycle_length``, ``residual_kernel_size``, ``residual_layer_num``, ``noise_schedule_factors``);
* same ``state_dict`` keys and shapes (``mel_preprocess.projection``, ``diffusion_embedding.projection{1,2}``,
  ``residual_layers.{i}.{dilated_conv, diffusion_projection, conditioner_projection, output_projection}``,
  ``skip_projection``, ``output_projection``), so a mapper checkpoint's ``state_dict`` loads unchanged;
* the step-embedding table is a non-persistent buffer built with the reference's own expression (``:45-55``) on the
  host at construction, like the reference does.

One step is a pre-built *program* (``include/bvg_b200.h``) issued by a single C call and, with ``use_cuda_graph``, replayed
as a CUDA graph: the step index lives in a device buffer the graph reads.  Per residual layer:

    rowop ADDVEC   y = x + diffusion_projection(step)           -> (hi, lo) operand planes          (:213)
    conv  k=3, d   dilated_conv(y) + [conditioner_projection(e)] C -> 2C on tcgen05, fp32 out       (:220)
    rowop GATE     sigmoid(gate) * tanh(filter)                  -> operand planes                   (:225-227)
    conv  1x1      output_projection C -> 2C, in place on the [x | skip] buffer:                    (:229-232, :307)
                   x = (x + r) / sqrt(2) | skip = s + skip   (epilogue, per-channel divisor)

The conditioner does not change between the steps of one utterance, so its 20 projections ``[B, L, cond] -> [B, L, 2C]``
(``:216-219``; the reference recomputes them every step) run once per conditioner tensor and enter the dilated
convolution's epilogue as the running sum.  Dense layers are the 3-pass split product of the vocoder's fp32 path
(``precision="fp32"``, 1e-4 parity) or bf16 (``"bf16"``); ``"fp32_simt"`` is the exact-fp32 FFMA anchor.  There is no
CPU path.
"""
from __future__ import annotations

import ctypes as C
import weakref
from math import sqrt

import numpy as np
import torch
from torch import nn

from .. import _lib as L
from .bigvgan import _Buf, _NULL, _PackedConv, _Program

__all__ = ["DiffSVC", "StepEncoder", "DiffusionState"]

_MODES = {"fp32": (L.UMMA, L.SPLIT), "bf16": (L.UMMA, L.BF16), "fp32_simt": (L.SIMT, L.F32)}
# rows of the sampler's schedule table (the globals of reference modules/diffsvcrepo_inference.py:177-197)
SCHEDULE_ROWS = ("sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_mean_coef1", "posterior_mean_coef2",
                 "posterior_log_variance_clipped", "alphas_cumprod")


class _Conv(nn.Module):
    """Plain (not weight-normed) ``nn.Conv1d`` parameters, ``weight [Cout, Cin, K]`` + ``bias`` with the reference's
    ``kaiming_normal_`` init (``modules/diffsvc.py:23-26``)."""

    def __init__(self, cin, cout, ksize, dilation=1, padding=0):
        super().__init__()
        self.cin, self.cout, self.ksize, self.dilation, self.padding = cin, cout, ksize, dilation, padding
        self.stride, self.transposed = 1, False
        w = torch.empty(cout, cin, ksize)
        nn.init.kaiming_normal_(w)
        self.weight = nn.Parameter(w)
        bound = 1.0 / sqrt(cin * ksize)
        self.bias = nn.Parameter(torch.empty(cout).uniform_(-bound, bound))


class _View:
    """What ``_PackedConv`` reads from a layer (``weight_v`` / ``weight_g`` / ``bias`` + geometry): a slice of a ``_Conv``."""

    def __init__(self, conv: _Conv, rows=None, bias=True, zero_rows=0):
        w = conv.weight.detach()
        b = conv.bias.detach()
        if rows is not None:
            w, b = w[rows], b[rows]
        if zero_rows:  # extra output channels that are identically zero (the skip half of the [x | skip] buffer)
            w = torch.cat([w, w.new_zeros(zero_rows, *w.shape[1:])])
            b = torch.cat([b, b.new_zeros(zero_rows)])
        self.weight_v, self.weight_g = w.contiguous(), None
        self.bias = b.contiguous() if bias else torch.zeros_like(b)
        self.cin, self.cout, self.ksize = conv.cin, w.shape[0], conv.ksize
        self.dilation, self.stride, self.padding, self.transposed = conv.dilation, 1, conv.padding, False


class StepEncoder(nn.Module):
    """Reference ``modules/diffsvc.py:29-93``: sin / cos table lookup + two SiLU layers."""

    def __init__(self, max_steps: int, FC_size: int):
        super().__init__()
        self.max_steps = max_steps
        self.register_buffer("embedding", self.build_embedding(max_steps), persistent=False)
        self.projection1 = nn.Linear(128, FC_size)
        self.projection2 = nn.Linear(FC_size, FC_size)

    @staticmethod
    def build_embedding(max_steps: int) -> torch.Tensor:
        # the reference's expression verbatim (:51-54): the table's large arguments (up to 1e7 rad) make it sensitive to
        # the last bit of 10 ** x, so it is evaluated by the same library on the host, at construction, like there
        steps = torch.arange(max_steps).unsqueeze(1)
        dims = torch.arange(64).unsqueeze(0)
        table = steps * 10.0 ** (dims * 4.0 / 63.0)
        return torch.cat([torch.sin(table), torch.cos(table)], dim=1)


class _Preprocessor(nn.Module):
    def __init__(self, n_mel, channels):
        super().__init__()
        self.projection = _Conv(n_mel, channels, 1)


class _ResidualBlock(nn.Module):
    def __init__(self, conditioner_size, fc_size, channels, dilation, kernel_size=3):
        super().__init__()
        if dilation == 1:
            if (kernel_size - 1) % 2 != 0:
                raise ValueError("Wrong kernel size for Conv1d")
            pad = (kernel_size - 1) // 2
        else:
            assert kernel_size == 3
            pad = dilation
        self.dilated_conv = _Conv(channels, 2 * channels, kernel_size, dilation=dilation, padding=pad)
        self.diffusion_projection = nn.Linear(fc_size, channels)
        self.conditioner_projection = _Conv(conditioner_size, 2 * channels, 1)
        self.output_projection = _Conv(channels, 2 * channels, 1)


class DiffSVC(nn.Module):
    """Denoiser of the DiffSVC acoustic model (reference ``modules/diffsvc.py:235-321``) running on libbvg_b200."""

    def __init__(self, cfg, precision: str = "fp32"):
        super().__init__()
        if precision not in _MODES:
            raise ValueError(f"precision must be one of {sorted(_MODES)}")
        self.cfg = cfg
        self.precision = precision
        f = cfg.noise_schedule_factors
        self.noise_schedule = np.linspace(f[0], f[1], int(f[2])).tolist()
        self.n_mel, self.channels = int(cfg.n_mel), int(cfg.residual_channels)
        self.cond_size, self.fc = int(cfg.conditioner_size), int(cfg.diffusion_fc_size)
        self.mel_preprocess = _Preprocessor(self.n_mel, self.channels)
        self.diffusion_embedding = StepEncoder(len(self.noise_schedule), self.fc)
        self.residual_layers = nn.ModuleList(
            [_ResidualBlock(self.cond_size, self.fc, self.channels, 2 ** (i % cfg.dilation_cycle_length), kernel_size=cfg.residual_kernel_size)
             for i in range(cfg.residual_layer_num)]
        )
        self.skip_projection = _Conv(self.channels, self.channels, 1)
        self.output_projection = _Conv(self.channels, self.n_mel, 1)
        nn.init.zeros_(self.output_projection.weight)
        if self.channels % 8 or self.cond_size % 8:
            raise ValueError("residual_channels and conditioner_size must be multiples of 8 for the tensor-core path")
        self.use_cuda_graph = True
        # programmatic dependent launch between the ~106 short kernels of a step (bvg_program_set_pdl)
        self.pdl = True
        self._packed = {}   # tile cap -> packed weights
        self._programs = {}
        self._cond_key = None
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._invalidate())

    # -- plumbing ---------------------------------------------------------------------------------
    def _invalidate(self):
        self._packed = {}
        self._programs = {}
        self._cond_key = None

    def _apply(self, fn, *args, **kwargs):
        self._invalidate()
        return super()._apply(fn, *args, **kwargs)

    def set_pdl(self, on: bool):
        self.pdl = bool(on)
        self._programs = {}
        self._cond_key = None
        return self

    def set_precision(self, precision: str):
        if precision not in _MODES:
            raise ValueError(f"precision must be one of {sorted(_MODES)}")
        if precision != self.precision:
            self.precision = precision
            self._invalidate()
        return self

    def _device(self):
        return self.skip_projection.weight.device

    def _require_cuda(self):
        dev = self._device()
        if dev.type != "cuda":
            raise RuntimeError("svc_inference_pipeline_b200.DiffSVC has no CPU path: move the model to a B200 (model.cuda())")
        L.require_sm100(dev.index if dev.index is not None else torch.cuda.current_device())
        return dev

    @staticmethod
    def tile_cap(rows: int) -> int:
        """Widest N tile of the dense layers for ``rows = B * L`` activation rows.  A step is ~86 dependent launches of a few
        microseconds; with one utterance (379 rows = 3 row blocks) 128-column tiles occupy 12-24 of the 148 SMs and every
        launch is one long serial chain of MMAs.  Narrow tiles spread a launch over more SMs; wide tiles re-read the
        activations less often once the rows alone fill the GPU.  Measured (tools/profile_diffsvc.py, ms per step, fp32 /
        bf16 path): 379 rows 0.54 / 0.47 at 32 columns, 0.68 / 0.49 at 64, 0.76 / 0.62 at 128; 1516 rows 0.79 / 0.70, 0.77 /
        0.57, 0.83 / 0.70; 15 008 rows 4.6 / 3.1, 4.2 / 2.7, 3.1 / 2.6."""
        return 32 if rows <= 512 else 64 if rows <= 4096 else 128

    def _pack(self, cap: int):
        dev = self._require_cuda()
        backend, op_dt = _MODES[self.precision]
        split = op_dt == L.SPLIT
        stream = torch.cuda.current_stream(dev).cuda_stream
        Cc = self.channels
        pk = {"tune": L.new_tuning(umma_ntile_cap=cap)}
        tune = C.pointer(pk["tune"])

        def pack(name, view):
            pk[name] = _PackedConv(view, backend, split, stream, tune=tune)

        with torch.cuda.device(dev):
            # x and the running skip sum share one [rows, 2C] buffer (see _build): the preprocessor writes [relu(.) | 0]
            pack("pre", _View(self.mel_preprocess.projection, zero_rows=Cc))
            for i, rl in enumerate(self.residual_layers):
                # the conditioner projection's bias rides along with the cached projection; the dilated conv keeps its own
                pack(f"dil{i}", _View(rl.dilated_conv))
                pack(f"cond{i}", _View(rl.conditioner_projection))
                pack(f"outp{i}", _View(rl.output_projection))  # both chunks (:230): [residual | skip]
            pack("skipproj", _View(self.skip_projection))
            pack("out", _View(self.output_projection))
            f32 = lambda t: t.detach().float().contiguous()
            pk["w1"], pk["b1"] = f32(self.diffusion_embedding.projection1.weight), f32(self.diffusion_embedding.projection1.bias)
            pk["w2"], pk["b2"] = f32(self.diffusion_embedding.projection2.weight), f32(self.diffusion_embedding.projection2.bias)
            pk["wd"] = torch.stack([f32(rl.diffusion_projection.weight) for rl in self.residual_layers]).contiguous()
            pk["bd"] = torch.stack([f32(rl.diffusion_projection.bias) for rl in self.residual_layers]).contiguous()
            pk["table"] = f32(self.diffusion_embedding.embedding)
            # per-channel divisor of the output projection's epilogue: (x + residual) / sqrt(2) | skip sum (:232, :307)
            pk["coldiv"] = torch.cat([torch.full((Cc,), sqrt(2.0)), torch.ones(Cc)]).to(device=dev, dtype=torch.float32)
            torch.cuda.current_stream(dev).synchronize()
        self._packed[cap] = pk

    # -- programs -----------------------------------------------------------------------------------
    def _build(self, B: int, Ln: int, float_steps: bool = False):
        rows = B * Ln
        cap = self.tile_cap(rows)
        if cap not in self._packed:
            self._pack(cap)
        dev = self._device()
        backend, op_dt = _MODES[self.precision]
        pk, Cc, nl = self._packed[cap], self.channels, len(self.residual_layers)
        tune = C.pointer(pk["tune"])
        mel_pitch = pk["pre"].x_pitch
        keep = []

        def f32buf(n):
            t = torch.empty(n, dtype=torch.float32, device=dev)
            keep.append(t)
            return t

        def f32t(t):
            return L.Tensor(t.data_ptr(), None, L.F32, 0)

        mel_in = torch.empty(B, Ln, self.n_mel, dtype=torch.float32, device=dev)
        cond_in = torch.empty(B, Ln, self.cond_size, dtype=torch.float32, device=dev)
        step_in = torch.zeros(B, dtype=torch.float32 if float_steps else torch.int32, device=dev)
        out = torch.empty(B, Ln, self.n_mel, dtype=torch.float32, device=dev)
        mel_op = _Buf(op_dt, rows * mel_pitch, dev)
        cond_op = _Buf(op_dt, rows * self.cond_size, dev)
        y_op = _Buf(op_dt, rows * Cc, dev)
        z_op = _Buf(op_dt, rows * Cc, dev)
        xs, y2 = f32buf(rows * 2 * Cc), f32buf(rows * 2 * Cc)   # xs = [x | skip] per row
        condproj = [f32buf(rows * 2 * Cc) for _ in range(nl)]
        dproj = f32buf(nl * B * Cc)
        keep += [mel_in, cond_in, step_in, out, mel_op, cond_op, y_op, z_op]

        def rowop(ops, kind, x, out_t, C_, x_pitch, out_pitch, vec=None, div=1.0):
            op = L.Op()
            op.kind = L.OP_ROWOP
            d = op.u.rowop
            d.kind, d.d_x, d.d_vec, d.out, d.div = kind, x, vec, out_t, float(div)
            d.B, d.L, d.C, d.x_pitch, d.out_pitch = B, Ln, C_, x_pitch, out_pitch
            ops.append(op)

        def conv(ops, name, x, out_t, res=None, acc=None, div=1.0, relu=False, coldiv=None):
            op = L.Op()
            op.kind = L.OP_CONV
            d = op.u.conv
            d.x, d.out = x.tensor() if isinstance(x, _Buf) else x, out_t
            d.res = res if res is not None else _NULL
            d.acc_in = acc if acc is not None else _NULL
            d.div, d.B, d.L = float(div), B, Ln
            d.w = C.pointer(pk[name].desc)
            d.relu = int(relu)
            d.d_coldiv = coldiv.data_ptr() if coldiv is not None else None
            d.tune = tune
            ops.append(op)

        def operand(buf):  # the SIMT anchor reads fp32 operands
            return buf.tensor()

        # conditioner program: runs when the conditioner tensor changes (once per utterance in the sampler)
        cops = []
        rowop(cops, L.ROW_ADDVEC, cond_in.data_ptr(), cond_op.tensor(), self.cond_size, self.cond_size, self.cond_size)
        for i in range(nl):
            conv(cops, f"cond{i}", operand(cond_op), f32t(condproj[i]))
        # step program
        ops = []
        op = L.Op()
        op.kind = L.OP_DIFFEMBED
        d = op.u.diffembed
        if float_steps:
            d.d_step_f = step_in.data_ptr()
        else:
            d.d_step = step_in.data_ptr()
        d.d_table = pk["table"].data_ptr()
        d.d_w1, d.d_b1, d.d_w2, d.d_b2 = pk["w1"].data_ptr(), pk["b1"].data_ptr(), pk["w2"].data_ptr(), pk["b2"].data_ptr()
        d.d_wd, d.d_bd, d.d_out = pk["wd"].data_ptr(), pk["bd"].data_ptr(), dproj.data_ptr()
        d.B, d.emb, d.fc, d.C, d.n_layers, d.max_steps = B, pk["table"].shape[1], self.fc, Cc, nl, pk["table"].shape[0]
        ops.append(op)
        rowop(ops, L.ROW_ADDVEC, mel_in.data_ptr(), mel_op.tensor(), self.n_mel, self.n_mel, mel_pitch)
        conv(ops, "pre", operand(mel_op), f32t(xs), relu=True)                                     # mel_preprocess (:118-128) -> [x | 0]
        for i in range(nl):
            rowop(ops, L.ROW_ADDVEC, xs.data_ptr(), y_op.tensor(), Cc, 2 * Cc, Cc, vec=dproj.data_ptr() + 4 * i * B * Cc)
            conv(ops, f"dil{i}", operand(y_op), f32t(y2), acc=f32t(condproj[i]))                    # dilated_conv(y) + conditioner (:220)
            rowop(ops, L.ROW_GATE, y2.data_ptr(), z_op.tensor(), Cc, 2 * Cc, Cc)
            # output_projection, both chunks in one launch, in place: x = (x + residual) / sqrt(2) | skip = s + skip (:229-232, :307)
            conv(ops, f"outp{i}", operand(z_op), f32t(xs), res=f32t(xs), coldiv=pk["coldiv"])
        rowop(ops, L.ROW_SCALE, xs.data_ptr() + 4 * Cc, y_op.tensor(), Cc, 2 * Cc, Cc, div=sqrt(nl))  # skip / sqrt(n) (:313)
        conv(ops, "skipproj", operand(y_op), z_op.tensor(), relu=True)                             # skip_projection + relu (:315-316)
        conv(ops, "out", operand(z_op), f32t(out))                                                  # output_projection (:317)
        prog = _Program(ops, keep, mel_in, out, len(ops))
        prog.cond_prog = _Program(cops, keep, cond_in, None, len(cops))
        prog.cond_in, prog.step_in = cond_in, step_in
        if not float_steps:
            # sampler (modules/diffsvcrepo_inference.py): the step followed by the DDPM update of the sample, which
            # lives in mel_in -- one program, one graph replay per diffusion step
            n_steps = pk["table"].shape[0]
            prog.tables = torch.zeros(6, n_steps, dtype=torch.float32, device=dev)   # rows: SCHEDULE_ROWS
            prog.noise = torch.zeros(B, 1, self.n_mel, Ln, dtype=torch.float32, device=dev)
            sop = L.Op()
            sop.kind = L.OP_SAMPLE
            sd = sop.u.sample
            sd.mode, sd.clip = L.SAMPLE_DDPM, 1
            sd.d_x, sd.d_x_out, sd.d_eps, sd.d_noise, sd.d_step = mel_in.data_ptr(), mel_in.data_ptr(), out.data_ptr(), prog.noise.data_ptr(), step_in.data_ptr()
            row = lambda k: prog.tables.data_ptr() + 4 * k * n_steps
            sd.d_sqrt_recip, sd.d_sqrt_recipm1, sd.d_coef1, sd.d_coef2, sd.d_logvar, sd.d_alphas_cumprod = (row(k) for k in range(6))
            sd.B, sd.L, sd.n_mel, sd.n_steps = B, Ln, self.n_mel, n_steps
            prog.ddpm_prog = _Program(ops + [sop], keep, mel_in, out, len(ops) + 1)
            prog.ddpm_prog.set_pdl(self.pdl)
        prog.set_pdl(self.pdl)
        prog.cond_prog.set_pdl(self.pdl)
        return prog

    def _program(self, B, Ln, float_steps=False):
        key = (B, Ln, self.precision, bool(float_steps))
        prog = self._programs.get(key)
        if prog is None:
            with torch.cuda.device(self._device()):
                prog = self._build(B, Ln, float_steps)
            if len(self._programs) >= 4:
                self._programs.clear()
                self._cond_key = None
            self._programs[key] = prog
        return prog

    def launches_per_step(self, B, Ln):
        return self._program(B, Ln).launches

    # -- forward ------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, mel_spec: torch.Tensor, conditioner: torch.Tensor, diffusion_step):
        """Reference ``modules/diffsvc.py:284-321``.  ``mel_spec [B, L, n_mel]``, ``conditioner [B, L, cond]``,
        ``diffusion_step [B, 1]`` (or ``[B]``): integer steps (the sampler's ``t``) or fractional ones (interpolated embedding).  Returns ``(noise [B, L, n_mel], stats)``; ``stats`` (the
        reference's dictionary of intermediate tensors, which its sampler never reads) is empty."""
        dev = self._require_cuda()
        if mel_spec.dim() != 3 or mel_spec.shape[2] != self.n_mel:
            raise ValueError(f"expected mel_spec of shape [B, L, {self.n_mel}], got {tuple(mel_spec.shape)}")
        B, Ln, _ = mel_spec.shape
        if tuple(conditioner.shape) != (B, Ln, self.cond_size):
            raise ValueError(f"expected conditioner of shape [{B}, {Ln}, {self.cond_size}], got {tuple(conditioner.shape)}")
        step = torch.as_tensor(diffusion_step)
        float_steps = step.dtype not in (torch.int32, torch.int64)  # StepEncoder.forward :79-85: table lookup or lerp_embedding
        step = step.reshape(-1)
        if step.numel() == 1 and B > 1:
            step = step.expand(B)
        prog = self._program(B, Ln, float_steps)
        with torch.cuda.device(dev):
            self._set_conditioner(prog, conditioner, dev)
            prog.mel_in.copy_(mel_spec, non_blocking=True)
            prog.step_in.copy_(step.to(prog.step_in.dtype), non_blocking=True)
            self._launch(prog, dev)
            return prog.out.clone(), {}

    def _launch(self, prog, dev):
        """Issue a program on the current stream: a CUDA-graph replay (captured on first use) or the launch list."""
        stream = torch.cuda.current_stream(dev)
        if not self.use_cuda_graph:
            prog.run(stream.cuda_stream)
            return
        if prog.graph is None:
            keep = [t.clone() for t in (prog.mel_in, prog.out)]  # the warm-up and the capture must not move the sampler's state
            prog.run(stream.cuda_stream)  # warm-up outside capture (lazy function attributes)
            stream.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                prog.run(torch.cuda.current_stream(dev).cuda_stream)
            prog.graph = g
            prog.mel_in.copy_(keep[0])
            prog.out.copy_(keep[1])
        prog.graph.replay()

    def _set_conditioner(self, prog, conditioner, dev):
        stream = torch.cuda.current_stream(dev)
        # same tensor OBJECT (weak reference: a freed tensor's address may be reused), unmodified since, same program
        ck = self._cond_key
        fresh = ck is None or ck[0]() is not conditioner or ck[1] != conditioner._version or ck[2] is not prog
        if fresh:  # new conditioner: its 20 projections once, reused by every step that follows
            prog.cond_in.copy_(conditioner, non_blocking=True)
            prog.cond_prog.run(stream.cuda_stream)
            self._cond_key = (weakref.ref(conditioner), conditioner._version, prog)

    def sampler(self, conditioner: torch.Tensor, schedule_tables) -> "DiffusionState":
        """State of one sampling run over ``conditioner [B, L, cond]`` (used by ``modules/diffsvcrepo_inference.py``)."""
        dev = self._require_cuda()
        B, Ln, cs = conditioner.shape
        if cs != self.cond_size:
            raise ValueError(f"expected conditioner of shape [B, L, {self.cond_size}], got {tuple(conditioner.shape)}")
        prog = self._program(B, Ln, False)
        tables = torch.as_tensor(schedule_tables, dtype=torch.float32)
        if tables.dim() != 2 or tables.shape[0] != 6 or tables.shape[1] > prog.tables.shape[1]:
            raise ValueError(f"schedule tables must be [6, n <= {prog.tables.shape[1]}] (the denoiser's step embedding has {prog.tables.shape[1]} rows), got {tuple(tables.shape)}")
        with torch.cuda.device(dev):
            self._set_conditioner(prog, conditioner, dev)
            prog.tables.zero_()
            prog.tables[:, : tables.shape[1]].copy_(tables)
        return DiffusionState(self, prog, dev, tables.shape[1])


class DiffusionState:
    """The sample ``x [B, L, n_mel]`` of one run of the sampler and the launches that move it (reference
    ``modules/diffsvcrepo_inference.py``).  ``x`` is the denoiser program's input buffer and ``eps`` its output buffer, so
    a diffusion step is the program (a CUDA-graph replay) plus one ``bvg_sample_fwd`` update and nothing is copied."""

    def __init__(self, model: DiffSVC, prog, dev, n_steps: int):
        self.model, self.prog, self.dev, self.n_steps = model, prog, dev, n_steps
        self.x, self.eps, self.noise = prog.mel_in, prog.out, prog.noise
        self._hist = []      # PLMS: earlier predictions, most recent first (the reference's noise_list, deque(maxlen=4))
        self._spare = [torch.empty_like(prog.out) for _ in range(4)]
        self._x_keep = None

    def _check(self, step: int):
        if not 0 <= int(step) < self.n_steps:
            raise IndexError(f"diffusion step {step} outside the schedule of {self.n_steps} steps")

    def denoise(self, step: int):
        """``eps = denoise_fn(x, cond, step)`` (every batch item at the same step, like the sampler's ``torch.full``)."""
        self._check(step)
        with torch.cuda.device(self.dev):
            self.prog.step_in.fill_(int(step))
            self.model._launch(self.prog, self.dev)

    def ddpm_step(self, step: int):
        """``p_sample`` (``:88-97``) with the noise in ``self.noise`` ([B, 1, n_mel, L], the reference's ``randn(x.shape)``)."""
        self._check(step)
        with torch.cuda.device(self.dev):
            self.prog.step_in.fill_(int(step))
            self.model._launch(self.prog.ddpm_prog, self.dev)

    def _update(self, step, interval, combine, x_src, hist, save):
        d = L.SampleDesc()
        d.mode, d.clip = L.SAMPLE_PLMS, 0
        d.d_x, d.d_x_out, d.d_eps, d.d_step = x_src.data_ptr(), self.x.data_ptr(), self.eps.data_ptr(), self.prog.step_in.data_ptr()
        d.d_alphas_cumprod = self.prog.tables.data_ptr() + 4 * 5 * self.prog.tables.shape[1]
        for k, h in enumerate(hist):
            d.d_hist[k] = h.data_ptr()
        d.d_eps_save = save.data_ptr() if save is not None else None
        d.interval, d.combine = int(interval), int(combine)
        d.B, d.L, d.n_mel = self.x.shape
        d.n_steps = self.prog.tables.shape[1]
        with torch.cuda.device(self.dev):
            self.prog.step_in.fill_(int(step))
            L.check(L.lib().bvg_sample_fwd(C.byref(d), torch.cuda.current_stream(self.dev).cuda_stream), "sample")

    def plms_step(self, step: int, interval: int):
        """``p_sample_plms`` (``:100-150``): pseudo linear multistep update with the last three predictions."""
        self._check(step)
        self.denoise(step)
        save = self._spare.pop()
        if not self._hist:
            # first step: predictor with eps, second evaluation at the predicted point, average (:126-139)
            if self._x_keep is None:
                self._x_keep = torch.empty_like(self.x)
            self._x_keep.copy_(self.x)
            self._update(step, interval, 0, self.x, [], save)            # x <- get_x_pred(x, eps, t); save <- eps
            self.denoise(max(int(step) - int(interval), 0))
            self._update(step, interval, 1, self._x_keep, [save], None)  # x <- get_x_pred(x_keep, (save + eps) / 2, t)
        else:
            self._update(step, interval, min(len(self._hist), 3) + 1, self.x, self._hist[:3], save)
        self._hist.insert(0, save)
        if len(self._hist) > 3:
            self._spare.append(self._hist.pop())
