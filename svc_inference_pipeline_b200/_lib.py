"""ctypes binding of ``libbvg_b200.so`` (C ABI: ``include/bvg_b200.h``).

The library is the product's only compute path: if it is missing or the device is not sm_100,
loading / calling fails loudly -- there is no CPU or PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# BVG_B200_LIB: A/B-test another build of the same library (development aid; never a fallback)
LIB_PATH = os.environ.get("BVG_B200_LIB") or os.path.join(_HERE, "libbvg_b200.so")

F32, BF16, SPLIT = 0, 1, 2
SIMT, UMMA = 0, 1
OP_PACK, OP_AMP, OP_CONV, OP_POST, OP_ROWOP, OP_DIFFEMBED, OP_SAMPLE = 0, 1, 2, 3, 4, 5, 6
N_OP_KINDS = 7
ROW_ADDVEC, ROW_GATE, ROW_SCALE = 0, 1, 2
SAMPLE_DDPM, SAMPLE_PLMS = 0, 1
MAX_TAPS, MAX_NTILES = 16, 32
ABI_VERSION = 5


class BvgError(RuntimeError):
    pass


class Tuning(C.Structure):
    """``bvg_tuning``: test / A-B knobs attached to a descriptor (``tune`` field; NULL = the library's defaults)."""

    _fields_ = [(n, C.c_int32) for n in (
        "amp_vec", "amp_chunk", "amp_mma", "amp_mma_tiles", "amp_packed", "amp_stream", "amp_stream_bf16", "amp_ct",
        "umma_mb", "umma_wide_mb2", "umma_max_ctas", "umma_ntile_cap", "umma_tap_group", "umma_a_stages", "umma_stack", "umma_pair", "umma_pair_smem_kb", "umma_pair_min", "_r1", "_r2")]


class Tensor(C.Structure):
    _fields_ = [("d_ptr", C.c_void_p), ("d_lo", C.c_void_p), ("dtype", C.c_int32), ("_pad", C.c_int32)]


class AmpDesc(C.Structure):
    _fields_ = [
        ("x", Tensor),
        ("y", Tensor),
        ("d_a", C.c_void_p),
        ("d_invb", C.c_void_p),
        ("taps_up", C.c_float * 12),
        ("taps_down", C.c_float * 12),
        ("B", C.c_int32),
        ("L", C.c_int32),
        ("C", C.c_int32),
        ("fast_sin", C.c_int32),
        ("tune", C.POINTER(Tuning)),
    ]


class ConvWeights(C.Structure):
    _fields_ = [
        ("backend", C.c_int32),
        ("cin", C.c_int32),
        ("n_total", C.c_int32),
        ("n_tile", C.c_int32),
        ("n_tiles", C.c_int32),
        ("cin_pad", C.c_int32),
        ("x_pitch", C.c_int32),
        ("tap_stride", C.c_int32),
        ("split", C.c_int32),
        ("n_taps", C.c_int32 * MAX_NTILES),
        ("shift", (C.c_int32 * MAX_TAPS) * MAX_NTILES),
        ("d_w", C.c_void_p),
        ("d_w_lo", C.c_void_p),
        ("d_bias", C.c_void_p),
    ]


class ConvDesc(C.Structure):
    _fields_ = [
        ("x", Tensor),
        ("out", Tensor),
        ("res", Tensor),
        ("acc_in", Tensor),
        ("div", C.c_float),
        ("B", C.c_int32),
        ("L", C.c_int32),
        ("w", C.POINTER(ConvWeights)),
        ("pre_amp", C.c_void_p),  # const bvg_amp_desc*: Activation1d fused in front of the convolution (or NULL)
        ("tune", C.POINTER(Tuning)),
        ("relu", C.c_int32),
        ("_pad", C.c_int32),
        ("d_coldiv", C.c_void_p),  # const float* [n_total]: per-channel divisor (or NULL)
    ]


class ConvGeom(C.Structure):
    _fields_ = [
        ("transposed", C.c_int32),
        ("cin", C.c_int32),
        ("cout", C.c_int32),
        ("ksize", C.c_int32),
        ("dilation", C.c_int32),
        ("stride", C.c_int32),
        ("padding", C.c_int32),
        ("backend", C.c_int32),
        ("split", C.c_int32),
        ("n_tile", C.c_int32),
        ("fold", C.c_int32),
        ("_pad", C.c_int32),
        ("tune", C.POINTER(Tuning)),
    ]


class PostDesc(C.Structure):
    _fields_ = [
        ("x", Tensor),
        ("d_w", C.c_void_p),
        ("bias", C.c_float),
        ("d_out", C.c_void_p),
        ("B", C.c_int32),
        ("L", C.c_int32),
        ("C", C.c_int32),
        ("ksize", C.c_int32),
    ]


class PackDesc(C.Structure):
    _fields_ = [
        ("d_mel", C.c_void_p),
        ("out", Tensor),
        ("B", C.c_int32),
        ("C", C.c_int32),
        ("T", C.c_int32),
        ("c_pad", C.c_int32),
        ("d_range", C.c_void_p),
        ("d_min", C.c_void_p),
    ]


class TailDesc(C.Structure):
    _fields_ = [
        ("d_wave", C.c_void_p),
        ("d_pcm", C.c_void_p),
        ("d_peak", C.c_void_p),
        ("B", C.c_int32),
        ("L", C.c_int32),
        ("fade_len", C.c_int32),
        ("silence", C.c_int32),
        ("volume_peak", C.c_float),
        ("_pad", C.c_int32),
    ]


class StitchDesc(C.Structure):
    _fields_ = [
        ("d_wave", C.c_void_p),
        ("wave_stride", C.c_int64),
        ("d_out", C.c_void_p),
        ("out_len", C.c_int64),
        ("d_table", C.c_void_p),
        ("h_table", C.c_void_p),
        ("n_chunks", C.c_int32),
        ("_pad", C.c_int32),
    ]


class LogmelDesc(C.Structure):
    _fields_ = [
        ("d_wave", C.c_void_p),
        ("wave_stride", C.c_int64),
        ("d_out", C.c_void_p),
        ("d_basis", C.c_void_p),
        ("d_band", C.c_void_p),
        ("B", C.c_int32),
        ("n", C.c_int32),
        ("n_fft", C.c_int32),
        ("hop", C.c_int32),
        ("win", C.c_int32),
        ("n_mels", C.c_int32),
        ("frames", C.c_int32),
        ("clip", C.c_float),
    ]


class RowopDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("_pad", C.c_int32),
        ("d_x", C.c_void_p),
        ("d_vec", C.c_void_p),
        ("out", Tensor),
        ("div", C.c_float),
        ("B", C.c_int32),
        ("L", C.c_int32),
        ("C", C.c_int32),
        ("x_pitch", C.c_int32),
        ("out_pitch", C.c_int32),
    ]


class DiffEmbedDesc(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("d_step", "d_table", "d_w1", "d_b1", "d_w2", "d_b2", "d_wd", "d_bd", "d_out")] + [
        (n, C.c_int32) for n in ("B", "emb", "fc", "C", "n_layers", "max_steps")] + [("d_step_f", C.c_void_p)]


class SampleDesc(C.Structure):
    _fields_ = [("mode", C.c_int32), ("clip", C.c_int32)] + [
        (n, C.c_void_p) for n in ("d_x", "d_x_out", "d_eps", "d_noise", "d_step", "d_sqrt_recip", "d_sqrt_recipm1", "d_coef1", "d_coef2", "d_logvar", "d_alphas_cumprod")] + [
        ("d_hist", C.c_void_p * 3), ("d_eps_save", C.c_void_p)] + [(n, C.c_int32) for n in ("interval", "combine", "B", "L", "n_mel", "n_steps")]


class _OpUnion(C.Union):
    _fields_ = [("pack", PackDesc), ("amp", AmpDesc), ("conv", ConvDesc), ("post", PostDesc), ("rowop", RowopDesc), ("diffembed", DiffEmbedDesc), ("sample", SampleDesc)]


class Op(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("u", _OpUnion)]


_lib = None

# every symbol include/bvg_b200.h declares (checked by tests/test_capi_symbols.py)
EXPORTS = [
    "bvg_abi_version",
    "bvg_last_error",
    "bvg_device_check",
    "bvg_sizeof_op",
    "bvg_sizeof_conv_weights",
    "bvg_amp_fwd",
    "bvg_conv_fwd",
    "bvg_conv_geometry",
    "bvg_conv_pack_bytes",
    "bvg_pack_conv_weights",
    "bvg_post_fwd",
    "bvg_pack_post_weights",
    "bvg_pack_mel",
    "bvg_tail_fwd",
    "bvg_stitch_fwd",
    "bvg_logmel_fwd",
    "bvg_rowop_fwd",
    "bvg_diffembed_fwd",
    "bvg_sample_fwd",
    "bvg_convert",
    "bvg_program_create",
    "bvg_program_run",
    "bvg_program_run_interleaved",
    "bvg_program_run_timed",
    "bvg_tuning_defaults",
    "bvg_program_set_pdl",
    "bvg_program_num_launches",
    "bvg_program_destroy",
]


def lib():
    """Load (once) and return the shared library; raises BvgError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BvgError(
            f"{LIB_PATH} not found: build it with `make -C svc_inference_pipeline_b200/csrc -j` "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no fallback path."
        )
    L = C.CDLL(LIB_PATH)
    L.bvg_last_error.restype = C.c_char_p
    L.bvg_sizeof_op.restype = C.c_size_t
    L.bvg_sizeof_conv_weights.restype = C.c_size_t
    L.bvg_abi_version.restype = C.c_int
    for name, argtypes in {
        "bvg_device_check": [C.c_int],
        "bvg_tuning_defaults": [C.POINTER(Tuning)],
        "bvg_amp_fwd": [C.POINTER(AmpDesc), C.c_void_p],
        "bvg_conv_fwd": [C.POINTER(ConvDesc), C.c_void_p],
        "bvg_post_fwd": [C.POINTER(PostDesc), C.c_void_p],
        "bvg_pack_mel": [C.POINTER(PackDesc), C.c_void_p],
        "bvg_tail_fwd": [C.POINTER(TailDesc), C.c_void_p],
        "bvg_stitch_fwd": [C.POINTER(StitchDesc), C.c_void_p],
        "bvg_logmel_fwd": [C.POINTER(LogmelDesc), C.c_void_p],
        "bvg_rowop_fwd": [C.POINTER(RowopDesc), C.c_void_p],
        "bvg_diffembed_fwd": [C.POINTER(DiffEmbedDesc), C.c_void_p],
        "bvg_sample_fwd": [C.POINTER(SampleDesc), C.c_void_p],
        "bvg_convert": [C.POINTER(Tensor), C.POINTER(Tensor), C.c_size_t, C.c_void_p],
        "bvg_conv_geometry": [C.POINTER(ConvGeom), C.POINTER(ConvWeights)],
        "bvg_conv_pack_bytes": [C.POINTER(ConvGeom), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)],
        "bvg_pack_conv_weights": [C.POINTER(ConvGeom), C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(ConvWeights), C.c_void_p, C.c_void_p, C.c_void_p],
        "bvg_pack_post_weights": [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p],
        "bvg_program_create": [C.POINTER(Op), C.c_int32, C.POINTER(C.c_void_p)],
        "bvg_program_run": [C.c_void_p, C.c_void_p],
        "bvg_program_run_interleaved": [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int32],
        "bvg_program_run_timed": [C.c_void_p, C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_float)],
        "bvg_program_set_pdl": [C.c_void_p, C.c_int],
        "bvg_program_num_launches": [C.c_void_p],
        "bvg_program_destroy": [C.c_void_p],
    }.items():
        fn = getattr(L, name)
        fn.argtypes = argtypes
        fn.restype = None if name in ("bvg_program_destroy", "bvg_tuning_defaults") else C.c_int
    if L.bvg_abi_version() != ABI_VERSION:
        raise BvgError(f"ABI mismatch: library reports version {L.bvg_abi_version()}, binding expects {ABI_VERSION}")
    if L.bvg_sizeof_op() != C.sizeof(Op) or L.bvg_sizeof_conv_weights() != C.sizeof(ConvWeights):
        raise BvgError(
            f"struct layout mismatch: C sizeof(bvg_op)={L.bvg_sizeof_op()} vs ctypes {C.sizeof(Op)}; "
            f"sizeof(bvg_conv_weights)={L.bvg_sizeof_conv_weights()} vs ctypes {C.sizeof(ConvWeights)}"
        )
    _lib = L
    return L


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().bvg_last_error().decode("utf-8", "replace")
        raise BvgError(f"{what or 'libbvg_b200'} failed (code {rc}): {msg}")


_checked_devices = set()


def require_sm100(device_index: int) -> None:
    """Fail loudly unless ``device_index`` is a compute-capability-10.x GPU (checked once per device:
    cudaGetDeviceProperties costs ~100 ms)."""
    if device_index not in _checked_devices:
        check(lib().bvg_device_check(int(device_index)), "device_check")
        _checked_devices.add(device_index)


# ---- test / A-B knobs -----------------------------------------------------------------------------------
# The library has no global knobs: a ``bvg_tuning`` travels with each descriptor.  For tests and the A/B tools the
# BINDING keeps one process-wide ``Tuning`` object that the descriptor builders (ops.py, modules/bigvgan.py) attach
# when any field differs from the defaults; programs deep-copy it at creation.
_tuning = None
_tuning_defaults = None


def _tuning_state():
    global _tuning, _tuning_defaults
    if _tuning is None:
        _tuning, _tuning_defaults = Tuning(), Tuning()
        lib().bvg_tuning_defaults(C.byref(_tuning))
        lib().bvg_tuning_defaults(C.byref(_tuning_defaults))
    return _tuning, _tuning_defaults


def set_tuning(name: str, value: int) -> None:
    """Set one knob of the binding's ``Tuning`` object (applies to descriptors built afterwards)."""
    t, _ = _tuning_state()
    if name not in dict(Tuning._fields_) or name.startswith("_"):
        raise BvgError(f"unknown tuning knob '{name}'")
    setattr(t, name, int(value))


def reset_tuning() -> None:
    t, _ = _tuning_state()
    lib().bvg_tuning_defaults(C.byref(t))


def new_tuning(**knobs):
    """A ``Tuning`` of the caller's own: the binding's current knobs with ``knobs`` on top (pass ``ctypes.pointer`` of it as
    a descriptor's ``tune``; geometry and programs read it at creation and keep their own copy)."""
    t, _ = _tuning_state()
    own = Tuning()
    C.memmove(C.byref(own), C.byref(t), C.sizeof(Tuning))
    for name, value in knobs.items():
        if name not in dict(Tuning._fields_) or name.startswith("_"):
            raise BvgError(f"unknown tuning knob '{name}'")
        setattr(own, name, int(value))
    return own


def tuning_ptr():
    """Pointer to the binding's ``Tuning`` for a descriptor's ``tune`` field, or NULL when every knob is at its default."""
    t, d = _tuning_state()
    if bytes(t) == bytes(d):
        return None
    return C.pointer(t)
