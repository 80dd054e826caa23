"""Time / batch sharding of the vocoder across chunks and GPUs.

The reference vocodes one whole utterance in one forward (``modules/bigvgan_inference.py:34-36``)
on one device and has no distributed code.  BigVGAN is convolutional with a finite receptive
field -- exactly +-38 mel frames for the repo generator -- so the path shards trivially:

* **batch**: utterances are independent -> split them across ranks, no communication except
  returning the waveforms (one ``all_gather``);
* **time** (long-form): cut the mel into contiguous ranges, extend each by a halo of ``H >= 38``
  frames per interior side, vocode the pieces independently, drop the halo and cross-fade
  linearly over ``fade_frames`` around every cut.  With ``H = 48`` a 20-frame fade window lies
  inside the region where *both* neighbours are exact, so in fp32 the fade blends two copies of
  the same samples (difference ~1e-7) and in bf16 it hides their rounding difference.  True
  sequence ends keep the reference's own zero / replicate padding.

One process per GPU; the only collective is a single ``all_gather`` of fp32 waveform pieces
(43 MB per rank for an hour of 24 kHz audio) over NCCL / NVLink.  Everything here is
model-agnostic (``model(mel[B, C, T]) -> wave[B, 1, T * hop]``), so the logic is tested on CPU with
a small convolutional stand-in and the ``gloo`` backend.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import torch

__all__ = [
    "RECEPTIVE_FIELD_FRAMES",
    "receptive_field_frames",
    "Chunk",
    "chunk_plan",
    "balanced_plan",
    "shard_range",
    "stitch",
    "vocode_long",
    "vocode_long_distributed",
    "vocode_batch_distributed",
]

RECEPTIVE_FIELD_FRAMES = 38  # repo generator, measured with fp64 autograd; == receptive_field_frames(repo cfg)


def _cfg(vcfg, key):
    return vcfg[key] if isinstance(vcfg, dict) else getattr(vcfg, key)


def receptive_field_frames(vcfg) -> int:
    """One-sided receptive field of a BigVGAN generator in mel frames, from its hyper-parameters (reference
    ``modules/bigvgan.py:521-598``): the halo a time chunk needs so that its interior equals the unchunked forward.
    Walks the network backwards from the first and the last sample of one output frame, carrying the exact interval
    ``[a, b]`` of samples needed at the current rate (interval ends are monotonic in the output position, so the two
    extreme samples of a frame bound every sample in it):

    * ``conv_post`` / ``conv_pre`` (k = 7): ``[a - 3, b + 3]``;
    * every ``Activation1d`` (2x kaiser-sinc up, snake, 12-tap low-pass down): ``[a - 5, b + 5]``;
    * a dilated conv ``(k, d)``: ``[a - d (k - 1) / 2, b + d (k - 1) / 2]``; AMPBlock1 layer = act, conv(k, d), act,
      conv(k, 1); AMPBlock2 layer = act, conv(k, d); the widest of the parallel resblocks of a stage counts;
    * ``ConvTranspose1d(k, u, padding = (k - u) // 2)``: output ``t = q u + j - pad`` reads inputs
      ``ceil((t + pad - (k - 1)) / u) .. floor((t + pad) / u)``.

    38 for the repo generator, as measured with fp64 autograd on the reference.
    """
    rates, uks = list(_cfg(vcfg, "upsample_rates")), list(_cfg(vcfg, "upsample_kernel_sizes"))
    rks, rds = list(_cfg(vcfg, "resblock_kernel_sizes")), list(_cfg(vcfg, "resblock_dilation_sizes"))
    block1 = str(_cfg(vcfg, "resblock")) == "1"
    hop = 1
    for u in rates:
        hop *= int(u)
    widest = 0
    for k, dils in zip(rks, rds):
        tot = 0
        for d in dils:
            tot += 5 + d * (k - 1) // 2
            if block1:
                tot += 5 + (k - 1) // 2
        widest = max(widest, tot)
    frame = 1 << 20  # far from either end
    a, b = frame * hop, frame * hop + hop - 1
    a, b = a - 3 - 5, b + 3 + 5  # conv_post, activation_post
    for u, ku in zip(reversed(rates), reversed(uks)):
        a, b = a - widest, b + widest
        pad = (ku - u) // 2
        a, b = -((-(a + pad - (ku - 1))) // u), (b + pad) // u  # ceil, floor
    a, b = a - 3, b + 3  # conv_pre
    return max(frame - a, b - frame)


@dataclass(frozen=True)
class Chunk:
    start: int   # first frame this chunk is responsible for
    end: int     # one past the last
    in_lo: int   # input window [in_lo, in_hi) = [start - halo, end + halo) clipped to the sequence
    in_hi: int
    keep_lo: int  # frames whose samples are kept: [start - fade/2, end + fade/2) clipped
    keep_hi: int


def chunk_plan(total_frames: int, chunk_frames: int, halo: int = 48, fade_frames: int = 20, lo: int = 0, hi: Optional[int] = None,
               receptive_field: int = RECEPTIVE_FIELD_FRAMES) -> List[Chunk]:
    """Partition frames ``[lo, hi)`` of a ``total_frames`` sequence into chunks of ``chunk_frames``.
    ``receptive_field``: the model's one-sided receptive field in frames (``receptive_field_frames(cfg)`` /
    ``Generator.receptive_field_frames()``; default = the repo generator's 38)."""
    hi = total_frames if hi is None else hi
    if chunk_frames <= 0:
        raise ValueError("chunk_frames must be positive")
    if fade_frames % 2 or fade_frames < 0:
        raise ValueError("fade_frames must be even and non-negative")
    if fade_frames // 2 + receptive_field > halo and total_frames > chunk_frames:
        raise ValueError(f"halo {halo} too small: need >= {receptive_field} (receptive field) + fade/2 = {fade_frames // 2}")
    if chunk_frames < fade_frames:
        raise ValueError("chunk_frames must be at least fade_frames")
    fh = fade_frames // 2
    out = []
    s = lo
    while s < hi:
        e = min(s + chunk_frames, hi)
        if hi - e < fade_frames and e < hi:  # do not leave a sliver shorter than the fade window
            e = hi
        out.append(Chunk(s, e, max(0, s - halo), min(total_frames, e + halo), max(0, s - fh), min(total_frames, e + fh)))
        s = e
    return out


def shard_range(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced split of ``range(total)`` (first ``total % world`` ranks get one extra)."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _fade_weights(n: int, device, dtype) -> torch.Tensor:
    # n samples of a linear ramp that sums to 1 with its mirror: w[i] + w[n-1-i] = 1
    return (torch.arange(n, device=device, dtype=dtype) + 0.5) / n


def stitch(pieces: Sequence[Tuple[Chunk, torch.Tensor]], total_frames: int, hop: int, fade_frames: int = 20) -> torch.Tensor:
    """Overlap-add ``(chunk, wave[keep_lo*hop : keep_hi*hop])`` pieces with linear cross-fades at the cuts."""
    if not pieces:
        return torch.zeros(total_frames * hop)
    ref = pieces[0][1]
    out = torch.zeros(total_frames * hop, dtype=ref.dtype, device=ref.device)
    fh = fade_frames // 2
    for ch, wave in pieces:
        n = (ch.keep_hi - ch.keep_lo) * hop
        if wave.numel() != n:
            raise ValueError(f"piece for frames [{ch.keep_lo}, {ch.keep_hi}) has {wave.numel()} samples, expected {n}")
        w = wave
        if fade_frames:
            w = wave.clone()
            if ch.start > 0:  # ramp up across [start - fh, start + fh)
                k = (min(ch.start + fh, ch.keep_hi) - ch.keep_lo) * hop
                full = _fade_weights(fade_frames * hop, wave.device, wave.dtype)
                off = (ch.keep_lo - (ch.start - fh)) * hop
                w[:k] *= full[off : off + k]
            if ch.end < total_frames:  # ramp down across [end - fh, end + fh)
                a = (max(ch.end - fh, ch.keep_lo) - ch.keep_lo) * hop
                full = 1.0 - _fade_weights(fade_frames * hop, wave.device, wave.dtype)
                off = (max(ch.end - fh, ch.keep_lo) - (ch.end - fh)) * hop
                w[a:] *= full[off : off + (n - a)]
        out[ch.keep_lo * hop : ch.keep_hi * hop] += w
    return out


def balanced_plan(total_frames: int, chunk_frames: int = 4096, halo: int = 48, fade_frames: int = 20, world: int = 1,
                  receptive_field: int = RECEPTIVE_FIELD_FRAMES) -> List[Chunk]:
    """Chunk plan for the long-form path: the number of chunks is a multiple of ``world`` (every rank gets the same
    count), the chunks are equally long to within one frame (none longer than ``chunk_frames``), and every input
    window has the SAME length ``W = longest chunk + 2 * halo`` -- a chunk at a sequence end, which has no halo on
    that side, takes the spare frames as extra context on the other -- so one generator program (one batch shape)
    serves every chunk.  With 4096-frame chunks the halo is 2.3 % of the work (938-frame chunks: 10 %)."""
    if chunk_frames <= 0 or world <= 0:
        raise ValueError("chunk_frames and world must be positive")
    if fade_frames % 2 or fade_frames < 0:
        raise ValueError("fade_frames must be even and non-negative")
    n = world * max(1, -(-total_frames // (world * chunk_frames)))
    shortest_ok = max(fade_frames, 1)  # a chunk must hold its own fade windows
    while n > world and total_frames // n < shortest_ok:
        n -= world
    if total_frames // n < shortest_ok:  # fewer chunks than ranks: the last ranks stay idle
        n = max(1, total_frames // shortest_ok)
    cuts = [(k * total_frames) // n for k in range(n + 1)]
    if n > 1 and fade_frames // 2 + receptive_field > halo:
        raise ValueError(f"halo {halo} too small: need >= {receptive_field} (receptive field) + fade/2 = {fade_frames // 2}")
    longest = max(cuts[k + 1] - cuts[k] for k in range(n))
    W = min(total_frames, longest + 2 * halo)
    fh = fade_frames // 2
    out = []
    for k in range(n):
        s_, e_ = cuts[k], cuts[k + 1]
        in_lo = min(max(0, s_ - halo), total_frames - W)
        out.append(Chunk(s_, e_, in_lo, in_lo + W, max(0, s_ - fh), min(total_frames, e_ + fh)))
    return out


def _stitch_rows(part: Sequence[Chunk], total_frames: int, hop: int, fade_frames: int, span_lo: int):
    """Rows {dst, src, n, fade_in, fade_out} (samples) of ``bvg_stitch_desc`` for chunks that went through the model
    as one batch of equal windows; ``dst`` is relative to frame ``span_lo``."""
    rows = []
    for c in part:
        n = (c.keep_hi - c.keep_lo) * hop
        fin = fade_frames * hop if (c.start > 0 and fade_frames) else 0
        fout = fade_frames * hop if (c.end < total_frames and fade_frames) else 0
        rows.append([(c.keep_lo - span_lo) * hop, (c.keep_lo - c.in_lo) * hop, n, fin, fout])
    return rows


def _stitch_batch(out: torch.Tensor, y: torch.Tensor, part: Sequence[Chunk], total_frames: int, hop: int, fade_frames: int, span_lo: int) -> None:
    """Cross-fade the model output ``y [b, 1, W * hop]`` of one batch of chunks into ``out`` (zero-filled by the
    caller; sample 0 = frame ``span_lo``): the ``bvg_stitch_fwd`` kernel for CUDA tensors, the same arithmetic
    with torch slices on the host otherwise (CPU models: the toy vocoder of the gloo tests)."""
    rows = _stitch_rows(part, total_frames, hop, fade_frames, span_lo)
    if y.is_cuda:
        import ctypes as C

        from . import _lib as L

        y2 = y.reshape(len(part), -1)
        if y2.dtype != torch.float32 or out.dtype != torch.float32:
            raise ValueError("stitch: model output and destination must be float32")
        y2 = y2.contiguous()
        h = torch.tensor(rows, dtype=torch.int64)
        t = h.to(y.device, non_blocking=False)
        d = L.StitchDesc()
        d.d_wave, d.wave_stride = y2.data_ptr(), y2.stride(0)
        d.d_out, d.out_len = out.data_ptr(), out.numel()
        d.d_table, d.h_table, d.n_chunks = t.data_ptr(), h.data_ptr(), len(part)
        with torch.cuda.device(y.device):
            L.check(L.lib().bvg_stitch_fwd(C.byref(d), torch.cuda.current_stream(y.device).cuda_stream), "stitch_fwd")
        return
    for j, (dst, src, n, fin, fout) in enumerate(rows):
        w = y[j].reshape(-1)[src : src + n]
        o = out[dst : dst + n]
        if fin:
            o[:fin] += w[:fin] * _fade_weights(fin, w.device, w.dtype)
        if fout:
            o[n - fout :] += w[n - fout :] * (1.0 - _fade_weights(fout, w.device, w.dtype))
        o[fin : n - fout] = w[fin : n - fout]


def _forward(model: Callable, x: torch.Tensor) -> torch.Tensor:
    # the B200 Generator lends its output buffer (consumed at once by the stitch on the same stream)
    return getattr(model, "forward_borrowed", model)(x)


def _run_chunks(model: Callable, mel: torch.Tensor, chunks: Sequence[Chunk], hop: int, batch_chunks: int) -> List[Tuple[Chunk, torch.Tensor]]:
    """Vocode the chunks of one ``[C, T]`` mel; equal-length input windows are batched.  Returns per-chunk pieces
    (the kept samples) for ``stitch``: the path for ragged plans (``chunk_plan``)."""
    by_len = {}
    for c in chunks:
        by_len.setdefault(c.in_hi - c.in_lo, []).append(c)
    done = {}
    for ln, group in by_len.items():
        for i in range(0, len(group), max(1, batch_chunks)):
            part = group[i : i + max(1, batch_chunks)]
            x = torch.stack([mel[:, c.in_lo : c.in_hi] for c in part]).contiguous()
            y = model(x)  # [b, 1, ln * hop]
            for j, c in enumerate(part):
                a = (c.keep_lo - c.in_lo) * hop
                done[c.start] = (c, y[j, 0, a : a + (c.keep_hi - c.keep_lo) * hop])
    return [done[c.start] for c in chunks]


def _vocode_span(model: Callable, mel: torch.Tensor, chunks: Sequence[Chunk], total: int, hop: int, fade_frames: int, batch_chunks: int,
                 out: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, int]:
    """Vocode consecutive chunks of a balanced plan and cross-fade them into one zero-initialised buffer that
    starts at frame ``chunks[0].keep_lo``.  Returns ``(buffer, span_lo)``."""
    span_lo, span_hi = chunks[0].keep_lo, chunks[-1].keep_hi
    n = (span_hi - span_lo) * hop
    if out is not None:
        out[:n].zero_()
    for i in range(0, len(chunks), max(1, batch_chunks)):
        part = chunks[i : i + max(1, batch_chunks)]
        x = torch.stack([mel[:, c.in_lo : c.in_hi] for c in part]).contiguous()
        y = _forward(model, x)
        if out is None:  # the model's own output dtype (fp32 for the generator)
            out = torch.zeros(n, dtype=y.dtype, device=y.device)
        _stitch_batch(out, y, part, total, hop, fade_frames, span_lo)
    return out, span_lo


def _rf(model, receptive_field):
    if receptive_field is not None:
        return int(receptive_field)
    fn = getattr(model, "receptive_field_frames", None)
    return int(fn()) if callable(fn) else RECEPTIVE_FIELD_FRAMES


@torch.no_grad()
def vocode_long(model: Callable, mel: torch.Tensor, hop: int, chunk_frames: int = 4096, halo: int = 48, fade_frames: int = 20, batch_chunks: int = 8,
                receptive_field: Optional[int] = None) -> torch.Tensor:
    """Chunked vocoding of one long ``[C, T]`` mel on the model's device; returns ``[T * hop]``.  ``receptive_field``
    defaults to ``model.receptive_field_frames()`` (derived from the generator's config) and is checked against the halo."""
    total = mel.shape[-1]
    chunks = balanced_plan(total, chunk_frames, halo, fade_frames, 1, _rf(model, receptive_field))
    return _vocode_span(model, mel, chunks, total, hop, fade_frames, batch_chunks)[0]


def _world(group):
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


@torch.no_grad()
def vocode_long_distributed(model: Callable, mel: torch.Tensor, hop: int, chunk_frames: int = 4096, halo: int = 48, fade_frames: int = 20,
                            batch_chunks: int = 16, group=None, receptive_field: Optional[int] = None, gather: str = "all") -> torch.Tensor:
    """Long-form vocoding sharded along time across the ranks of ``group``.

    Every rank holds the whole mel (2.5 MB per minute), vocodes the same number of equal chunks (``balanced_plan``)
    -- a contiguous range, cross-faded on the device into one span buffer by ``bvg_stitch_fwd`` -- and contributes
    that span to ONE collective; neighbouring spans overlap only in a fade window, where each carries its half of
    the blend, so the result is the sum of the spans.  ``gather="all"``: ``all_gather``, the full ``[T * hop]``
    waveform on every rank; ``gather="root"``: ``gather`` to rank 0 (the others return their own span only).
    """
    import torch.distributed as dist

    world, rank = _world(group)
    total = mel.shape[-1]
    chunks = balanced_plan(total, chunk_frames, halo, fade_frames, world, _rf(model, receptive_field))
    per = len(chunks) // world
    mine = chunks[rank * per : (rank + 1) * per] if len(chunks) % world == 0 else chunks[slice(*shard_range(len(chunks), world, rank))]
    if world == 1:
        return _vocode_span(model, mel, chunks, total, hop, fade_frames, batch_chunks)[0]
    spans = []
    for r in range(world):
        a, b = (r * per, (r + 1) * per) if len(chunks) % world == 0 else shard_range(len(chunks), world, r)
        spans.append((chunks[a].keep_lo, chunks[b - 1].keep_hi) if b > a else (0, 0))
    max_len = max(h - l for l, h in spans) * hop
    dev = mel.device
    local = torch.zeros(max_len, dtype=torch.float32, device=dev)
    if mine:
        _vocode_span(model, mel, mine, total, hop, fade_frames, batch_chunks, out=local)
    if gather == "root":
        parts = [torch.empty(max_len, dtype=torch.float32, device=dev) for _ in range(world)] if rank == 0 else None
        dist.gather(local, parts, dst=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        if rank != 0:
            return local[: (spans[rank][1] - spans[rank][0]) * hop]
        gathered = parts
    else:
        flat = torch.empty(world * max_len, dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(flat, local, group=group)
        gathered = [flat[r * max_len : (r + 1) * max_len] for r in range(world)]
    out = torch.zeros(total * hop, dtype=torch.float32, device=dev)
    for r, (lo, hi) in enumerate(spans):
        out[lo * hop : hi * hop] += gathered[r][: (hi - lo) * hop]
    return out


@torch.no_grad()
def vocode_batch_distributed(model: Callable, mels: torch.Tensor, hop: int, group=None) -> torch.Tensor:
    """Batch of equal-length utterances ``[B, C, T]`` split across ranks; returns all ``[B, T * hop]`` waveforms
    on every rank (one ``all_gather``).  Ranks may receive unequal shares when ``B % world != 0``."""
    import torch.distributed as dist

    world, rank = _world(group)
    B, _, T = mels.shape
    lo, hi = shard_range(B, world, rank)
    dev = mels.device
    y = model(mels[lo:hi].contiguous()).reshape(hi - lo, T * hop) if hi > lo else torch.zeros(0, T * hop, device=dev)
    if world == 1:
        return y
    per = -(-B // world)
    local = torch.zeros(per, T * hop, dtype=torch.float32, device=dev)
    local[: hi - lo] = y
    gathered = torch.empty(world * per, T * hop, dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(gathered, local, group=group)
    rows = []
    for r in range(world):
        a, b = shard_range(B, world, r)
        rows.append(gathered[r * per : r * per + (b - a)])
    return torch.cat(rows)
