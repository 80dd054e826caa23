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
    "Chunk",
    "chunk_plan",
    "shard_range",
    "stitch",
    "vocode_long",
    "vocode_long_distributed",
    "vocode_batch_distributed",
]

RECEPTIVE_FIELD_FRAMES = 38  # repo generator, measured with fp64 autograd


@dataclass(frozen=True)
class Chunk:
    start: int   # first frame this chunk is responsible for
    end: int     # one past the last
    in_lo: int   # input window [in_lo, in_hi) = [start - halo, end + halo) clipped to the sequence
    in_hi: int
    keep_lo: int  # frames whose samples are kept: [start - fade/2, end + fade/2) clipped
    keep_hi: int


def chunk_plan(total_frames: int, chunk_frames: int, halo: int = 48, fade_frames: int = 20, lo: int = 0, hi: Optional[int] = None) -> List[Chunk]:
    """Partition frames ``[lo, hi)`` of a ``total_frames`` sequence into chunks of ``chunk_frames``."""
    hi = total_frames if hi is None else hi
    if chunk_frames <= 0:
        raise ValueError("chunk_frames must be positive")
    if fade_frames % 2 or fade_frames < 0:
        raise ValueError("fade_frames must be even and non-negative")
    if fade_frames // 2 + RECEPTIVE_FIELD_FRAMES > halo and total_frames > chunk_frames:
        raise ValueError(f"halo {halo} too small: need >= {RECEPTIVE_FIELD_FRAMES} (receptive field) + fade/2 = {fade_frames // 2}")
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


def _run_chunks(model: Callable, mel: torch.Tensor, chunks: Sequence[Chunk], hop: int, batch_chunks: int) -> List[Tuple[Chunk, torch.Tensor]]:
    """Vocode the chunks of one ``[C, T]`` mel; equal-length input windows are batched."""
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


@torch.no_grad()
def vocode_long(model: Callable, mel: torch.Tensor, hop: int, chunk_frames: int = 4096, halo: int = 48, fade_frames: int = 20, batch_chunks: int = 8) -> torch.Tensor:
    """Chunked vocoding of one long ``[C, T]`` mel on the model's device; returns ``[T * hop]``."""
    total = mel.shape[-1]
    chunks = chunk_plan(total, chunk_frames, halo, fade_frames)
    return stitch(_run_chunks(model, mel, chunks, hop, batch_chunks), total, hop, fade_frames)


def _world(group):
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


@torch.no_grad()
def vocode_long_distributed(model: Callable, mel: torch.Tensor, hop: int, chunk_frames: int = 4096, halo: int = 48, fade_frames: int = 20,
                            batch_chunks: int = 8, group=None) -> torch.Tensor:
    """Long-form vocoding sharded along time across the ranks of ``group``.

    Every rank holds the whole mel (2.5 MB per minute), vocodes a contiguous range of chunks and
    contributes its already cross-faded partial waveform to ONE ``all_gather``; the ranges only
    overlap in the fade windows, where the partial sums add up.  Returns the full ``[T * hop]``
    waveform on every rank.
    """
    import torch.distributed as dist

    world, rank = _world(group)
    total = mel.shape[-1]
    chunks = chunk_plan(total, chunk_frames, halo, fade_frames)
    lo, hi = shard_range(len(chunks), world, rank)
    mine = chunks[lo:hi]
    pieces = _run_chunks(model, mel, mine, hop, batch_chunks) if mine else []
    if world == 1:
        return stitch(pieces, total, hop, fade_frames)
    # local partial sum over this rank's span (cuts inside the span are fully cross-faded; the two
    # outer fade windows carry this rank's half of the blend)
    span_lo, span_hi = [], []
    for r in range(world):
        a, b = shard_range(len(chunks), world, r)
        span_lo.append(chunks[a].keep_lo if b > a else 0)
        span_hi.append(chunks[b - 1].keep_hi if b > a else 0)
    max_len = max(h - l for l, h in zip(span_lo, span_hi)) * hop
    dev = mel.device
    local = torch.zeros(max_len, dtype=torch.float32, device=dev)
    if mine:
        part = stitch(pieces, total, hop, fade_frames)
        local[: (span_hi[rank] - span_lo[rank]) * hop] = part[span_lo[rank] * hop : span_hi[rank] * hop]
    gathered = torch.empty(world * max_len, dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(gathered, local, group=group)
    out = torch.zeros(total * hop, dtype=torch.float32, device=dev)
    for r in range(world):
        n = (span_hi[r] - span_lo[r]) * hop
        out[span_lo[r] * hop : span_hi[r] * hop] += gathered[r * max_len : r * max_len + n]
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
