"""Chunk / shard / stitch logic on CPU with a small convolutional stand-in for the vocoder
(receptive field < 38 frames, hop 8) and a 2-rank gloo group.  The GPU version of the same
property (chunked == unchunked on the real generator) is in tests/test_gpu_generator.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from svc_inference_pipeline_b200 import sharding as S

HOP = 8


class ToyVocoder(torch.nn.Module):
    """mel [B, 6, T] -> wave [B, 1, 8T]; receptive field +-(3 + 2 + 1*...) << 38 frames."""

    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(0)
        self.pre = torch.nn.Conv1d(6, 8, 7, padding=3)
        self.up = torch.nn.ConvTranspose1d(8, 4, 16, 8, padding=4)
        self.post = torch.nn.Conv1d(4, 1, 9, padding=4)
        for p in self.parameters():
            p.data = torch.randn(p.shape, generator=g) * 0.3
        self.double()

    def forward(self, x):
        return torch.tanh(self.post(torch.sin(self.up(torch.tanh(self.pre(x.double()))))))


def test_chunk_plan_covers_and_halo():
    chunks = S.chunk_plan(1000, 256, halo=48, fade_frames=20)
    assert chunks[0].start == 0 and chunks[-1].end == 1000
    assert all(a.end == b.start for a, b in zip(chunks, chunks[1:]))
    assert chunks[0].in_lo == 0 and chunks[0].in_hi == 256 + 48 and chunks[0].keep_lo == 0
    assert chunks[1].in_lo == 256 - 48 and chunks[1].keep_lo == 256 - 10 and chunks[1].keep_hi == 512 + 10
    assert chunks[-1].in_hi == 1000 and chunks[-1].keep_hi == 1000
    # no sliver shorter than the fade window
    assert S.chunk_plan(266, 256)[-1].end - S.chunk_plan(266, 256)[-1].start >= 20
    with pytest.raises(ValueError):
        S.chunk_plan(1000, 256, halo=40, fade_frames=20)  # 38 + 10 > 40
    assert [S.shard_range(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert S.shard_range(2, 4, 3) == (2, 2)


def test_balanced_plan():
    """Long-form plan: chunk count a multiple of the world size, equal lengths (+-1), one window length for all."""
    for total, chunk, world in [(337500, 4096, 8), (1000, 256, 2), (517, 100, 3), (90, 40, 1), (60, 4096, 4), (25, 10, 2)]:
        chunks = S.balanced_plan(total, chunk, halo=48, fade_frames=20, world=world)
        assert chunks[0].start == 0 and chunks[-1].end == total
        assert all(a.end == b.start for a, b in zip(chunks, chunks[1:]))
        lens = [c.end - c.start for c in chunks]
        assert max(lens) - min(lens) <= 1 and max(lens) < max(chunk + 1, 40) and min(lens) >= min(20, total)
        assert len({c.in_hi - c.in_lo for c in chunks}) == 1
        if total >= world * 20:
            assert len(chunks) % world == 0
        for c in chunks:
            assert 0 <= c.in_lo <= c.keep_lo <= c.start < c.end <= c.keep_hi <= c.in_hi <= total
            assert c.in_lo == 0 or c.in_lo <= c.start - 48
            assert c.in_hi == total or c.in_hi >= c.end + 48
    hour = S.balanced_plan(337500, 4096, world=8)
    assert len(hour) == 88 and hour[0].in_hi - hour[0].in_lo == 3836 + 96
    with pytest.raises(ValueError):
        S.balanced_plan(1000, 256, halo=40, fade_frames=20)
    S.balanced_plan(1000, 256, halo=40, fade_frames=20, receptive_field=21)  # the 512x generator's field fits


@pytest.mark.parametrize("T,chunk", [(1000, 256), (517, 100), (300, 300), (90, 40), (30, 8)])
def test_chunked_equals_unchunked(T, chunk):
    model = ToyVocoder()
    mel = torch.randn(6, T, generator=torch.Generator().manual_seed(1)).double()
    full = model(mel[None])[0, 0]
    out = S.vocode_long(model, mel, HOP, chunk_frames=chunk, batch_chunks=3)
    assert out.shape == full.shape
    assert (out - full).abs().max().item() < 1e-12


def test_stitch_crossfade_weights_sum_to_one():
    total, hop = 120, 4
    chunks = S.chunk_plan(total, 40, halo=48, fade_frames=20)
    ones = [(c, torch.ones((c.keep_hi - c.keep_lo) * hop, dtype=torch.float64)) for c in chunks]
    assert (S.stitch(ones, total, hop, 20) - 1).abs().max().item() < 1e-12
    # the blend really is a cross-fade: piece k contributes 1 - w, piece k+1 contributes w
    a = [(c, torch.full(((c.keep_hi - c.keep_lo) * hop,), float(i), dtype=torch.float64)) for i, c in enumerate(chunks)]
    y = S.stitch(a, total, hop, 20)
    assert y[0] == 0 and y[-1] == len(chunks) - 1
    seg = y[(40 - 10) * hop : (40 + 10) * hop]
    assert torch.all(seg[1:] > seg[:-1]) and 0 < seg[0] < 0.02 and 0.98 < seg[-1] < 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, T, chunk, B, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        model = ToyVocoder()
        mel = torch.randn(6, T, generator=torch.Generator().manual_seed(1)).double()
        full = model(mel[None])[0, 0]
        out = S.vocode_long_distributed(model, mel, HOP, chunk_frames=chunk, batch_chunks=2)
        err_long = (out.double() - full).abs().max().item()
        mels = torch.randn(B, 6, 50, generator=torch.Generator().manual_seed(2)).double()
        allw = S.vocode_batch_distributed(model, mels, HOP)
        err_batch = (allw.double() - model(mels)[:, 0]).abs().max().item()
        ret[rank] = (err_long, err_batch, tuple(out.shape), tuple(allw.shape))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,T,chunk,B", [(2, 700, 128, 5), (2, 150, 150, 1), (3, 400, 64, 4)])
def test_distributed_gloo(world, T, chunk, B):
    """N > 1 path: time-sharded long-form + batch-sharded vocoding, one all_gather each."""
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), T, chunk, B, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        err_long, err_batch, shp, shpb = ret[r]
        assert shp == (T * HOP,) and shpb == (B, 50 * HOP)
        assert err_long < 1e-6 and err_batch < 1e-6  # the gather buffer is fp32


@pytest.mark.parametrize("variant", ["b1_snakebeta_log", "b2_snake_lin"])
def test_receptive_field_from_cfg(variant):
    """sharding.receptive_field_frames(cfg) is exact: the gradient of one output sample of the PyTorch-CPU port
    with respect to the mel is non-zero in frames [f - rf, f + rf] and nowhere else, and reaches both ends.  38 for the repo generator (the value
    SURVEY.md appendix A measured on the reference with fp64 autograd), 21 for the 512x v2 generator."""
    import numpy as np

    from oracle import bigvgan_torch_cpu as port
    from svc_inference_pipeline_b200.sharding import RECEPTIVE_FIELD_FRAMES, receptive_field_frames
    from util_cases import REPO, V2, tiny_cfg_sd

    assert receptive_field_frames(REPO) == RECEPTIVE_FIELD_FRAMES == 38
    assert receptive_field_frames(V2) == 21
    cfg, sd = tiny_cfg_sd(variant)
    rf = receptive_field_frames(cfg)
    hop = int(np.prod(cfg["upsample_rates"]))
    sd = {k: torch.from_numpy(v).double() for k, v in sd.items()}
    T, f0 = 2 * rf + 21, rf + 10
    # structural support by autograd (a perturbation test underestimates it: the outermost paths run through the
    # 0.002-sized end taps of every anti-aliasing filter and vanish below fp64 resolution of the output)
    lo, hi = [], []
    for t in (f0 * hop, f0 * hop + hop - 1):
        mel = torch.randn(1, cfg["input_dim"], T, dtype=torch.float64, generator=torch.Generator().manual_seed(3)).requires_grad_(True)
        with torch.enable_grad():
            y = port.generator_forward.__wrapped__(sd, cfg, mel)
            y[0, 0, t].backward()
        cols = torch.nonzero(mel.grad[0].abs().amax(dim=0) > 0).reshape(-1)
        lo.append(f0 - int(cols.min()))
        hi.append(int(cols.max()) - f0)
    assert max(lo + hi) == rf, (rf, lo, hi)
    assert lo[0] == rf and hi[1] == rf  # the first sample of a frame reaches furthest back, the last furthest ahead
