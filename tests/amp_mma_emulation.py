"""Lane-level numpy emulation of the tensor-core formulation of the fused Activation1d kernel
(svc_inference_pipeline_b200/csrc/amp_mma.cu).  It mirrors the CUDA kernel's index arithmetic --
the banded-Toeplitz B fragments of both FIRs, the ldmatrix.trans / mma.m16n8k16 / D->A fragment
chaining, the 3-row offset of the output tiles and the two replicate clamps -- with exact (float64)
arithmetic, so that a mismatch against the oracle is an indexing bug, not rounding.

    python tests/amp_mma_emulation.py          # self-check against oracle/bigvgan_oracle.py
(also run by tests/test_host_logic.py::test_amp_mma_index_math)

Formulation (per warp: 16 channels, time in blocks of 8 steps; "s-block" m = 2x-rate samples
16m .. 16m+15, "z-tile" m = output steps 8m+3 .. 8m+10):
    U^T[ch, j]  = X^T[ch, k] . Gup[k, j]     k = 16 x-rows starting at 8m-3, j = 16 s-times of block m
    S           = snake(U)                    (registers; D fragment == A fragment of the next MMA)
    Z^T[ch, n]  = S^T[ch, kk] . Fdn[kk, n]   kk = 32 s-times of blocks m, m+1; n = 8 output steps
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root
sys.path.insert(0, ROOT)

LANES = np.arange(32)


# ---- Toeplitz fragment coefficient formulas (the CUDA kernel evaluates exactly these) ------------
def up_coeff(gu, k, j):
    """Gup[k][j]: weight of x-row k (relative to row 8m-3) in s-time j (relative to 16m); gu = 2*f."""
    if j % 2 == 0:
        mm = j // 2 + 5 - k
        return gu[2 * mm + 1] if 0 <= mm <= 5 else 0.0
    mm = (j - 1) // 2 + 6 - k
    return gu[2 * mm] if 0 <= mm <= 5 else 0.0


def down_coeff(fd, kk, n):
    """Fdn[kk][n]: weight of s-time kk (relative to 16m, kk in [0, 32)) in output step 8m+3+n."""
    t = kk - 2 * n - 1
    return fd[t] if 0 <= t <= 11 else 0.0


# ---- warp-level primitives -------------------------------------------------------------------------
def b_fragment(fn, n0, k0):
    """B fragment of mma.m16n8k16 (16 x 8, "col"): reg0 = (k = 2*(T%4)+{0,1}, n = T/4), reg1 = k+8."""
    frag = np.zeros((32, 2, 2))
    for T in LANES:
        for r in range(2):
            for e in range(2):
                frag[T, r, e] = fn(k0 + 8 * r + 2 * (T % 4) + e, n0 + T // 4)
    return frag


def ldmatrix_x4_trans(smem, row0, col0):
    """A fragment (16 ch x 16 times) from smem[time][ch]: matrix j covers times row0+8*(j/2).. and
    channels col0+8*(j%2)..; .trans hands thread T the elements stored[2*(T%4)+{0,1}][T/4]."""
    frag = np.zeros((32, 4, 2))
    for T in LANES:
        for j in range(4):
            for e in range(2):
                frag[T, j, e] = smem[row0 + 8 * (j // 2) + 2 * (T % 4) + e, col0 + 8 * (j % 2) + T // 4]
    return frag


def mma(a, b, d):
    """d[T][4] += A(16x16) . B(16x8) with the PTX fragment layouts."""
    A = np.zeros((16, 16))
    B = np.zeros((16, 8))
    for T in LANES:
        r, c = T // 4, 2 * (T % 4)
        for e in range(2):
            A[r, c + e] = a[T, 0, e]
            A[r + 8, c + e] = a[T, 1, e]
            A[r, c + 8 + e] = a[T, 2, e]
            A[r + 8, c + 8 + e] = a[T, 3, e]
            B[c + e, T // 4] = b[T, 0, e]
            B[c + 8 + e, T // 4] = b[T, 1, e]
    D = A @ B
    out = d.copy()
    for T in LANES:
        r, c = T // 4, 2 * (T % 4)
        out[T] += [D[r, c], D[r, c + 1], D[r + 8, c], D[r + 8, c + 1]]
    return out


def d_to_a(d0, d1):
    """Two D tiles (16 x 8 each, s-times 0-7 and 8-15) -> one A fragment (16 x 16)."""
    a = np.zeros((32, 4, 2))
    a[:, 0] = d0[:, 0:2]
    a[:, 1] = d0[:, 2:4]
    a[:, 2] = d1[:, 0:2]
    a[:, 3] = d1[:, 2:4]
    return a


def activation1d_mma(x, a_par, invb, f_up, f_down, nb=4):
    """x [L, C] (C % 16 == 0) -> z [L, C], emulating one CTA column of warps sliding over time."""
    L, C = x.shape
    gu = 2.0 * np.asarray(f_up, dtype=np.float64)
    fd = np.asarray(f_down, dtype=np.float64)
    z = np.full((L, C), np.nan)
    jl = 2 * L - 1
    m_last = (L - 4) // 8 if L >= 4 else -1
    up_b = [b_fragment(lambda k, j: up_coeff(gu, k, j), 8 * h, 0) for h in range(2)]
    dn_b = [b_fragment(lambda kk, n: down_coeff(fd, kk, n), 0, 16 * blk) for blk in range(2)]

    for g in range(C // 16):
        ch_a = 16 * g + LANES // 4  # channel of regs 0,1; regs 2,3 are ch_a + 8
        for m0 in range(-1, m_last + 1, nb):  # a warp's chunk of nb z-tiles
            r_origin = 8 * m0 - 3             # time index of smem row 0
            rows = 8 * nb + 16
            smem = x[np.clip(r_origin + np.arange(rows), 0, L - 1)]  # replicate clamp on x at fill time
            s_last = np.zeros((32, 2))

            def s_block(m):
                nonlocal s_last
                mm = max(m, 0)                # block -1 is rebuilt from block 0 (left clamp)
                xa = ldmatrix_x4_trans(smem, 8 * (mm - m0), 16 * g)
                d = [mma(xa, up_b[h], np.zeros((32, 4))) for h in range(2)]
                for h in range(2):            # snake; regs 0,1 -> channel ch_a, regs 2,3 -> ch_a + 8
                    for r in range(4):
                        ch = ch_a + (8 if r >= 2 else 0)
                        u = d[h][:, r]
                        d[h][:, r] = u + invb[ch] * np.sin(a_par[ch] * u) ** 2
                if m < 0:                     # s[j < 0] = s[0]: tile 0, column 0 -> lanes T%4 == 0, regs 0 / 2
                    src = LANES & ~3
                    va, vb = d[0][src, 0], d[0][src, 2]
                    for h in range(2):
                        d[h][:, 0] = d[h][:, 1] = va
                        d[h][:, 2] = d[h][:, 3] = vb
                elif 16 * m + 15 >= jl:       # right clamp: s[j > 2L-1] = s[2L-1]
                    if 16 * m <= jl:          # this block holds s[2L-1] (an odd column): keep it for later blocks
                        jj = jl - 16 * m
                        hh, nn = jj // 8, jj % 8
                        src = (LANES & ~3) | (nn // 2)
                        s_last = np.stack([d[hh][src, 1], d[hh][src, 3]], axis=1)
                    for h in range(2):
                        for e in range(2):
                            j = 16 * m + 8 * h + 2 * (LANES % 4) + e
                            d[h][:, e] = np.where(j > jl, s_last[:, 0], d[h][:, e])
                            d[h][:, 2 + e] = np.where(j > jl, s_last[:, 1], d[h][:, 2 + e])
                return d_to_a(d[0], d[1])

            prev = s_block(m0)
            for m in range(m0, min(m0 + nb, m_last + 1)):
                cur = s_block(m + 1)
                zt = mma(prev, dn_b[0], np.zeros((32, 4)))
                zt = mma(cur, dn_b[1], zt)
                for T in LANES:               # D fragment: (ch = T/4 [+8], t = 8m+3 + 2*(T%4) + {0,1})
                    for e in range(2):
                        t = 8 * m + 3 + 2 * (T % 4) + e
                        if 0 <= t < L:
                            z[t, 16 * g + T // 4] = zt[T, e]
                            z[t, 16 * g + T // 4 + 8] = zt[T, 2 + e]
                prev = cur
    return z


def main(lengths=(1, 2, 3, 4, 5, 6, 7, 8, 9, 11, 12, 13, 16, 19, 20, 31, 32, 33, 40, 67, 100), verbose=True):
    from oracle import bigvgan_oracle as O
    from svc_inference_pipeline_b200.utils import synth

    f = synth.aa_filter_taps().astype(np.float64)
    rng = np.random.default_rng(0)
    worst = 0.0
    for L in lengths:
        C = 16
        x = rng.standard_normal((1, C, L)) * 1.5
        alpha = rng.standard_normal(C) * 0.3
        beta = rng.standard_normal(C) * 0.3
        ref = O.activation1d(x, alpha, beta, True, f, f)[0].T  # [L, C]
        got = activation1d_mma(x[0].T.copy(), np.exp(alpha), 1.0 / (np.exp(beta) + 1e-9), f, f, nb=3)
        assert not np.isnan(got).any(), f"L={L}: outputs not written"
        err = np.abs(got - ref).max()
        worst = max(worst, err)
        if verbose:
            print(f"L={L:4d} max|err|={err:.2e}")
        assert err < 1e-12, f"L={L}: {err}"
    if verbose:
        print("ok, worst", worst)
    return worst


if __name__ == "__main__":
    main()
