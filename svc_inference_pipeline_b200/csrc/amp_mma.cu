// K-A (tensor-core variant): fused anti-aliased activation with both FIRs on the tensor cores.
//
// Replaces reference modules/bigvgan.py:251-256 (UpSample1d :278-287 -> Snake/SnakeBeta
// :84-95 / :146-159 -> DownSample1d :304-307 -> LowPassFilter1d :224-231), same math as
// amp_kernel.cu.  ncu on that kernel (profiles/r01_ncu_summary_v5.md) shows it is bound by
// instruction issue (85 % issue-active, 60 thread-instructions per element, of which the two
// 12-tap FIRs are 36 FFMA), not by HBM.  Here the FIRs are banded-Toeplitz GEMMs on
// mma.sync.m16n8k16 (bf16 operands, fp32 accumulate), in the transposed orientation that lets the
// accumulator fragment of the upsampler feed the downsampler without leaving registers:
//     U^T[ch, j]  = X^T[ch, k] . Gup[k, j]    k: 16 x rows from 8m-3,  j: the 16 2x-rate samples of block m
//     S           = snake(U)                   D fragment of the up-MMA == A fragment of the down-MMA
//     Z^T[ch, n]  = S^T[ch, kk] . Fdn[kk, n]  kk: the 32 samples of blocks m, m+1;  n: outputs 8m+3 .. 8m+10
// A warp owns 16 channels and slides over time in blocks of 8 steps, keeping the previous s-block
// as an A fragment.  Precision: operands are split into bf16 (hi, lo) pairs where the format needs
// it -- taps always (hi + lo = 16 mantissa bits), x when it arrives as fp32, s when the output is a
// SPLIT tensor -- and the cross terms hi*hi + hi*lo + lo*hi are accumulated in fp32, i.e. the
// same 2^-16 relative operand precision the tensor-core convolutions consume.
//
// Data movement: a CTA stages its [time tile + 16 halo rows] x [16 NG channels] of x in shared
// memory as bf16 planes (coalesced 16-byte global loads, replicate clamp applied while staging),
// ldmatrix.trans builds the channel-major A fragments, stmatrix.trans transposes the result
// back and every global store is a 16-byte piece of a channels-last row.
// Index arithmetic is pinned by the lane-level emulation tests/amp_mma_emulation.py.
#include "amp_mma.cuh"

#ifndef BVG_AMP_PAIR
#define BVG_AMP_PAIR 1  // interior tiles: two time blocks in flight per warp (0 = one at a time, for A/B builds)
#endif

namespace bvg {

// IN_BF16: x is bf16 (one operand plane) else fp32 (hi, lo planes).  OUT_MODE: BVG_BF16 | BVG_SPLIT.
// The activated signal s always enters the down-MMA as hi + lo (rounding it to one bf16 costs the
// bf16 path 0.5 dB of SNR and pushes its log-mel L1 over the 1e-2 gate).  NG: 16-channel groups per
// CTA = warps per CTA; a warp owns 16 channels for the whole walk along time.
template <bool IN_BF16, int OUT_MODE, bool FAST_SIN, int NG>
__global__ void __launch_bounds__(32 * NG) amp_mma_kernel(const __grid_constant__ AmpMmaParams p) {
  constexpr int NT = 32 * NG;
  constexpr int CT = 16 * NG;
  constexpr int PITCH = CT * 2 + 16;  // bytes per staged row: +16 keeps the 8 rows of an ldmatrix on distinct banks
  constexpr int NPL = IN_BF16 ? 1 : 2;
  constexpr int NOUT = OUT_MODE == BVG_SPLIT ? 2 : 1;
  // How the activated signal s enters the down-MMA.  SPLIT output (fp32 path): bf16 hi + lo, three k16
  // MMAs per s-block.  BF16 output: one tf32 term (cvt.rna, 11 significant bits: 4x finer than the bf16
  // the result is rounded to, and no exponent-range concern), k8 MMAs straight from the fp32 registers.
  //   S_F16 (BF16 output): s and the low-pass taps as single fp16 terms -- 11 significant bits, 8x finer than the
  //   bf16 the result is rounded to (a single *bf16* s term costs the bf16 path 0.5 dB and its log-mel gate);
  //   |s| above 65504 saturates.  2 HMMAs per z-tile instead of 6, no split arithmetic.
  constexpr bool S_F16 = OUT_MODE == BVG_BF16;
  constexpr bool S_SPLIT = true;  // (else-branch below: tf32 k8 MMAs, measured slower on B200 -- kept for reference)
  constexpr int PLANE = AM_ROWS * PITCH;
  constexpr int STG_PLANE = 8 * 48;   // 8 rows x 16 channels, 48-byte pitch (conflict-free stmatrix)
  constexpr int STG = STG_PLANE * NOUT;
  // bf16 input: two tile buffers filled by cp.async (no staging registers); fp32 input: the (hi, lo) planes of
  // one tile, written from registers after the split.  Then two output staging buffers per warp.
  constexpr int XBYTES = 2 * PLANE;
  __shared__ __align__(16) uint8_t smem[XBYTES + NG * 2 * STG];

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int g = tid >> 5;
  pdl_trigger();  // programmatic dependent launch (common.cuh): no global access before the wait
  pdl_wait();

  const int cgi = blockIdx.x % p.n_cg;
  const int rest = blockIdx.x / p.n_cg;
  const int cti = rest % p.n_ct;
  const int b = rest / p.n_ct;
  const int c0 = cgi * CT;
  const int L = p.L, C = p.C;
  const long long item = (long long)b * L * C;
  const int jl = 2 * L - 1;  // last valid 2x-rate sample

  // ---- per-thread constants: Toeplitz B fragments (hi, lo), snake parameters, addresses ----------------
  const int q = lane & 3, rw = lane >> 2;
  // dn_*: [s-block (prev, cur)][reg]; S_SPLIT: bf16 k16 fragments (reg r = k rows 8r + 2q, +1), else tf32
  // k8 fragments per 8-sample half h: regs 2h, 2h+1 = samples 8h + 2q, 8h + 2q + 1 (the K order of each
  // k8 MMA is permuted so that the D fragment of the up-MMA is its A fragment as it stands)
  uint32_t up_hi[2][2], up_lo[2][2], dn_hi[2][S_SPLIT ? 2 : 4], dn_lo[2][S_SPLIT ? 2 : 4];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int k = 8 * r + 2 * q;
      amm::split_pair(amm::up_coeff(p, k, 8 * h + rw), amm::up_coeff(p, k + 1, 8 * h + rw), up_hi[h][r], up_lo[h][r]);
      if constexpr (S_F16) {
        dn_hi[h][r] = amm::pack_f16x2_sat(amm::down_coeff(p, 16 * h + k, rw), amm::down_coeff(p, 16 * h + k + 1, rw));
        dn_lo[h][r] = 0u;
      } else if constexpr (S_SPLIT) {
        amm::split_pair(amm::down_coeff(p, 16 * h + k, rw), amm::down_coeff(p, 16 * h + k + 1, rw), dn_hi[h][r], dn_lo[h][r]);
      } else {
#pragma unroll
        for (int e = 0; e < 2; ++e) {  // block h, half r, sample 2q + e
          const float c = amm::down_coeff(p, 16 * h + k + e, rw);
          dn_hi[h][2 * r + e] = amm::to_tf32(c);
          dn_lo[h][2 * r + e] = __float_as_uint(c - __uint_as_float(dn_hi[h][2 * r + e]));
        }
      }
    }
  const int ch_a = c0 + 16 * g + rw, ch_b = ch_a + 8;
  float apar[2], invb[2];
  {
    const float a0 = ch_a < C ? __ldg(p.a + ch_a) : 0.f, a1 = ch_b < C ? __ldg(p.a + ch_b) : 0.f;
    apar[0] = FAST_SIN ? a0 : a0 * 0.318309886183790672f;
    apar[1] = FAST_SIN ? a1 : a1 * 0.318309886183790672f;
    invb[0] = ch_a < C ? __ldg(p.invb + ch_a) : 0.f;
    invb[1] = ch_b < C ? __ldg(p.invb + ch_b) : 0.f;
  }
  // ldmatrix row address of this lane: matrix j = lane / 8 -> times +8 * (j / 2), channels +8 * (j % 2)
  const uint32_t ld_base = amm::smem_u32(smem) + (uint32_t)(((lane & 7) + 8 * (lane >> 4)) * PITCH + 32 * g + 16 * ((lane >> 3) & 1));
  uint8_t* const stg = smem + XBYTES + g * 2 * STG;
  // stmatrix row address: matrix j = lane / 8 -> plane j / 2, channels +8 * (j % 2); row lane % 8
  const uint32_t st_addr = amm::smem_u32(stg) + (uint32_t)((lane >> 4) * STG_PLANE + (lane & 7) * 48 + ((lane >> 3) & 1) * 16);
  // read-back: lane -> plane lane / 16, row (lane % 16) / 2, channel half lane % 2
  const int rb_pl = lane >> 4, rb_row = (lane & 15) >> 1, rb_half = lane & 1;
  const uint8_t* const rb_ptr = stg + rb_pl * STG_PLANE + rb_row * 48 + rb_half * 16;
  const int rb_ch = c0 + 16 * g + 8 * rb_half;
  const bool rb_on = rb_pl < NOUT && rb_ch < C;
  uint16_t* const out_base = reinterpret_cast<uint16_t*>(rb_pl == 0 ? p.y : p.y_lo) + item + rb_ch;

  float s_last[2] = {0.f, 0.f};
  struct SFrag {
    uint32_t hi[S_SPLIT ? 4 : 8];  // S_SPLIT: bf16 pairs (k16 A fragment); else tf32 values d[h][i] at 4h + i
    uint32_t lo[S_SPLIT ? 4 : 1];
  };

  // s-block m (2x-rate samples 16m .. 16m+15) from staged rows starting at `addr`:
  // upsample (MMA) -> snake -> [EDGE: replicate clamps of the activated signal] -> A fragments
  auto s_block = [&](auto edge_tag, int m, uint32_t addr, SFrag& out) {
    constexpr bool EDGE = decltype(edge_tag)::value;
    [[maybe_unused]] uint32_t xh[4], xl[4];
    amm::ldmatrix_x4_trans(addr, xh);
    if constexpr (NPL == 2) amm::ldmatrix_x4_trans(addr + PLANE, xl);
    float d[2][4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
      for (int i = 0; i < 4; ++i) d[h][i] = 0.f;
      amm::mma_bf16(d[h], xh, up_hi[h]);
      amm::mma_bf16(d[h], xh, up_lo[h]);
      if constexpr (NPL == 2) amm::mma_bf16(d[h], xl, up_hi[h]);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {  // registers (0, 1) and (2, 3) are two time steps of one channel each
      amm::snake_pair<FAST_SIN>(d[h][0], d[h][1], apar[0], invb[0]);
      amm::snake_pair<FAST_SIN>(d[h][2], d[h][3], apar[1], invb[1]);
    }
    if constexpr (EDGE) {
      if (m < 0) {
        // left clamp (LowPassFilter1d pad, bigvgan.py:227): s[j < 0] = s[0]; `addr` pointed at block 0
        const float va = __shfl_sync(0xffffffffu, d[0][0], lane & ~3), vb = __shfl_sync(0xffffffffu, d[0][2], lane & ~3);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          d[h][0] = d[h][1] = va;
          d[h][2] = d[h][3] = vb;
        }
      } else if (16 * m + 15 >= jl) {
        // right clamp: s[j > 2L-1] = s[2L-1] (an odd column of this block, or kept from an earlier one)
        if (16 * m <= jl) {
          const int jj = jl - 16 * m;
          const int src = (lane & ~3) | ((jj & 7) >> 1);
          const float ta = (jj >> 3) ? d[1][1] : d[0][1], tb = (jj >> 3) ? d[1][3] : d[0][3];
          s_last[0] = __shfl_sync(0xffffffffu, ta, src);
          s_last[1] = __shfl_sync(0xffffffffu, tb, src);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            if (16 * m + 8 * h + 2 * q + e > jl) {
              d[h][e] = s_last[0];
              d[h][2 + e] = s_last[1];
            }
          }
      }
    }
    if constexpr (S_F16) {
      out.hi[0] = amm::pack_f16x2_sat(d[0][0], d[0][1]);
      out.hi[1] = amm::pack_f16x2_sat(d[0][2], d[0][3]);
      out.hi[2] = amm::pack_f16x2_sat(d[1][0], d[1][1]);
      out.hi[3] = amm::pack_f16x2_sat(d[1][2], d[1][3]);
    } else if constexpr (S_SPLIT) {
      amm::split_pair(d[0][0], d[0][1], out.hi[0], out.lo[0]);
      amm::split_pair(d[0][2], d[0][3], out.hi[1], out.lo[1]);
      amm::split_pair(d[1][0], d[1][1], out.hi[2], out.lo[2]);
      amm::split_pair(d[1][2], d[1][3], out.hi[3], out.lo[3]);
    } else {
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int i = 0; i < 4; ++i) out.hi[4 * h + i] = amm::to_tf32(d[h][i]);
    }
  };
  // z (+)= s-block `f` (blk 0: the older one) through the low-pass Toeplitz fragments
  auto down = [&](float (&z)[4], const SFrag& f, int blk) {
    if constexpr (S_F16) {
      amm::mma_f16(z, f.hi, dn_hi[blk]);
    } else if constexpr (S_SPLIT) {
      amm::mma_bf16(z, f.hi, dn_hi[blk]);
      amm::mma_bf16(z, f.hi, dn_lo[blk]);
      amm::mma_bf16(z, f.lo, dn_hi[blk]);
    } else {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        // k8 slots (q, q + 4) <-> samples (2q, 2q + 1) of half h; rows (rw, rw + 8) = regs (0|1, 2|3)
        amm::mma_tf32(z, f.hi[4 * h + 0], f.hi[4 * h + 2], f.hi[4 * h + 1], f.hi[4 * h + 3], dn_hi[blk][2 * h], dn_hi[blk][2 * h + 1]);
        amm::mma_tf32(z, f.hi[4 * h + 0], f.hi[4 * h + 2], f.hi[4 * h + 1], f.hi[4 * h + 3], dn_lo[blk][2 * h], dn_lo[blk][2 * h + 1]);
      }
    }
  };

  SFrag fa, fb;  // ping-pong: fa holds the s-block a staged tile starts from

  // one staged tile: z-tiles mt .. mt + AM_NB - 1 (EDGE: clamps, partial last tile, row predicates)
  auto run_tile = [&](auto edge_tag, int mt, uint32_t xoff) {
    constexpr bool EDGE = decltype(edge_tag)::value;
    uint32_t addr = ld_base + xoff + 8 * PITCH;                        // s-block mt + 1 starts at staged row 8
    // z-tile m from s-blocks m (prev) and m + 1 (cur, computed here)
    auto step = [&](int i, const SFrag& prev, SFrag& cur) -> bool {
      const int m = mt + i;
      if (EDGE && m > p.m_last) return false;
      s_block(edge_tag, m + 1, addr, cur);
      float z[4] = {0.f, 0.f, 0.f, 0.f}, z2[4] = {0.f, 0.f, 0.f, 0.f};  // two accumulation chains (MMA latency)
      down(z, prev, 0);
      down(z2, cur, 1);
#pragma unroll
      for (int r = 0; r < 4; ++r) z[r] += z2[r];
      // z fragment: (channel rw [+8], step 8m+3 + 2q + {0,1}) -> transposed through the warp's staging rows
      const uint32_t sbuf = (uint32_t)((i & 1) * STG);
      if constexpr (NOUT == 2) {
        uint32_t h0, l0, h1, l1;
        amm::split_pair(z[0], z[1], h0, l0);
        amm::split_pair(z[2], z[3], h1, l1);
        amm::stmatrix_x4_trans(st_addr + sbuf, h0, h1, l0, l1);
      } else {
        amm::stmatrix_x2_trans(st_addr + sbuf, pack_bf16x2(z[0], z[1]), pack_bf16x2(z[2], z[3]));
      }
      __syncwarp();
      bool on = rb_on;
      if constexpr (EDGE) {
        const int t = 8 * m + 3 + rb_row;
        on = on && t >= 0 && t < L;
      }
      // (address formed from the row index each time: a pointer carried across the nested lambdas ends up on the stack)
      if (on) *reinterpret_cast<uint4*>(out_base + (long long)(8 * m + 3 + rb_row) * C) = *reinterpret_cast<const uint4*>(rb_ptr + sbuf);
      addr += 8 * PITCH;
      return true;
    };
    static_assert(AM_NB % 2 == 0, "ping-pong needs an even number of z-tiles per staged tile");
#if BVG_AMP_PAIR
    if constexpr (!EDGE) {
      // Interior tiles: two z-tiles per iteration with their two s-blocks computed side by side -- two independent
      // ldmatrix -> HMMA -> snake -> HMMA chains in flight per warp instead of one (the stalls ncu showed were
      // dependent-issue waits on exactly that chain).  Same arithmetic in the same order per element: bit-identical.
#pragma unroll 1
      for (int i = 0; i < AM_NB; i += 2) {
        const int m = mt + i;
        SFrag c0, c1;
        {
          [[maybe_unused]] uint32_t xh0[4], xh1[4], xl0[4], xl1[4];
          amm::ldmatrix_x4_trans(addr, xh0);
          amm::ldmatrix_x4_trans(addr + 8 * PITCH, xh1);
          if constexpr (NPL == 2) {
            amm::ldmatrix_x4_trans(addr + PLANE, xl0);
            amm::ldmatrix_x4_trans(addr + 8 * PITCH + PLANE, xl1);
          }
          float d0[2][4], d1[2][4];
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int r = 0; r < 4; ++r) d0[h][r] = d1[h][r] = 0.f;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            amm::mma_bf16(d0[h], xh0, up_hi[h]);
            amm::mma_bf16(d1[h], xh1, up_hi[h]);
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            amm::mma_bf16(d0[h], xh0, up_lo[h]);
            amm::mma_bf16(d1[h], xh1, up_lo[h]);
          }
          if constexpr (NPL == 2) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              amm::mma_bf16(d0[h], xl0, up_hi[h]);
              amm::mma_bf16(d1[h], xl1, up_hi[h]);
            }
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            amm::snake_pair<FAST_SIN>(d0[h][0], d0[h][1], apar[0], invb[0]);
            amm::snake_pair<FAST_SIN>(d1[h][0], d1[h][1], apar[0], invb[0]);
            amm::snake_pair<FAST_SIN>(d0[h][2], d0[h][3], apar[1], invb[1]);
            amm::snake_pair<FAST_SIN>(d1[h][2], d1[h][3], apar[1], invb[1]);
          }
          auto pack = [&](const float (&d)[2][4], SFrag& out) {
            if constexpr (S_F16) {
              out.hi[0] = amm::pack_f16x2_sat(d[0][0], d[0][1]);
              out.hi[1] = amm::pack_f16x2_sat(d[0][2], d[0][3]);
              out.hi[2] = amm::pack_f16x2_sat(d[1][0], d[1][1]);
              out.hi[3] = amm::pack_f16x2_sat(d[1][2], d[1][3]);
            } else {
              amm::split_pair(d[0][0], d[0][1], out.hi[0], out.lo[0]);
              amm::split_pair(d[0][2], d[0][3], out.hi[1], out.lo[1]);
              amm::split_pair(d[1][0], d[1][1], out.hi[2], out.lo[2]);
              amm::split_pair(d[1][2], d[1][3], out.hi[3], out.lo[3]);
            }
          };
          pack(d0, c0);
          pack(d1, c1);
        }
        float za[4] = {0.f, 0.f, 0.f, 0.f}, za2[4] = {0.f, 0.f, 0.f, 0.f}, zb[4] = {0.f, 0.f, 0.f, 0.f}, zb2[4] = {0.f, 0.f, 0.f, 0.f};
        down(za, fa, 0);
        down(zb, c0, 0);
        down(za2, c0, 1);
        down(zb2, c1, 1);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          za[r] += za2[r];
          zb[r] += zb2[r];
        }
        if constexpr (NOUT == 2) {
          uint32_t h0, l0, h1, l1;
          amm::split_pair(za[0], za[1], h0, l0);
          amm::split_pair(za[2], za[3], h1, l1);
          amm::stmatrix_x4_trans(st_addr, h0, h1, l0, l1);
          amm::split_pair(zb[0], zb[1], h0, l0);
          amm::split_pair(zb[2], zb[3], h1, l1);
          amm::stmatrix_x4_trans(st_addr + STG, h0, h1, l0, l1);
        } else {
          amm::stmatrix_x2_trans(st_addr, pack_bf16x2(za[0], za[1]), pack_bf16x2(za[2], za[3]));
          amm::stmatrix_x2_trans(st_addr + STG, pack_bf16x2(zb[0], zb[1]), pack_bf16x2(zb[2], zb[3]));
        }
        __syncwarp();
        if (rb_on) {
          const uint4 va = *reinterpret_cast<const uint4*>(rb_ptr), vb = *reinterpret_cast<const uint4*>(rb_ptr + STG);
          *reinterpret_cast<uint4*>(out_base + (long long)(8 * m + 3 + rb_row) * C) = va;
          *reinterpret_cast<uint4*>(out_base + (long long)(8 * m + 11 + rb_row) * C) = vb;
        }
        __syncwarp();  // the staging rows are rewritten by the next iteration
        fa = c1;
        addr += 16 * PITCH;
      }
    } else
#endif
    {
#pragma unroll 1
      for (int i = 0; i < AM_NB; i += 2) {
        if (!step(i, fa, fb)) break;
        if (!step(i + 1, fb, fa)) break;
      }
    }
  };

  // ---- staging: global (coalesced 16-byte loads) -> registers -> bf16 planes in shared memory.  The loads
  // of tile k+1 are issued before tile k is computed and parked in registers, so their latency hides
  // behind the MMAs; the replicate clamp on x (UpSample1d pad, bigvgan.py:281) is applied to the row index.
  constexpr int VEC = IN_BF16 ? 8 : 4;
  constexpr int VPR = CT / VEC;  // vectors per row
  constexpr int NV = AM_ROWS * VPR / NT;
  static_assert(AM_ROWS * VPR % NT == 0, "staging assumes a whole number of vectors per thread");
  uint4 v[NV];
  auto load_tile = [&](int tile) {
    const int R0 = 8 * (-1 + tile * AM_NB) - 3;  // staged row r holds time R0 + r
#pragma unroll
    for (int it = 0; it < NV; ++it) {
      const int idx = tid + it * NT;
      const int r = idx / VPR, cv = idx - r * VPR;
      const int t = min(max(R0 + r, 0), L - 1);
      const int ch = c0 + cv * VEC;
      v[it] = make_uint4(0u, 0u, 0u, 0u);
      if (ch < C) {
        if constexpr (IN_BF16)
          v[it] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.x) + item + (long long)t * C + ch));
        else
          v[it] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(p.x) + item + (long long)t * C + ch));
      }
    }
  };
  auto store_tile = [&]() {
#pragma unroll
    for (int it = 0; it < NV; ++it) {
      const int idx = tid + it * NT;
      const int r = idx / VPR, cv = idx - r * VPR;
      if constexpr (IN_BF16) {
        *reinterpret_cast<uint4*>(smem + r * PITCH + cv * 16) = v[it];
      } else {
        uint2 hi, lo;
        amm::split_pair(__uint_as_float(v[it].x), __uint_as_float(v[it].y), hi.x, lo.x);
        amm::split_pair(__uint_as_float(v[it].z), __uint_as_float(v[it].w), hi.y, lo.y);
        *reinterpret_cast<uint2*>(smem + r * PITCH + cv * 8) = hi;
        *reinterpret_cast<uint2*>(smem + PLANE + r * PITCH + cv * 8) = lo;
      }
    }
  };

  // bf16 input: 16-byte cp.async straight into the tile buffer (replicate clamp = clamped source row,
  // channels past C = zero-fill), tile k+1 in flight while tile k is computed
  auto async_tile = [&](int tile, int buf) {
    const int R0 = 8 * (-1 + tile * AM_NB) - 3;
    const uint32_t dst0 = amm::smem_u32(smem) + (uint32_t)(buf * PLANE);
#pragma unroll
    for (int it = 0; it < NV; ++it) {
      const int idx = tid + it * NT;
      const int r = idx / VPR, cv = idx - r * VPR;
      const int t = min(max(R0 + r, 0), L - 1);
      const int ch = c0 + cv * VEC;
      const uint16_t* src = reinterpret_cast<const uint16_t*>(p.x) + item + (long long)t * C + (ch < C ? ch : 0);
      const uint32_t nbytes = ch < C ? 16u : 0u;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst0 + (uint32_t)(r * PITCH + cv * 16)), "l"(src), "r"(nbytes) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  const int tile0 = cti * p.tiles_per_cta;
  const int tile_end = min(tile0 + p.tiles_per_cta, p.n_tiles);
  if constexpr (IN_BF16) {
    async_tile(tile0, 0);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else {
    load_tile(tile0);
    store_tile();
  }
  __syncthreads();
  int buf = 0;
  for (int tile = tile0; tile < tile_end; ++tile) {
    const int mt = -1 + tile * AM_NB;  // first z-tile of the staged tile
    const bool has_next = tile + 1 < tile_end;
    const uint32_t xoff = IN_BF16 ? (uint32_t)(buf * PLANE) : 0u;
    if (has_next) {
      if constexpr (IN_BF16)
        async_tile(tile + 1, buf ^ 1);
      else
        load_tile(tile + 1);
    }
    if (tile == tile0) {
      // warm-up: s-block mt (rows from 0), or block 0 (rows from 8) broadcast as block -1
      s_block(std::true_type{}, mt, ld_base + xoff + (mt < 0 ? 8 * PITCH : 0), fa);
    }
    const bool edge = mt < 0 || 16 * (mt + AM_NB) + 15 >= jl;
    if (edge)
      run_tile(std::true_type{}, mt, xoff);
    else
      run_tile(std::false_type{}, mt, xoff);
    if (!has_next) break;
    if constexpr (IN_BF16) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();  // the next tile has landed for everyone, and every warp is done with this one
      buf ^= 1;
    } else {
      __syncthreads();  // every warp is done with this tile's rows
      store_tile();
      __syncthreads();
    }
  }
}

template <bool IN_BF16, int OUT_MODE, bool FAST_SIN, int NG>
static cudaError_t launch_amp_mma(const AmpMmaParams& p, cudaStream_t st) {
  const long long blocks = (long long)p.B * p.n_ct * p.n_cg;
  // same shared-memory carve-out as the convolution kernels (max shared): an SM hosts kernels of two streams at once
  // only when they agree on the L1 / shared split
  if (first_use_on_device(reinterpret_cast<const void*>(amp_mma_kernel<IN_BF16, OUT_MODE, FAST_SIN, NG>)))
    cudaFuncSetAttribute(amp_mma_kernel<IN_BF16, OUT_MODE, FAST_SIN, NG>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  return launch_k(amp_mma_kernel<IN_BF16, OUT_MODE, FAST_SIN, NG>, dim3((unsigned)blocks), dim3(32 * NG), 0, st, p);
}

template <bool IN_BF16, int OUT_MODE, bool FAST_SIN>
static cudaError_t launch_amp_mma_ng(const AmpMmaParams& p, int ng, cudaStream_t st) {
  if (ng == 4) return launch_amp_mma<IN_BF16, OUT_MODE, FAST_SIN, 4>(p, st);
  if (ng == 3) return launch_amp_mma<IN_BF16, OUT_MODE, FAST_SIN, 3>(p, st);
  return launch_amp_mma<IN_BF16, OUT_MODE, FAST_SIN, 2>(p, st);
}

// bvg_tuning.amp_mma: 0 = never, 1 (default) = where it measured faster than the FFMA2 kernel on B200
// (profiles/r01_ncu_summary_v7.md section 4: BF16 -> BF16, every C that is a multiple of 8), 2 = wherever it is supported;
// bvg_tuning.amp_mma_tiles: time tiles per CTA, 0 = choose

// The tensor-core kernel takes F32 -> SPLIT (fp32 path) and BF16 -> BF16 (bf16 path) with C a multiple
// of 8.  Everything else stays on amp_kernel.cu.
bool amp_mma_supported(const bvg_amp_desc* d) {
  const int amp_mma_enable = tune_of(d->tune).amp_mma;
  if (!amp_mma_enable) return false;
  if (d->C % 8 != 0) return false;
  const bool f32_split = d->x.dtype == BVG_F32 && d->y.dtype == BVG_SPLIT;
  const bool bf_bf = d->x.dtype == BVG_BF16 && d->y.dtype == BVG_BF16;
  if (!f32_split && !bf_bf) return false;
  if (amp_mma_enable == 1 && !bf_bf) return false;
  if (((uintptr_t)d->x.d_ptr & 15) || ((uintptr_t)d->y.d_ptr & 15) || (d->y.d_lo && ((uintptr_t)d->y.d_lo & 15))) return false;
  return true;
}

int amp_mma_forward(const bvg_amp_desc* d, cudaStream_t st) {
  AmpMmaParams p;
  p.x = d->x.d_ptr;
  p.y = d->y.d_ptr;
  p.y_lo = d->y.d_lo;
  p.a = d->d_a;
  p.invb = d->d_invb;
  for (int k = 0; k < 12; ++k) {
    p.gu[k] = 2.0f * d->taps_up[k];
    p.fd[k] = d->taps_down[k];
  }
  p.B = d->B;
  p.L = d->L;
  p.C = d->C;
  // channel groups per CTA: least padding, then widest
  int ng = 4, best = 1 << 30;
  for (int cand = 4; cand >= 2; --cand) {
    const int padded = ceil_div(d->C, 16 * cand) * 16 * cand;
    if (padded < best) {
      best = padded;
      ng = cand;
    }
  }
  p.n_cg = ceil_div(d->C, 16 * ng);
  p.m_last = d->L >= 4 ? (d->L - 4) / 8 : -1;
  p.n_tiles = ceil_div(p.m_last + 2, AM_NB);
  // tiles per CTA: long walks amortise the per-CTA set-up (Toeplitz fragments, warm-up block) while
  // the grid still covers every SM with a few rounds of co-resident CTAs
  const int amp_mma_tiles = tune_of(d->tune).amp_mma_tiles;
  int tpc = amp_mma_tiles > 0 ? amp_mma_tiles : 32;
  while (tpc > 1 && (long long)d->B * p.n_cg * ceil_div(p.n_tiles, tpc) < 148ll * 5 * 3) tpc >>= 1;
  p.tiles_per_cta = tpc;
  p.n_ct = ceil_div(p.n_tiles, tpc);
  BVG_REQUIRE((long long)d->B * p.n_ct * p.n_cg < (1ll << 31), "amp: grid too large");
  cudaError_t e;
  const bool fast = d->fast_sin != 0;
  if (d->x.dtype == BVG_F32)
    e = fast ? launch_amp_mma_ng<false, BVG_SPLIT, true>(p, ng, st) : launch_amp_mma_ng<false, BVG_SPLIT, false>(p, ng, st);
  else
    e = fast ? launch_amp_mma_ng<true, BVG_BF16, true>(p, ng, st) : launch_amp_mma_ng<true, BVG_BF16, false>(p, ng, st);
  if (e != cudaSuccess) return cuda_fail(e, "amp_mma_kernel launch");
  return BVG_OK;
}

}  // namespace bvg
