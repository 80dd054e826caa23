// K-A, fp32 path (F32 in -> SPLIT out): tensor-core Activation1d as a per-warp stream.
//
// Same math and MMA formulation as amp_mma.cu (banded-Toeplitz FIRs on mma.sync.m16n8k16, bf16 (hi, lo)
// operand pairs, the accumulator fragment of the upsampler feeding the low-pass after the snake;
// replaces reference modules/bigvgan.py:251-256), different data movement:
//   * every warp owns 16 channels and walks along time on its own -- there is no block-wide barrier;
//   * x arrives as raw fp32 through cp.async (16 bytes per lane per 8-row group) into a per-warp ring of
//     64 rows, 4 row groups ahead of the MMAs; the replicate clamp on x (UpSample1d pad,
//     bigvgan.py:281) is the clamped source row, channels past C are zero-filled by the copy;
//   * the A fragments are built straight from the fp32 ring (8 conflict-free LDS.32 per s-block) and
//     split into (hi, lo) bf16 in registers, so no staging registers live across the loop.
// ncu on the staged variant (profiles/r01_ncu_summary_v7.md): 46 instructions per element but only 15
// resident warps (128 registers, 40 of them prefetch) and a barrier per 64 steps.
#include "amp_mma.cuh"

namespace bvg {

constexpr int RS_ROWS = 64;    // ring rows per warp (power of two)
constexpr int RS_D = 4;        // row groups (8 rows) in flight ahead of the one being consumed
constexpr int RS_WARPS = 4;

// BF16_PATH = false: F32 in -> SPLIT out (fp32 path; x, s, taps as bf16 (hi, lo) pairs).
// BF16_PATH = true : BF16 in -> BF16 out (bf16 path; x as it is, s and low-pass taps as single fp16 terms,
//                    see amp_mma.cu): ring rows are bf16, fragments come from ldmatrix.trans.
template <bool BF16_PATH, bool FAST_SIN>
__global__ void __launch_bounds__(32 * RS_WARPS) amp_stream_kernel(const __grid_constant__ AmpMmaParams p) {
  // ring row: 16 channels + 16 bytes of padding -- fp32: the 4 row pairs of a fragment load hit distinct banks;
  // bf16: the 8 rows of an ldmatrix do
  constexpr int RS_PITCH = BF16_PATH ? 48 : 80;
  constexpr int NOUT = BF16_PATH ? 1 : 2;
  constexpr int STG_PLANE = 8 * 48;  // 8 rows x 16 channels of bf16, 48-byte pitch (conflict-free stmatrix)
  constexpr int STG = NOUT * STG_PLANE;
  __shared__ __align__(16) uint8_t smem[RS_WARPS * (RS_ROWS * RS_PITCH + 2 * STG)];

  const int lane = threadIdx.x & 31;
  const int g = threadIdx.x >> 5;
  const int cgi = blockIdx.x % p.n_cg;
  const int rest = blockIdx.x / p.n_cg;
  const int cti = rest % p.n_ct;
  const int b = rest / p.n_ct;
  const int L = p.L, C = p.C;
  const int c_w = cgi * (16 * RS_WARPS) + 16 * g;  // first channel of this warp
  if (c_w >= C) return;                            // no block-wide synchronisation anywhere below
  const long long item = (long long)b * L * C;
  const int jl = 2 * L - 1;                        // last valid 2x-rate sample

  const int m_begin = -1 + cti * p.tiles_per_cta * AM_NB;  // z-tiles m_begin .. m_end
  const int m_end = min(m_begin + p.tiles_per_cta * AM_NB - 1, p.m_last);
  if (m_begin > p.m_last) return;
  const int R0 = 8 * m_begin - 3;                  // time of ring row index 0

  // ---- per-thread constants ------------------------------------------------------------------------
  const int q = lane & 3, rw = lane >> 2;
  uint32_t up_hi[2][2], up_lo[2][2], dn_hi[2][2], dn_lo[2][2];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int k = 8 * r + 2 * q;
      amm::split_pair(amm::up_coeff(p, k, 8 * h + rw), amm::up_coeff(p, k + 1, 8 * h + rw), up_hi[h][r], up_lo[h][r]);
      if constexpr (BF16_PATH) {
        dn_hi[h][r] = amm::pack_f16x2_sat(amm::down_coeff(p, 16 * h + k, rw), amm::down_coeff(p, 16 * h + k + 1, rw));
        dn_lo[h][r] = 0u;
      } else {
        amm::split_pair(amm::down_coeff(p, 16 * h + k, rw), amm::down_coeff(p, 16 * h + k + 1, rw), dn_hi[h][r], dn_lo[h][r]);
      }
    }
  const int ch_a = c_w + rw, ch_b = ch_a + 8;
  float apar[2], invb[2];
  {
    const float a0 = ch_a < C ? __ldg(p.a + ch_a) : 0.f, a1 = ch_b < C ? __ldg(p.a + ch_b) : 0.f;
    apar[0] = FAST_SIN ? a0 : a0 * 0.318309886183790672f;
    apar[1] = FAST_SIN ? a1 : a1 * 0.318309886183790672f;
    invb[0] = ch_a < C ? __ldg(p.invb + ch_a) : 0.f;
    invb[1] = ch_b < C ? __ldg(p.invb + ch_b) : 0.f;
  }
  uint8_t* const ring = smem + g * (RS_ROWS * RS_PITCH + 2 * STG);
  uint8_t* const stg = ring + RS_ROWS * RS_PITCH;
  const uint32_t ring_u32 = amm::smem_u32(ring);
  // fp32 ring: fragment element (row 2q + e [+8], channel rw [+8]) of a 16-row window starting at a multiple of 8;
  // bf16 ring: ldmatrix row of this lane (matrix j = lane / 8: rows +8 * (j / 2), channels +8 * (j % 2))
  const uint32_t frag_off = BF16_PATH ? (uint32_t)((lane & 7) * RS_PITCH + 16 * ((lane >> 3) & 1)) : (uint32_t)(2 * q * RS_PITCH + rw * 4);
  const int frag_row8 = BF16_PATH ? 8 * (lane >> 4) : 0;
  // cp.async role of this lane inside an 8-row group.  fp32: row lane / 4, 16-byte chunk lane % 4 (4 channels);
  // bf16: lanes 0-15 only, row lane / 2, chunk lane % 2 (8 channels)
  constexpr int CPR = BF16_PATH ? 2 : 4;   // 16-byte chunks per ring row
  constexpr int CPE = BF16_PATH ? 8 : 4;   // channels per chunk
  const bool cp_lane = lane < 8 * CPR;
  const int cp_row = (lane / CPR) & 7, cp_ch = c_w + CPE * (lane % CPR);
  const uint32_t cp_bytes = cp_ch < C ? 16u : 0u;
  const uint8_t* const cp_src = reinterpret_cast<const uint8_t*>(p.x) + (item + (cp_ch < C ? cp_ch : 0)) * (BF16_PATH ? 2 : 4);
  const long long cp_stride = (long long)C * (BF16_PATH ? 2 : 4);
  const uint32_t cp_dst = ring_u32 + (uint32_t)(cp_row * RS_PITCH + (lane % CPR) * 16);
  // output: stmatrix row address (matrix j = lane / 8 -> plane j / 2, channels +8 * (j % 2)), read-back role
  const uint32_t st_addr = amm::smem_u32(stg) + (uint32_t)((lane >> 4) * STG_PLANE + (lane & 7) * 48 + ((lane >> 3) & 1) * 16);
  const int rb_pl = lane >> 4, rb_row = (lane & 15) >> 1, rb_half = lane & 1;
  const uint8_t* const rb_ptr = stg + rb_pl * STG_PLANE + rb_row * 48 + rb_half * 16;
  const int rb_ch = c_w + 8 * rb_half;
  const bool rb_on = rb_pl < NOUT && rb_ch < C;
  uint16_t* const out_base = reinterpret_cast<uint16_t*>(rb_pl == 0 ? p.y : p.y_lo) + item + rb_ch;

  // rows [8 j, 8 j + 8) of the walk -> ring (one commit group per call, always: uniform accounting)
  auto issue_group = [&](int j) {
    const int t = min(max(R0 + 8 * j + cp_row, 0), L - 1);
    if (cp_lane)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(cp_dst + (uint32_t)(((8 * j) & (RS_ROWS - 1)) * RS_PITCH)),
                   "l"(cp_src + t * cp_stride), "r"(cp_bytes)
                   : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  float s_last[2] = {0.f, 0.f};
  struct SFrag {
    uint32_t hi[4], lo[BF16_PATH ? 1 : 4];
  };

  // s-block m from the 16 ring rows starting at walk row `row0` (a multiple of 8)
  auto s_block = [&](auto edge_tag, int m, int row0, SFrag& out) {
    constexpr bool EDGE = decltype(edge_tag)::value;
    [[maybe_unused]] uint32_t xh[4], xl[4];
    if constexpr (BF16_PATH) {
      amm::ldmatrix_x4_trans(ring_u32 + (uint32_t)(((row0 + frag_row8) & (RS_ROWS - 1)) * RS_PITCH) + frag_off, xh);
    } else {
#pragma unroll
      for (int j2 = 0; j2 < 2; ++j2) {
        const uint32_t base = ring_u32 + (uint32_t)(((row0 + 8 * j2) & (RS_ROWS - 1)) * RS_PITCH) + frag_off;
#pragma unroll
        for (int i2 = 0; i2 < 2; ++i2) {
          float v0, v1;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v0) : "r"(base + (uint32_t)(i2 * 32)));
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v1) : "r"(base + (uint32_t)(i2 * 32 + RS_PITCH)));
          amm::split_pair(v0, v1, xh[2 * j2 + i2], xl[2 * j2 + i2]);
        }
      }
    }
    float d[2][4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
      for (int i = 0; i < 4; ++i) d[h][i] = 0.f;
      amm::mma_bf16(d[h], xh, up_hi[h]);
      amm::mma_bf16(d[h], xh, up_lo[h]);
      if constexpr (!BF16_PATH) amm::mma_bf16(d[h], xl, up_hi[h]);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {  // registers (0, 1) and (2, 3) are two time steps of one channel each
      amm::snake_pair<FAST_SIN>(d[h][0], d[h][1], apar[0], invb[0]);
      amm::snake_pair<FAST_SIN>(d[h][2], d[h][3], apar[1], invb[1]);
    }
    if constexpr (EDGE) {
      if (m < 0) {
        // left clamp (LowPassFilter1d pad, bigvgan.py:227): s[j < 0] = s[0]; row0 pointed at block 0
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
    if constexpr (BF16_PATH) {
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

  SFrag fa, fb;
  int it = 0;  // iterations done: iteration i consumes walk rows [8 (i + 1), 8 (i + 1) + 16)
  uint16_t* out = out_base + (long long)(8 * m_begin + 3 + rb_row) * C;

  // z-tile m_begin + it from (prev, cur = s-block m_begin + it + 1)
  auto step = [&](auto edge_tag, const SFrag& prev, SFrag& cur) {
    constexpr bool EDGE = decltype(edge_tag)::value;
    const int m = m_begin + it;
    // rows up to 8 it + 23 (groups 0 .. it + 2) have landed once at most RS_D - 1 groups are pending
    asm volatile("cp.async.wait_group %0;" ::"n"(RS_D - 1) : "memory");
    __syncwarp();
    issue_group(it + RS_D + 2);  // overwrites rows [8 it - 16, 8 it - 9]: last read two iterations ago
    s_block(edge_tag, m + 1, 8 * (it + 1), cur);
    float z[4] = {0.f, 0.f, 0.f, 0.f}, z2[4] = {0.f, 0.f, 0.f, 0.f};
    if constexpr (BF16_PATH) {
      amm::mma_f16(z, prev.hi, dn_hi[0]);
      amm::mma_f16(z2, cur.hi, dn_hi[1]);
    } else {
      amm::mma_bf16(z, prev.hi, dn_hi[0]);
      amm::mma_bf16(z2, cur.hi, dn_hi[1]);
      amm::mma_bf16(z, prev.hi, dn_lo[0]);
      amm::mma_bf16(z2, cur.hi, dn_lo[1]);
      amm::mma_bf16(z, prev.lo, dn_hi[0]);
      amm::mma_bf16(z2, cur.lo, dn_hi[1]);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) z[r] += z2[r];
    const uint32_t sbuf = (uint32_t)((it & 1) * STG);
    if constexpr (BF16_PATH) {
      amm::stmatrix_x2_trans(st_addr + sbuf, pack_bf16x2(z[0], z[1]), pack_bf16x2(z[2], z[3]));
    } else {
      uint32_t h0, l0, h1, l1;
      amm::split_pair(z[0], z[1], h0, l0);
      amm::split_pair(z[2], z[3], h1, l1);
      amm::stmatrix_x4_trans(st_addr + sbuf, h0, h1, l0, l1);
    }
    __syncwarp();
    bool on = rb_on;
    if constexpr (EDGE) {
      const int t = 8 * m + 3 + rb_row;
      on = on && t >= 0 && t < L;
    }
    if (on) *reinterpret_cast<uint4*>(out) = *reinterpret_cast<const uint4*>(rb_ptr + sbuf);
    out += 8 * (long long)C;
    ++it;
  };

  // prologue: RS_D + 2 row groups in flight, then the warm-up s-block (no output)
#pragma unroll
  for (int j = 0; j < RS_D + 2; ++j) issue_group(j);
  asm volatile("cp.async.wait_group %0;" ::"n"(RS_D - 1) : "memory");  // groups 0 .. 2
  __syncwarp();
  // s-block m_begin from rows [0, 16), or block 0 (rows [8, 24)) broadcast as block -1
  s_block(std::true_type{}, m_begin, m_begin < 0 ? 8 : 0, fa);

  for (int mt = m_begin; mt <= m_end; mt += AM_NB) {
    const bool edge = mt < 0 || 16 * (mt + AM_NB) + 15 >= jl;
    if (edge) {
#pragma unroll 1
      for (int i = 0; i < AM_NB && mt + i <= m_end; i += 2) {
        step(std::true_type{}, fa, fb);
        if (mt + i + 1 > m_end) break;
        step(std::true_type{}, fb, fa);
      }
    } else {
#pragma unroll 1
      for (int i = 0; i < AM_NB; i += 2) {
        step(std::false_type{}, fa, fb);
        step(std::false_type{}, fb, fa);
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");  // nothing may be in flight into shared memory at exit
}

// bvg_tuning.amp_stream: 1 = F32 -> SPLIT runs here.  Off by default: measured on B200 (in-program, fp32
// path) 3.0-3.2 TB/s against 3.5-3.6 TB/s for the FFMA2 kernel -- ncu: 46 thread-instructions per element, a third of
// them the bf16 (hi, lo) splitting of x, s and z that the fp32 path's 16-bit operands need, so the tensor-core
// formulation does not pay here the way it does on the bf16 path (amp_mma.cu, fp16 single-term s).

// bvg_tuning.amp_stream_bf16: BF16 -> BF16 on the streaming kernel (1) or on the staged amp_mma kernel (0, default: the
// staged kernel is level for C >= 48 and 15 % ahead for C = 24, gpurun_out/ab_streambf16.txt)

bool amp_stream_supported(const bvg_amp_desc* d) {
  if (d->C % 8 != 0) return false;
  const bool f32_split = d->x.dtype == BVG_F32 && d->y.dtype == BVG_SPLIT;
  const bool bf_bf = d->x.dtype == BVG_BF16 && d->y.dtype == BVG_BF16;
  const bvg_tuning T = tune_of(d->tune);
  if (!((f32_split && T.amp_stream) || (bf_bf && T.amp_stream_bf16))) return false;
  if (((uintptr_t)d->x.d_ptr & 15) || ((uintptr_t)d->y.d_ptr & 15) || (f32_split && ((uintptr_t)d->y.d_lo & 15))) return false;
  return true;
}

int amp_stream_forward(const bvg_amp_desc* d, cudaStream_t st) {
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
  p.n_cg = ceil_div(d->C, 16 * RS_WARPS);
  p.m_last = d->L >= 4 ? (d->L - 4) / 8 : -1;
  p.n_tiles = ceil_div(p.m_last + 2, AM_NB);
  const int amp_mma_tiles = tune_of(d->tune).amp_mma_tiles;
  int tpc = amp_mma_tiles > 0 ? amp_mma_tiles : 32;
  while (tpc > 1 && (long long)d->B * p.n_cg * ceil_div(p.n_tiles, tpc) < 148ll * 5 * 3) tpc >>= 1;
  p.tiles_per_cta = tpc;
  p.n_ct = ceil_div(p.n_tiles, tpc);
  const long long blocks = (long long)d->B * p.n_ct * p.n_cg;
  BVG_REQUIRE(blocks < (1ll << 31), "amp: grid too large");
  const bool bf = d->x.dtype == BVG_BF16;
  if (bf && d->fast_sin)
    amp_stream_kernel<true, true><<<(unsigned)blocks, 32 * RS_WARPS, 0, st>>>(p);
  else if (bf)
    amp_stream_kernel<true, false><<<(unsigned)blocks, 32 * RS_WARPS, 0, st>>>(p);
  else if (d->fast_sin)
    amp_stream_kernel<false, true><<<(unsigned)blocks, 32 * RS_WARPS, 0, st>>>(p);
  else
    amp_stream_kernel<false, false><<<(unsigned)blocks, 32 * RS_WARPS, 0, st>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "amp_stream_kernel launch");
  return BVG_OK;
}

}  // namespace bvg
