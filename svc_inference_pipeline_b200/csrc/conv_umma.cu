// K-C / K-T: dense convolutions as a tap GEMM on the 5th-gen tensor cores (sm_100a).
//
// Replaces the weight-normed Conv1d layers of the AMP blocks and conv_pre, and the
// ConvTranspose1d upsamplers (reference modules/bigvgan.py:319-386/:428-431, :529-537/:602,
// :547-561/:607) in the formulation of include/bvg_b200.h:
//     D[t, n] = sum_{tap} sum_{ci} X[t + shift(tap), ci] * W[n][tap][ci]
// GEMM view per CTA tile: M = 128 time rows (TMEM lanes), N = n_tile output channels (TMEM
// columns, fp32), K = taps x Cin streamed in 64-channel slices.
//
//  * operands are bf16, staged in shared memory by TMA (cp.async.bulk.tensor, 128B swizzle):
//      A: one [rows x 64ch] box of the channels-last activation per Cin slice.  In HALO mode the
//         box carries 128 + (max shift - min shift) rows and every tap reads it through a UMMA
//         descriptor whose start address is advanced by (shift - min shift) rows, so a dilated
//         k=11 layer loads each activation row once instead of 11 times.  Rows outside [0, L)
//         are zero-filled by TMA (a 3-D map [B][L][C] keeps batch items apart) = Conv1d padding.
//      B: one [n_tile x 64ch] box of the packed weights per (tap, Cin slice).
//  * tcgen05.mma (cta_group::1, kind::f16, M=128, N=n_tile, K=16) issued by one thread,
//    accumulating in TMEM; a tile is mb row blocks of 128 sharing every weight box, with as many
//    accumulator stages as fit in the 512 columns so the epilogue of tile i overlaps tile i+1;
//  * SPLIT operands (fp32-parity path): A and W come as (hi, lo) bf16 planes and each K slice
//    issues hi*hi + lo*hi + hi*lo (3 MMAs, fp32 accumulate): 16 mantissa bits per operand
//    (narrow layers: [W_hi; W_lo] stacked along N, 2 MMAs);
//  * warp roles: warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
//    warps 4-11 = epilogue (tcgen05.ld -> swizzled staging -> bias/residual/accumulate/divide
//    -> F32 | BF16 | SPLIT stores), fused mode: warps 12-19 = Activation1d operand producer.
//    Persistent: grid = min(#tiles, #SMs), static round-robin.
//
// Every mbarrier wait is bounded: a pipeline bug traps (launch failure) instead of hanging the GPU.
#include <cuda.h>

#include <cstring>

#include "amp_p2.cuh"
#include "conv_umma.cuh"

namespace bvg {

// MMAs of one weight stage: gtaps taps x mb M blocks x KS K steps (twice with the lo plane of a
// SPLIT activation).  The issuing warp executes a few instructions per tcgen05.mma at one dependent
// uniform-datapath instruction every ~4-5 cycles, so for narrow N (an N = 32 MMA lasts 16-40 cycles)
// and for single-M-block tiles this warp is on the critical path (ncu, v5: 89 cycles per MMA at 20
// instructions each): K steps are compile-time, the accumulate flag is a predicate, descriptors
// advance by immediate adds, and the row shift of the next tap is fetched before this tap's MMAs.
//   sh      : shift of the first tap of the stage (prefetched by the caller, before the barrier wait)
//   nxt_tap : tap whose shift is to be returned for the caller's next stage
//   STACKED : the stage holds [W_hi; W_lo] per tap: A_hi x both (idesc2, N = 2 n_tile) then A_lo x W_hi (idesc, N = n_tile)
template <int KS, bool WITH_LO, bool STACKED = false>
__device__ __forceinline__ int issue_stage(const UmmaParams& p, const int* __restrict__ shifts, int g0, int gtaps, int sh, int nxt_tap,
                                           int min_shift, uint32_t a_stage16, uint32_t b16, uint32_t tmem_acc, uint32_t a_plane16, uint32_t idesc,
                                           uint64_t desc_hi, uint32_t acc0, uint32_t b_tap16, uint32_t col_stride, int mb, uint32_t idesc2 = 0) {
  const uint32_t corr_col = (uint32_t)p.n_tile;
  for (int g = 0; g < gtaps; ++g) {
    uint32_t a16 = a_stage16 + (uint32_t)(sh - min_shift) * 8u;  // 128 B per row
    sh = shifts[g + 1 < gtaps ? g0 + g + 1 : nxt_tap];
    uint32_t tmem_d = tmem_acc;
    const uint32_t first = g == 0 ? acc0 : 1u;
    for (int mbi = 0; mbi < mb; ++mbi) {
#pragma unroll
      for (int k = 0; k < KS; ++k) {
        const uint64_t bd = desc_hi | (uint64_t)(b16 + 2u * k);
        if (ptx::elect_one()) ptx::umma_f16(tmem_d, desc_hi | (uint64_t)(a16 + 2u * k), bd, STACKED ? idesc2 : idesc, k == 0 ? first : 1u);
        if (WITH_LO) {
          // STACKED: the lo*hi products join the hi*lo ones in the correction columns [n_tile, 2 n_tile), so the main
          // accumulator takes one addition per K step instead of two.  The tensor core truncates when it adds into an
          // fp32 accumulator (error ~ steps * 2^-24 * |acc|: measured 3.4e-9 * K relative, tools/conv_precision_diag.py)
          // and the correction sum is 2^-8 of the main one, so its own truncation does not count.
          if (ptx::elect_one()) ptx::umma_f16(tmem_d + (STACKED ? corr_col : 0u), desc_hi | (uint64_t)(a16 + a_plane16 + 2u * k), bd, idesc, 1u);
        }
      }
      a16 += (UM_BM * 128u) >> 4;
      tmem_d += col_stride;
    }
    b16 += b_tap16;
  }
  return sh;
}

// ------------------------------------------------------------------------------------------------
// Fused Activation1d producer (bvg_conv_desc.pre_amp): warps 12-19 compute the A operand of
// the convolution -- z = DownSample1d(snake(UpSample1d(x))) of reference modules/bigvgan.py:251-256 -- from the
// Activation1d's fp32 input and write the (hi, lo) bf16 planes straight into the 128-byte-swizzled A stages the
// UMMA descriptors read, so the activated tensor never exists in HBM (saves its 4-byte write and 4-byte read per
// element and the separate kernel) and the FFMA work runs under the tile's MMAs.  Arithmetic and operation order
// are amp_kernel_p2's (amp_kernel.cu); a thread owns one channel pair and a run of consecutive rows of the
// tile, rows outside [0, L) are the convolution's zero padding.  Narrow layers only (Cin <= 64 * a_stages): all
// Cin slices of a tile are produced together.
// ------------------------------------------------------------------------------------------------
template <bool FAST_SIN>
__device__ __forceinline__ void fused_block6(const UmmaParams& p, const P2 (&xa)[6], P2 (&xb)[6], const P2 (&sa)[12], P2 (&sb)[12], int tau0,
                                             const float* __restrict__ xcol, int t_lo, int t_hi, int trow0, uint32_t pair_base, P2 apar, P2 invb,
                                             P2 zc, bool store) {
  const int L = p.L, C = p.f_C;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const int r = min(max(tau0 + 6 + j, 0), L - 1);
    const float2 t = *reinterpret_cast<const float2*>(xcol + (long long)r * C);
    xb[j] = pk2(t.x, t.y);
  }
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    P2 pa = 0ull, pb = 0ull;
#pragma unroll
    for (int m = 0; m < 6; ++m) {
      const int w = j + 5 - m;
      const P2 xv = (w < 6) ? xa[w < 6 ? w : 0] : xb[w >= 6 ? w - 6 : 0];
      pb = fma2(pk2(p.f_gu[2 * m + 1], p.f_gu[2 * m + 1]), xv, pb);
      pa = fma2(pk2(p.f_gu[2 * m], p.f_gu[2 * m]), xv, pa);
    }
    pa = snake_two<FAST_SIN>(pa, apar, invb);
    pb = snake_two<FAST_SIN>(pb, apar, invb);
    const int tau = tau0 + j;
    if (tau >= L - 3) {  // right replicate clamp of the activated signal: s[j > 2L-1] = s[2L-1]
      const P2 prev = (j == 0) ? sa[11] : sb[j == 0 ? 0 : 2 * j - 1];
      if (tau >= L - 2) pa = prev;
      pb = pa;
    }
    sb[2 * j] = pa;
    sb[2 * j + 1] = pb;
    if (store) {
      P2 z = zc;
#pragma unroll
      for (int k = 0; k < 12; ++k) {
        const int i = 2 * j + 2 + k;
        const P2 sv = (i < 12) ? sa[i < 12 ? i : 0] : sb[i >= 12 ? i - 12 : 0];
        z = fma2(pk2(p.f_fd[k], p.f_fd[k]), sv, z);
      }
      if (tau >= t_lo && tau < t_hi) {
        float z0, z1;
        upk2(z, z0, z1);
        const uint32_t hi = pack_bf16x2(z0, z1);
        const uint32_t row = (uint32_t)(tau - trow0);
        const uint32_t addr = (pair_base + row * 128u) ^ ((row & 7u) << 4);
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(hi) : "memory");
        if (p.planes == 2) {
          const P2 lo = fma2(pk2(__uint_as_float(hi << 16), __uint_as_float(hi & 0xffff0000u)), pk2(-1.0f, -1.0f), z);
          float l0, l1;
          upk2(lo, l0, l1);
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr + (uint32_t)p.a_plane_bytes), "r"(pack_bf16x2(l0, l1)) : "memory");
        }
      }
    }
  }
}

__device__ __forceinline__ void fused_amp_producer(const UmmaParams& p, uint32_t a_base, uint32_t bar_base) {
  const int ptid = (int)threadIdx.x - UM_THREADS;  // 0 .. 255
  const int lane = threadIdx.x & 31;
  const int C = p.f_C, L = p.L;
  const int pairs = C >> 1;
  const int n_sub = (32 * UM_AMP_WARPS) / pairs;  // row runs per tile
  const int pi = ptid % pairs, sub = ptid / pairs;
  const bool active = sub < n_sub;
  const int need = p.f_need;
  const int rps = (need + n_sub - 1) / n_sub;
  const int r0 = sub * rps, r1 = min(r0 + rps, need);
  const int c = 2 * pi;
  const int cb_mine = c >> 6;
  const uint32_t col_bytes = (uint32_t)((c & 63) * 2);
  const float a0 = __ldg(p.f_a + c), a1 = __ldg(p.f_a + c + 1);
  const float ib0 = __ldg(p.f_invb + c), ib1 = __ldg(p.f_invb + c + 1);
  const bool fast = p.f_fast_sin != 0;
  // cosine-form constants (amp_p2.cuh): 2a | a/pi, -invb/2, (invb/2) sum(taps)
  const P2 apar = fast ? pk2(snake_apar2<true>(a0), snake_apar2<true>(a1)) : pk2(snake_apar2<false>(a0), snake_apar2<false>(a1));
  const P2 invb = pk2(snake_hbn(ib0), snake_hbn(ib1));
  const P2 zc = pk2(snake_zc(ib0, p.f_fsum), snake_zc(ib1, p.f_fsum));
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (UM_MAX_A_STAGES + s); };

  int sa = 0, pa = 0;
  for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
    const int nt = (int)(tile % p.n_tiles);
    const long long mt = tile / p.n_tiles;
    const int b = (int)(mt / p.m_tiles_per_item);
    const int t0 = (int)(mt % p.m_tiles_per_item) * p.tile_rows;
    const int trow0 = t0 + p.min_shift[nt];  // time of stage row 0
    // stages of this tile's Cin slices (n_cb <= a_stages): wait until the MMAs of their previous use retired
    uint32_t my_stage = 0;
    {
      int s = sa, par = pa;
      for (int cb = 0; cb < p.n_cb; ++cb) {
        ptx::mbar_wait(a_empty(s), par ^ 1, p.err_flag, 7);
        if (cb == cb_mine) my_stage = a_base + (uint32_t)(s * p.a_stage_bytes);
        if (++s == p.a_stages) { s = 0; par ^= 1; }
      }
    }
    // channels [C, round_up(C, 16)) of the last slice are read by its last K step: zero them (TMA's out-of-bounds fill)
    if (C & 15) {
      int s_last = sa + p.n_cb - 1;
      if (s_last >= p.a_stages) s_last -= p.a_stages;
      const uint32_t base = a_base + (uint32_t)(s_last * p.a_stage_bytes) + (uint32_t)((C & 63) * 2);
      for (int r = ptid; r < need; r += 32 * UM_AMP_WARPS) {
        const uint32_t addr = (base + (uint32_t)r * 128u) ^ (((uint32_t)r & 7u) << 4);
        for (int pl = 0; pl < p.planes; ++pl)
          asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(addr + (uint32_t)(pl * p.a_plane_bytes)), "r"(0u) : "memory");
      }
    }
    if (active && r0 < r1) {
      const uint32_t pair_base = my_stage + col_bytes;
      const int T0 = trow0 + r0, T1 = trow0 + r1;
      const int ta = max(T0, 0), tb = min(T1, L);
      // rows outside the sequence: Conv1d's zero padding
      for (int t = T0; t < T1; ++t) {
        if (t >= ta && t < tb) continue;
        const uint32_t row = (uint32_t)(t - trow0);
        const uint32_t addr = (pair_base + row * 128u) ^ ((row & 7u) << 4);
        for (int pl = 0; pl < p.planes; ++pl) asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr + (uint32_t)(pl * p.a_plane_bytes)), "r"(0u) : "memory");
      }
      if (ta < tb) {
        const float* xcol = p.f_x + (long long)b * L * C + c;
        P2 xa[6], xb[6], sa_w[12], sb_w[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) sa_w[k] = 0ull;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          const int r = min(max(ta - 6 + k, 0), L - 1);
          const float2 t = *reinterpret_cast<const float2*>(xcol + (long long)r * C);
          xa[k] = pk2(t.x, t.y);
        }
#define BVG_FBLOCK(XA, XB, SA, SB, TAU, STORE)                                                                  \
  if (fast) fused_block6<true>(p, XA, XB, SA, SB, TAU, xcol, ta, tb, trow0, pair_base, apar, invb, zc, STORE); \
  else fused_block6<false>(p, XA, XB, SA, SB, TAU, xcol, ta, tb, trow0, pair_base, apar, invb, zc, STORE)
        BVG_FBLOCK(xa, xb, sa_w, sb_w, ta - 6, false);  // warm-up: fills the s window, no output
        // left replicate clamp of the activated signal: s[j < 0] = s[0].  The window holds s[2 ta - 7 + k]; a run may
        // start at any row here (amp_kernel_p2's chunks start at 0 or far from it), so s[0] sits in slot 7 - 2 ta
        if (ta == 0) {
#pragma unroll
          for (int k = 0; k < 7; ++k) sb_w[k] = sb_w[7];
        } else if (ta == 1) {
#pragma unroll
          for (int k = 0; k < 5; ++k) sb_w[k] = sb_w[5];
        } else if (ta == 2) {
#pragma unroll
          for (int k = 0; k < 3; ++k) sb_w[k] = sb_w[3];
        }
#pragma unroll 1
        for (int t = ta; t < tb; t += 12) {
          BVG_FBLOCK(xb, xa, sb_w, sa_w, t, true);
          BVG_FBLOCK(xa, xb, sa_w, sb_w, t + 6, true);  // past tb: computed, not stored (keeps the windows in step)
        }
#undef BVG_FBLOCK
      }
    }
    // generic-proxy writes -> visible to the tensor core (async proxy), then one arrive per warp and slice
    ptx::fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      int s = sa;
      for (int cb = 0; cb < p.n_cb; ++cb) {
        ptx::mbar_arrive(a_full(s));
        if (++s == p.a_stages) s = 0;
      }
    }
    for (int cb = 0; cb < p.n_cb; ++cb)
      if (++sa == p.a_stages) { sa = 0; pa ^= 1; }
  }
}

// Epilogue specialisation (compile-time, so the per-element code is a handful of instructions):
//   OUT  : output format (BVG_F32 | BVG_BF16 | BVG_SPLIT)
//   SBF  : residual / running-sum tensors are bf16 (else fp32)
//   RES  : a residual is added            ACC : a running sum is added (and maybe divided)
//   GEN  : generic fallback -- every choice read from the descriptor at run time, ragged N allowed
template <int OUT, bool SBF, bool RES, bool ACC, bool GEN, bool FUSED = false>
__global__ void __launch_bounds__(FUSED ? UM_THREADS_FUSED : 128 + 32 * UM_EPI_WARPS_MAX, 1) conv_umma_kernel(const __grid_constant__ UmmaParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve-up: [A stages][B stages][epilogue staging][barriers][tmem ptr]; base rounded up to 1024 B
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = smem_base;
  const uint32_t b_base = a_base + p.a_stages * p.a_stage_bytes;
  const uint32_t stg_base = b_base + p.b_stages * p.b_stage_bytes;
  const int n_epi = FUSED ? UM_EPI_WARPS : p.epi_warps;
  const uint32_t bar_base = stg_base + (uint32_t)n_epi * 2048u;  // 32 rows x 16 fp32 of staging per epilogue warp
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (UM_MAX_A_STAGES + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (2 * UM_MAX_A_STAGES + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (2 * UM_MAX_A_STAGES + UM_MAX_B_STAGES + s); };
  auto t_full = [&](int s) { return bar_base + 8u * (2 * UM_MAX_A_STAGES + 2 * UM_MAX_B_STAGES + s); };
  auto t_empty = [&](int s) { return bar_base + 8u * (2 * UM_MAX_A_STAGES + 2 * UM_MAX_B_STAGES + UM_MAX_T_STAGES + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * UM_MAX_A_STAGES + 2 * UM_MAX_B_STAGES + 2 * UM_MAX_T_STAGES);
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));  // generic pointer to smem_base
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_gen + (tmem_slot - smem_base));

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // provably warp-uniform
  const int lane = threadIdx.x & 31;
  pdl_trigger();  // the next launch of the chain may set itself up while this one runs (common.cuh)

  if (warp == 0 && lane == 0) {
    if (!FUSED) ptx::prefetch_tmap(&p.tm_x[0]);
    ptx::prefetch_tmap(&p.tm_w[0]);
    if (p.planes == 2) {
      if (!FUSED) ptx::prefetch_tmap(&p.tm_x[1]);
      if (!p.stacked) ptx::prefetch_tmap(&p.tm_w[1]);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.a_stages; ++s) {
      ptx::mbar_init(a_full(s), FUSED ? UM_AMP_WARPS : 1);  // fused: one arrive per producer warp
      ptx::mbar_init(a_empty(s), 1);
    }
    for (int s = 0; s < p.b_stages; ++s) {
      ptx::mbar_init(b_full(s), 1);
      ptx::mbar_init(b_empty(s), 1);
    }
    for (int s = 0; s < p.t_stages; ++s) {
      ptx::mbar_init(t_full(s), 1);
      ptx::mbar_init(t_empty(s), n_epi);  // one arrive per epilogue warp
    }
    ptx::fence_barrier_init();
    ptx::fence_proxy_async();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();  // everything above overlapped the previous launch; its results are visible from here on

  const int planes = p.planes;
  const int stage_cols = p.mb * p.col_stride;

  if (warp == 0) {
    // ================================ TMA producer ================================
    // The whole warp runs the loop convergently (all operands stay in uniform registers); only
    // the TMA / mbarrier instructions themselves are predicated on one elected lane.
    {
      int sa = 0, pa = 0, sb = 0, pb = 0;
      const int box_bytes = p.a_box_rows * 128;
      for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int nt = (int)(tile % p.n_tiles);
        const long long mt = tile / p.n_tiles;
        const int b = (int)(mt / p.m_tiles_per_item);
        const int t0 = (int)(mt % p.m_tiles_per_item) * p.tile_rows;
        const int ntaps = p.n_taps[nt];
        const int row0 = t0 + p.min_shift[nt];
        for (int cb = 0; cb < p.n_cb; ++cb) {
          // A super-tile: rows [row0, row0 + a_boxes * a_box_rows) x 64 channels, every plane
          // (fused mode: written by the Activation1d producer warps instead)
          if constexpr (!FUSED) {
            ptx::mbar_wait(a_empty(sa), pa ^ 1, p.err_flag, 1);
            if (ptx::elect_one()) {
              ptx::mbar_expect_tx(a_full(sa), (uint32_t)p.a_stage_bytes);
              for (int pl = 0; pl < planes; ++pl)
                for (int bx = 0; bx < p.a_boxes; ++bx)
                  ptx::tma_load_3d(a_base + sa * p.a_stage_bytes + pl * p.a_plane_bytes + bx * box_bytes, &p.tm_x[pl], a_full(sa), cb * UM_KB,
                                   row0 + bx * p.a_box_rows, b);
            }
            __syncwarp();
            if (++sa == p.a_stages) { sa = 0; pa ^= 1; }
          }
          // weights: one box of tap_group consecutive taps per (group, plane); rows past the last
          // tap of this N tile belong to the next tile (or are zero-filled past the end) and are unused
          const int w_planes = p.stacked ? 1 : planes;   // stacked: hi and lo rows arrive in one box
          const int w_rows = p.stacked ? 2 * p.n_tile : p.n_tile;
          for (int g0 = 0; g0 < ntaps; g0 += p.tap_group) {
            for (int wp = 0; wp < w_planes; ++wp) {
              ptx::mbar_wait(b_empty(sb), pb ^ 1, p.err_flag, 2);
              if (ptx::elect_one()) {
                ptx::mbar_expect_tx(b_full(sb), (uint32_t)p.b_stage_bytes);
                ptx::tma_load_2d(b_base + sb * p.b_stage_bytes, &p.tm_w[wp], b_full(sb), cb * UM_KB,
                                 (nt * p.tap_stride + g0) * w_rows);
              }
              __syncwarp();
              if (++sb == p.b_stages) { sb = 0; pb ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    // One thread feeds the tensor pipe, so its instruction count per MMA is the budget that
    // matters (a lone warp issues one dependent instruction every few cycles): descriptors are
    // kept as a constant high word plus a low word that only needs integer adds, and the warp
    // stays converged so that every tcgen05 operand lives in a uniform register -- issuing from
    // inside an `if (lane == 0)` region makes the compiler wrap each UTCHMMA in an
    // ELECT / R2UR.BROADCAST "waterfall" loop (~15 instructions per MMA, measured 2-10x slower).
    {
      const uint32_t idesc = make_idesc(p.n_tile);
      const uint64_t desc_hi = (uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29)) << 32;  // SBO, version, SWIZZLE_128B
      const uint32_t a_plane16 = (uint32_t)p.a_plane_bytes >> 4;
      const bool split = p.planes == 2;
      const int tap_group = p.tap_group;
      const int mb = p.mb;
      const bool stacked = p.stacked != 0;
      const uint32_t idesc2 = make_idesc(2 * p.n_tile);
      const uint32_t b_tap16 = (uint32_t)p.n_tile * (stacked ? 16u : 8u);  // rows of 128 B per tap
      const uint32_t col_stride = (uint32_t)p.col_stride;
      const int ks_last = (p.cin - (p.n_cb - 1) * UM_KB + 15) >> 4;
      int sa = 0, pa = 0, sb = 0, pb = 0, as = 0, ap = 0;
      for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int nt = (int)(tile % p.n_tiles);
        const int ntaps = p.n_taps[nt];
        const int min_shift = p.min_shift[nt];
        ptx::mbar_wait(t_empty(as), ap ^ 1, p.err_flag, 3);
        ptx::tc_fence_after();
        const uint32_t tmem_acc = tmem_base + (uint32_t)(as * stage_cols);
        const int* shifts = p.shift[nt];
        int sh = shifts[0];
        for (int cb = 0; cb < p.n_cb; ++cb) {
          const int ksteps = cb + 1 < p.n_cb ? 4 : ks_last;
          ptx::mbar_wait(a_full(sa), pa, p.err_flag, 4);
          const uint32_t a_stage16 = (((a_base + sa * p.a_stage_bytes) & 0x3FFFFu) >> 4) | (1u << 16);
          for (int g0 = 0; g0 < ntaps; g0 += tap_group) {
            const int gtaps = min(tap_group, ntaps - g0);
            const int g_next = g0 + tap_group < ntaps ? g0 + tap_group : 0;
            const uint32_t acc0 = (cb == 0 && g0 == 0) ? 0u : 1u;  // 0 => the first tap overwrites the accumulator
            // one straight-line body per (K steps, lo plane) combination, chosen once per stage
#define BVG_ISSUE(KS, LO, NXT, ACC)                                                                                                          \
  sh = (LO && stacked) ? issue_stage<KS, true, true>(p, shifts, g0, gtaps, sh, NXT, min_shift, a_stage16, b16, tmem_acc, a_plane16, idesc,    \
                                                      desc_hi, ACC, b_tap16, col_stride, mb, idesc2)                                          \
                       : issue_stage<KS, LO>(p, shifts, g0, gtaps, sh, NXT, min_shift, a_stage16, b16, tmem_acc, a_plane16, idesc, desc_hi,   \
                                             ACC, b_tap16, col_stride, mb)
#define BVG_STAGE(LO, NXT, ACC)                                                              \
  {                                                                                          \
    ptx::mbar_wait(b_full(sb), pb, p.err_flag, 5);                                           \
    ptx::tc_fence_after();                                                                   \
    const uint32_t b16 = (((b_base + sb * p.b_stage_bytes) & 0x3FFFFu) >> 4) | (1u << 16);   \
    if (ksteps == 4) BVG_ISSUE(4, LO, NXT, ACC);                                             \
    else if (ksteps == 2) BVG_ISSUE(2, LO, NXT, ACC);                                        \
    else if (ksteps == 3) BVG_ISSUE(3, LO, NXT, ACC);                                        \
    else BVG_ISSUE(1, LO, NXT, ACC);                                                         \
    if (ptx::elect_one()) ptx::umma_commit(b_empty(sb)); /* frees the stage when its MMAs retire */ \
    __syncwarp();                                                                            \
    if (++sb == p.b_stages) { sb = 0; pb ^= 1; }                                             \
  }
            if (stacked) {
              BVG_STAGE(true, g_next, acc0);   // [W hi; W lo] in one stage: A hi x both, A lo x W hi
            } else if (split) {
              BVG_STAGE(true, g0, acc0);       // W hi: A hi and A lo
              BVG_STAGE(false, g_next, 1u);    // W lo: A hi
            } else {
              BVG_STAGE(false, g_next, acc0);
            }
#undef BVG_STAGE
#undef BVG_ISSUE
          }
          if (ptx::elect_one()) ptx::umma_commit(a_empty(sa));
          __syncwarp();
          if (++sa == p.a_stages) { sa = 0; pa ^= 1; }
        }
        if (ptx::elect_one()) ptx::umma_commit(t_full(as));  // accumulators complete -> epilogue
        __syncwarp();
        if (++as == p.t_stages) { as = 0; ap ^= 1; }
      }
    }
  } else if (warp >= 4 && warp < 4 + n_epi) {
    // ================================ epilogue ====================================
    // Warp e owns TMEM lane quarter q = e % 4 (rows 32q..32q+31 of every M block) and every second
    // (M block, 16-column chunk) work item.  Each chunk goes TMEM -> registers (thread = row) ->
    // swizzled shared staging -> registers (4 lanes per row, 4 columns each), so that all global
    // traffic (residual, running sum, output) is made of whole 32-byte sectors per row.
    const int e = warp - 4;
    const int q = e & 3;
    const int grp = e >> 2;
    const uint32_t stg = stg_base + (uint32_t)e * 2048u;
    const int n_chunks = p.n_tile >> 4;
    const int n_items = p.mb * n_chunks;
    const int rrow = lane >> 2;   // row within an 8-row group after the transposition
    const int g = lane & 3;       // 16-byte granule (4 columns) within the 16-column chunk
    const int NGRP = n_epi >> 2;  // warps per lane quarter
    const int n_my = n_items > grp ? (n_items - grp + NGRP - 1) / NGRP : 0;  // items grp, grp + NGRP, ...
    const bool has_res = GEN ? (p.epi.res != nullptr) : RES;
    const bool has_acc = GEN ? (p.epi.acc != nullptr) : ACC;
    const bool res_bf = GEN ? (p.epi.res_dtype != BVG_F32) : SBF;
    const bool acc_bf = GEN ? (p.epi.acc_dtype != BVG_F32) : SBF;
    const int out_dt = GEN ? p.epi.out_dtype : OUT;
    const int N = p.epi.N;
    int as = 0, ap = 0;
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int nt = (int)(tile % p.n_tiles);
      const long long mt = tile / p.n_tiles;
      const int b = (int)(mt / p.m_tiles_per_item);
      const int t0 = (int)(mt % p.m_tiles_per_item) * p.tile_rows;
      const long long row_base = (long long)b * p.L;
      const uint32_t tmem_q = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * stage_cols);
      // items are handled two at a time: the residual of both is requested before anything waits
      // (for the first pair: before the accumulator is even complete), so that the loads of a
      // whole pair -- 8 per lane -- are in flight together instead of one latency per chunk
      for (int pb2 = 0; pb2 < n_my; pb2 += 2) {
        uint4 raw[2][4];
        int it_mbi[2], it_c0[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int it = grp + NGRP * (pb2 + u);
          const int mbi = it / n_chunks;
          it_mbi[u] = mbi;
          it_c0[u] = (it - mbi * n_chunks) << 4;
          if (has_res && (!GEN || p.vec_ok)) {
            const int tbase = t0 + mbi * UM_BM + q * 32;
            const int n0 = nt * p.n_tile + it_c0[u] + 4 * g;
            const bool ok = pb2 + u < n_my && n0 < N;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              raw[u][i] = make_uint4(0u, 0u, 0u, 0u);
              const int t = tbase + rrow + 8 * i;
              if (ok && t < p.L) {
                const long long off = (row_base + t) * N + n0;
                if (!res_bf) {
                  raw[u][i] = *reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(p.epi.res) + off);
                } else {
                  const uint2 h = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(p.epi.res) + off);
                  raw[u][i].x = h.x;
                  raw[u][i].y = h.y;
                }
              }
            }
          }
        }
        if (pb2 == 0) {
          ptx::mbar_wait(t_full(as), ap, p.err_flag, 6);
          ptx::tc_fence_after();
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (pb2 + u >= n_my) continue;
          const int mbi = it_mbi[u], c0 = it_c0[u];
          const int tbase = t0 + mbi * UM_BM + q * 32;
          if (tbase >= p.L) continue;  // whole 32-row slab past the end of the sequence (warp-uniform)
          uint32_t r[16];
          ptx::tmem_ld16(tmem_q + (uint32_t)(mbi * p.col_stride + c0), r);
          if (p.stacked) {  // the hi*lo products were accumulated n_tile columns further
            uint32_t r2[16];
            ptx::tmem_ld16(tmem_q + (uint32_t)(mbi * p.col_stride + p.n_tile + c0), r2);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
          }
          ptx::tmem_ld_wait();
          // stage: thread = row `lane`; granule gg of row r lives at r*64 + ((gg ^ ((r >> 1) & 3)) * 16)
          const uint32_t wsw = (uint32_t)((lane >> 1) & 3);
#pragma unroll
          for (int gg = 0; gg < 4; ++gg) {
            const uint32_t addr = stg + (uint32_t)lane * 64u + (((uint32_t)gg ^ wsw) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(r[4 * gg]), "r"(r[4 * gg + 1]), "r"(r[4 * gg + 2]), "r"(r[4 * gg + 3]) : "memory");
          }
          __syncwarp();
          const int n0 = nt * p.n_tile + c0 + 4 * g;
          float v[4][4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = rrow + 8 * i;
            const uint32_t addr = stg + (uint32_t)rr * 64u + (((uint32_t)g ^ (uint32_t)((rr >> 1) & 3)) << 4);
            uint32_t a0, a1, a2, a3;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(addr) : "memory");
            v[i][0] = __uint_as_float(a0); v[i][1] = __uint_as_float(a1); v[i][2] = __uint_as_float(a2); v[i][3] = __uint_as_float(a3);
          }
          __syncwarp();
          if (n0 >= N) continue;
          if (!GEN || p.vec_ok) {
            const float4 bias = *reinterpret_cast<const float4*>(p.epi.bias + n0);
            [[maybe_unused]] float4 cd = make_float4(1.f, 1.f, 1.f, 1.f);  // per-channel divisor: one load per chunk
            if (GEN && p.epi.coldiv) cd = __ldg(reinterpret_cast<const float4*>(p.epi.coldiv + n0));
            const long long off0 = (row_base + tbase + rrow) * N + n0;  // row i adds 8 * i * N
            float ac[4][4];
            if (has_acc) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
#pragma unroll
                for (int j = 0; j < 4; ++j) ac[i][j] = 0.f;
                if (tbase + rrow + 8 * i < p.L) epi_load4(p.epi.acc, acc_bf ? BVG_BF16 : BVG_F32, off0 + (long long)(8 * i) * N, ac[i]);
              }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (tbase + rrow + 8 * i >= p.L) continue;
              v[i][0] += bias.x; v[i][1] += bias.y; v[i][2] += bias.z; v[i][3] += bias.w;
              if (has_res) {
                float rs[4];
                if (!res_bf) {
                  rs[0] = __uint_as_float(raw[u][i].x); rs[1] = __uint_as_float(raw[u][i].y);
                  rs[2] = __uint_as_float(raw[u][i].z); rs[3] = __uint_as_float(raw[u][i].w);
                } else {
                  unpack_bf16x2(raw[u][i].x, rs[0], rs[1]);
                  unpack_bf16x2(raw[u][i].y, rs[2], rs[3]);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) v[i][j] += rs[j];
              }
              if (has_acc) {
#pragma unroll
                for (int j = 0; j < 4; ++j) v[i][j] += ac[i][j];
                if (p.epi.use_div) {
#pragma unroll
                  for (int j = 0; j < 4; ++j) v[i][j] = __fdiv_rn(v[i][j], p.epi.div);
                }
              } else if (GEN && p.epi.use_div) {
#pragma unroll
                for (int j = 0; j < 4; ++j) v[i][j] = __fdiv_rn(v[i][j], p.epi.div);
              }
              if (GEN && p.epi.coldiv) {
                v[i][0] = __fdiv_rn(v[i][0], cd.x); v[i][1] = __fdiv_rn(v[i][1], cd.y);
                v[i][2] = __fdiv_rn(v[i][2], cd.z); v[i][3] = __fdiv_rn(v[i][3], cd.w);
              }
              if (GEN && p.epi.relu) {
#pragma unroll
                for (int j = 0; j < 4; ++j) v[i][j] = fmaxf(v[i][j], 0.f);
              }
              const long long off = off0 + (long long)(8 * i) * N;
              if (out_dt == BVG_F32) {
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.epi.out) + off) = make_float4(v[i][0], v[i][1], v[i][2], v[i][3]);
              } else if (out_dt == BVG_BF16) {
                epi_store_bf16x4(p.epi.out, off, v[i]);
              } else {
                float hi[4], lo[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) split_bf16(v[i][j], hi[j], lo[j]);
                epi_store_bf16x4(p.epi.out, off, hi);
                epi_store_bf16x4(p.epi.out_lo, off, lo);
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int t = tbase + rrow + 8 * i;
              if (t >= p.L) continue;
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (n0 + j < p.N) epilogue1(p.epi, row_base + t, n0 + j, v[i][j]);
            }
          }
        }
      }
      if (n_my == 0) {
        ptx::mbar_wait(t_full(as), ap, p.err_flag, 6);
        ptx::tc_fence_after();
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(t_empty(as));
      if (++as == p.t_stages) { as = 0; ap ^= 1; }
    }
  }

  if constexpr (FUSED) {
    if (warp >= 4 + UM_EPI_WARPS) fused_amp_producer(p, a_base, bar_base);
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

int umma_encode_bf16_map(CUtensorMap* map, void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                           const cuuint32_t* box, const char* what) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return BVG_ECUDA;
  }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, base, dims, strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(%s) failed with CUresult %d (base %p, rank %d, dims %llu/%llu/%llu, box %u/%u/%u)", what, (int)r, base, rank,
              (unsigned long long)dims[0], (unsigned long long)dims[1], rank > 2 ? (unsigned long long)dims[2] : 0ull, box[0], box[1],
              rank > 2 ? box[2] : 0u);
    return BVG_ECUDA;
  }
  return BVG_OK;
}

struct UmmaLaunch {
  UmmaParams p;
  int grid;
  size_t smem;
};

// bvg_tuning knobs read here (bvg_conv_desc.tune): umma_mb (force M blocks per tile), umma_max_ctas (cap the persistent
// grid), umma_tap_group (taps per weight stage), umma_a_stages (activation stages), and
// umma_wide_mb2 -- mb = 2 with a single TMEM stage (2 x 256 columns): 0 = for SPLIT operands only (three MMA passes make a
// tile three times longer, so the un-overlapped epilogue costs ~1 % while every weight box feeds two row blocks:
// stage-0 convolutions 19.2 -> 17.9 ms per forward, gpurun_out/ab_mb2_fp32.txt; bf16 operands lose 13 %),
// 1 = for every tile that fits (also 192 columns), -1 = never
static const int kBarBytes = 8 * (2 * UM_MAX_A_STAGES + 2 * UM_MAX_B_STAGES + 2 * UM_MAX_T_STAGES) + 16;

// shared-memory plan for a given number of M blocks: picks the taps per weight stage (narrow N
// tiles group several taps into one TMA box / one barrier round trip) and returns the number of
// weight stages that fit
static int plan_smem(int staging_bytes, int umma_tap_group, int mb, int a_stages, int max_span, int planes, int n_tile, int max_taps, int* box_rows, int* boxes, int* tap_group) {
  const int rows = mb * UM_BM + max_span;
  const int nb = (rows + 255) / 256;
  const int br = (((rows + nb - 1) / nb) + 7) / 8 * 8;
  *box_rows = br;
  *boxes = nb;
  const int a_stage = nb * br * 128 * planes;
  const int avail = UM_SMEM_LIMIT - 1024 - kBarBytes - staging_bytes - a_stages * a_stage;
  if (avail <= 0) return 0;
  int g = umma_tap_group > 0 ? umma_tap_group : 24 * 1024 / (n_tile * 128);  // ~24 KB per stage
  if (g > 256 / n_tile) g = 256 / n_tile;                                      // TMA box <= 256 rows
  if (g > max_taps) g = max_taps;
  if (g < 1) g = 1;
  while (g > 1 && avail / (g * n_tile * 128) < 3) --g;
  *tap_group = g;
  int bs = avail / (g * n_tile * 128);
  return bs > UM_MAX_B_STAGES ? UM_MAX_B_STAGES : bs;
}

int conv_umma_prepare(const bvg_conv_desc* d, UmmaLaunch* out) {
  const bvg_conv_weights* w = d->w;
  BVG_REQUIRE(w->backend == BVG_UMMA, "conv_umma: weights were packed for another backend");
  // fused Activation1d producer: the operand is computed in the kernel from pre_amp->x (fp32); it has the format
  // the weights were packed for (SPLIT planes for split weights, else BF16)
  const bvg_amp_desc* fa = d->pre_amp;
  const bool fused = fa != nullptr;
  if (fused) {
    BVG_REQUIRE(fa->x.dtype == BVG_F32 && fa->x.d_ptr && fa->d_a && fa->d_invb, "conv_umma: the fused Activation1d needs an fp32 input and its parameters");
    BVG_REQUIRE(fa->B == d->B && fa->L == d->L && fa->C == w->cin, "conv_umma: fused Activation1d shape [%d, %d, %d] does not match the convolution [%d, %d, %d]",
                fa->B, fa->L, fa->C, d->B, d->L, w->cin);
    BVG_REQUIRE(w->cin % 8 == 0 && w->cin <= UM_KB && w->cin >= 8, "conv_umma: the fused Activation1d takes 8 <= Cin <= 64 (one K slice), Cin %% 8 == 0");
    BVG_REQUIRE(((uintptr_t)fa->x.d_ptr & 7) == 0, "conv_umma: fused Activation1d input must be 8-byte aligned");
  } else {
    BVG_REQUIRE(d->x.dtype == BVG_BF16 || d->x.dtype == BVG_SPLIT, "conv_umma: input must be BF16 or SPLIT");
  }
  const int planes = fused ? (w->split ? 2 : 1) : (d->x.dtype == BVG_SPLIT ? 2 : 1);
  BVG_REQUIRE(planes == 1 || (w->split && (w->split == 2 || w->d_w_lo) && (fused || d->x.d_lo)), "conv_umma: SPLIT input needs split-packed weights and a lo plane");
  const bool stacked = planes == 2 && w->split == 2;
  BVG_REQUIRE(!stacked || 2 * w->n_tile <= 256, "conv_umma: stacked weights need n_tile <= 128");
  BVG_REQUIRE((fused || d->x.d_ptr) && w->d_w, "conv_umma: null pointer");
  BVG_REQUIRE(d->B > 0 && d->L > 0, "conv_umma: bad shape");
  BVG_REQUIRE(w->n_tile % 16 == 0 && w->n_tile >= 16 && w->n_tile <= 256, "conv_umma: bad n_tile %d", w->n_tile);
  BVG_REQUIRE(w->n_tiles <= BVG_MAX_NTILES, "conv_umma: too many N tiles");
  BVG_REQUIRE(w->x_pitch % 8 == 0, "conv_umma: channel pitch %d must be a multiple of 8 (16-byte TMA strides)", w->x_pitch);
  BVG_REQUIRE((fused || ((uintptr_t)d->x.d_ptr & 15) == 0) && ((uintptr_t)w->d_w & 15) == 0, "conv_umma: operands must be 16-byte aligned");

  const bvg_tuning T = tune_of(d->tune);
  const int umma_mb = T.umma_mb, umma_wide_mb2 = T.umma_wide_mb2, umma_max_ctas = T.umma_max_ctas, umma_tap_group = T.umma_tap_group,
            umma_a_stages = T.umma_a_stages;
  UmmaParams& p = out->p;
  memset(&p, 0, sizeof(p));
  // four epilogue warps per TMEM lane quarter for bf16 operands (conv_umma.cuh), two otherwise and with the fused producer
  p.epi_warps = (planes == 1 && !fused) ? UM_EPI_WARPS_MAX : UM_EPI_WARPS;
  const int staging_bytes = p.epi_warps * 2048;
  int rc = fill_epilogue(d, p.epi);
  if (rc != BVG_OK) return rc;
  p.planes = planes;
  p.stacked = stacked ? 1 : 0;
  p.B = d->B;
  p.L = d->L;
  p.N = w->n_total;
  p.n_tile = w->n_tile;
  p.n_tiles = w->n_tiles;
  p.tap_stride = w->tap_stride;
  p.n_cb = w->cin_pad / UM_KB;
  p.cin = w->cin;
  p.vec_ok = (w->n_total % 4 == 0) ? 1 : 0;
  int max_span = 0, max_taps = 1;
  for (int t = 0; t < w->n_tiles; ++t) {
    if (w->n_taps[t] > max_taps) max_taps = w->n_taps[t];
    BVG_REQUIRE(w->n_taps[t] > 0 && w->n_taps[t] <= BVG_MAX_TAPS, "conv_umma: bad tap count");
    int lo = w->shift[t][0], hi = w->shift[t][0];
    for (int k = 0; k < w->n_taps[t]; ++k) {
      p.shift[t][k] = w->shift[t][k];
      lo = w->shift[t][k] < lo ? w->shift[t][k] : lo;
      hi = w->shift[t][k] > hi ? w->shift[t][k] : hi;
    }
    p.n_taps[t] = w->n_taps[t];
    p.min_shift[t] = lo;
    if (hi - lo > max_span) max_span = hi - lo;
  }

  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);

  // M blocks per tile: every weight box is then reused by mb MMAs and the per-tile pipeline
  // latencies (TMA round trip, accumulator hand-over, epilogue) are paid once per mb*128 rows.
  // Constraints: >= 2 accumulator stages in the 512 TMEM columns, >= 3 weight stages in shared
  // memory next to two A super-tiles, and enough tiles to fill the machine twice.
  const int w_rows = stacked ? 2 * w->n_tile : w->n_tile;  // weight rows per tap in a stage = accumulator columns
  p.col_stride = (w_rows + 31) / 32 * 32;
  int mb = 1;
  for (int cand = UM_MAX_MB; cand >= 1; cand >>= 1) {
    int br, nb, tg;
    const bool single_stage_ok = cand == 2 && cand * p.col_stride <= 512 &&
                                 (umma_wide_mb2 > 0 || (umma_wide_mb2 == 0 && planes == 2 && p.col_stride == 256));
    const bool tmem_ok = cand * p.col_stride * 2 <= 512 || (cand == 1) || single_stage_ok;
    if (!tmem_ok) continue;
    if (plan_smem(staging_bytes, umma_tap_group, cand, 2, max_span, planes, w_rows, max_taps, &br, &nb, &tg) < (cand == 1 ? 2 : 3)) continue;
    const long long tiles = (long long)d->B * ceil_div(d->L, cand * UM_BM) * w->n_tiles;
    if (cand > 1 && tiles < 2ll * sms) continue;
    mb = cand;
    break;
  }
  if (umma_mb > 0) mb = umma_mb;
  BVG_REQUIRE(mb >= 1 && mb <= UM_MAX_MB && mb * p.col_stride <= 512, "conv_umma: bad M-block count %d for n_tile %d", mb, w->n_tile);
  p.mb = mb;
  p.tile_rows = mb * UM_BM;
  p.m_tiles_per_item = ceil_div(d->L, p.tile_rows);
  p.total_tiles = (long long)d->B * p.m_tiles_per_item * w->n_tiles;
  p.t_stages = 512 / (mb * p.col_stride);
  if (p.t_stages > UM_MAX_T_STAGES) p.t_stages = UM_MAX_T_STAGES;
  // activation stages: a tile's MMAs cannot start before its whole super-tile has landed, so when one
  // tile is short (narrow N: ~2k cycles of MMA) two stages leave the TMA round trip exposed; take as many
  // (up to 4) as still leave three weight stages
  int a_stages = 2;
  for (int cand = UM_MAX_A_STAGES; cand > 2; --cand) {
    int br, nb, tg;
    if (umma_a_stages == -1 && p.n_cb <= 2 && plan_smem(staging_bytes, umma_tap_group, mb, cand, max_span, planes, w_rows, max_taps, &br, &nb, &tg) >= 3) {
      a_stages = cand;
      break;
    }
  }
  if (fused) {  // the producer of tile i+1 runs under the MMAs of tile i: one spare stage when it fits
    int br, nb, tg;
    if (plan_smem(staging_bytes, umma_tap_group, mb, 3, max_span, planes, w_rows, max_taps, &br, &nb, &tg) >= 3) a_stages = 3;
  }
  if (umma_a_stages >= 2 && umma_a_stages <= UM_MAX_A_STAGES) a_stages = umma_a_stages;
  p.a_stages = a_stages;
  const int bs = plan_smem(staging_bytes, umma_tap_group, mb, a_stages, max_span, planes, w_rows, max_taps, &p.a_box_rows, &p.a_boxes, &p.tap_group);
  BVG_REQUIRE(bs >= 2, "conv_umma: tile does not fit in shared memory (mb %d, span %d, n_tile %d, planes %d)", mb, max_span, w->n_tile, planes);
  p.b_stages = bs;
  p.a_plane_bytes = p.a_boxes * p.a_box_rows * 128;
  p.a_stage_bytes = p.a_plane_bytes * planes;
  p.b_stage_bytes = p.tap_group * w_rows * 128;
  size_t smem = 1024 + (size_t)p.a_stages * p.a_stage_bytes + (size_t)bs * p.b_stage_bytes + (size_t)staging_bytes + kBarBytes;
  // keep one CTA per SM (each allocates all 512 TMEM columns): ask for more than half the SM's smem
  if (smem < 120 * 1024) smem = 120 * 1024;
  BVG_REQUIRE(smem <= (size_t)UM_SMEM_LIMIT, "conv_umma: shared memory plan exceeds the limit");
  out->smem = smem;

  // tensor maps
  if (fused) {
    p.f_x = reinterpret_cast<const float*>(fa->x.d_ptr);
    p.f_a = fa->d_a;
    p.f_invb = fa->d_invb;
    for (int k = 0; k < 12; ++k) {
      p.f_gu[k] = 2.0f * fa->taps_up[k];
      p.f_fd[k] = fa->taps_down[k];
    }
    p.f_fsum = 0.f;
    for (int k = 0; k < 12; ++k) p.f_fsum += fa->taps_down[k];
    p.f_C = fa->C;
    p.f_need = p.tile_rows + max_span;
    p.f_fast_sin = fa->fast_sin;
  }
  for (int pl = 0; pl < planes; ++pl) {
    void* xb = pl == 0 ? d->x.d_ptr : d->x.d_lo;
    cuuint64_t dims[3] = {(cuuint64_t)w->x_pitch, (cuuint64_t)d->L, (cuuint64_t)d->B};
    cuuint64_t strides[2] = {(cuuint64_t)w->x_pitch * 2, (cuuint64_t)w->x_pitch * 2 * (cuuint64_t)d->L};
    cuuint32_t box[3] = {(cuuint32_t)UM_KB, (cuuint32_t)p.a_box_rows, 1};
    rc = fused ? BVG_OK : umma_encode_bf16_map(&p.tm_x[pl], xb, 3, dims, strides, box, "activation");
    if (rc != BVG_OK) return rc;
    void* wb = pl == 0 ? w->d_w : w->d_w_lo;
    if (stacked && pl == 1) continue;  // both weight planes live in d_w
    cuuint64_t wdims[2] = {(cuuint64_t)w->cin_pad, (cuuint64_t)w->n_tiles * w->tap_stride * w_rows};
    cuuint64_t wstrides[1] = {(cuuint64_t)w->cin_pad * 2};
    cuuint32_t wbox[2] = {(cuuint32_t)UM_KB, (cuuint32_t)(p.tap_group * w_rows)};
    rc = umma_encode_bf16_map(&p.tm_w[pl], wb, 2, wdims, wstrides, wbox, "weights");
    if (rc != BVG_OK) return rc;
  }

  long long grid = p.total_tiles < sms ? p.total_tiles : sms;
  if (umma_max_ctas > 0 && grid > umma_max_ctas) grid = umma_max_ctas;
  out->grid = (int)grid;
  return BVG_OK;
}

typedef void (*UmmaKernel)(const UmmaParams);

// pick the epilogue specialisation for a descriptor (nullptr-free: falls back to the generic one)
static UmmaKernel select_kernel(const UmmaParams& p) {
  const EpiParams& e = p.epi;
  const bool res = e.res != nullptr, acc = e.acc != nullptr;
  const bool sbf = (res && e.res_dtype == BVG_BF16) || (acc && e.acc_dtype == BVG_BF16);
  const bool mixed = (res && acc && e.res_dtype != e.acc_dtype);
  const bool plain = p.vec_ok && !mixed && (acc || !e.use_div) && (!acc || res) && !e.relu && !e.coldiv;
  if (p.f_x) {  // fused Activation1d producer: the epilogues the fp32 path's resblock convolutions use, else the generic one
    if (plain && !sbf) {
      if (e.out_dtype == BVG_F32 && !res) return conv_umma_kernel<BVG_F32, false, false, false, false, true>;
      if (e.out_dtype == BVG_F32 && res && !acc) return conv_umma_kernel<BVG_F32, false, true, false, false, true>;
      if (e.out_dtype == BVG_F32 && res && acc) return conv_umma_kernel<BVG_F32, false, true, true, false, true>;
      if (e.out_dtype == BVG_SPLIT && res && acc) return conv_umma_kernel<BVG_SPLIT, false, true, true, false, true>;
    }
    return conv_umma_kernel<BVG_F32, false, false, false, true, true>;
  }
  if (plain) {
    if (!sbf) {
      if (e.out_dtype == BVG_F32 && !res) return conv_umma_kernel<BVG_F32, false, false, false, false>;
      if (e.out_dtype == BVG_F32 && res && !acc) return conv_umma_kernel<BVG_F32, false, true, false, false>;
      if (e.out_dtype == BVG_F32 && res && acc) return conv_umma_kernel<BVG_F32, false, true, true, false>;
      if (e.out_dtype == BVG_SPLIT && !res) return conv_umma_kernel<BVG_SPLIT, false, false, false, false>;
      if (e.out_dtype == BVG_SPLIT && res && acc) return conv_umma_kernel<BVG_SPLIT, false, true, true, false>;
    }
    if (e.out_dtype == BVG_BF16 && !res) return conv_umma_kernel<BVG_BF16, true, false, false, false>;
    if (sbf && e.out_dtype == BVG_BF16 && res && !acc && e.res_dtype == BVG_BF16) return conv_umma_kernel<BVG_BF16, true, true, false, false>;
    if (sbf && e.out_dtype == BVG_BF16 && res && acc && e.res_dtype == BVG_BF16) return conv_umma_kernel<BVG_BF16, true, true, true, false>;
  }
  return conv_umma_kernel<BVG_F32, false, false, false, true>;
}

int conv_umma_launch(const UmmaLaunch* l, cudaStream_t st) {
  if (l->grid <= 0) return BVG_OK;
  UmmaKernel k = select_kernel(l->p);
  // opt in to > 48 KB of dynamic shared memory once per (device, specialisation): the attribute is per device
  if (first_use_on_device(reinterpret_cast<const void*>(k)))
    BVG_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, UM_SMEM_LIMIT));
  BVG_CHECK_CUDA(launch_k(k, dim3(l->grid), dim3(l->p.f_x ? UM_THREADS_FUSED : 128 + 32 * l->p.epi_warps), l->smem, st, l->p));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "conv_umma_kernel launch");
  return BVG_OK;
}

size_t umma_launch_size() { return sizeof(UmmaLaunch); }

int conv_umma_forward(const bvg_conv_desc* d, cudaStream_t st) {
  UmmaLaunch l;
  int rc = conv_umma_prepare(d, &l);
  if (rc != BVG_OK) return rc;
  return conv_umma_launch(&l, st);
}

}  // namespace bvg
