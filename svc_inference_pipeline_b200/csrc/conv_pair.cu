// K-C / K-T on a CTA PAIR: the wide convolutions (N tile >= 128 columns) as a tap GEMM issued with
// tcgen05.mma.cta_group::2 -- two SMs of one TPC work on one 256-row tile (reference layers: the same as conv_umma.cu,
// modules/bigvgan.py:319-386/:428-431, :529-537/:602, :547-561/:607; formulation of include/bvg_b200.h).
//
// Why: what bounded the single-CTA kernel on the wide layers was the tensor core's operand traffic from shared
// memory (ncu r02: 86 % tensor-pipe active; a 128 x 128 x 16 MMA needs 8 KB in 65 cycles, ~the 128 B / clk an SM
// delivers, and the TMA writes share that port).  With cta_group::2 one instruction computes D[256, N] from
// A[256, K] and B[N, K]: CTA r of the pair holds rows [128 r, 128 r + 128) of the activation tile and HALF of the weight
// box (rows [r N / 2, (r + 1) N / 2)), so every SM reads half the B bytes per MMA and loads half of every weight box
// from L2, while each weight box still serves 256 output rows.
//
//   * cluster of 2 CTAs (cudaLaunchKernelEx, cluster dimension 2), persistent: pair i walks tiles i, i + #pairs, ...
//   * warp 0 of BOTH CTAs is a TMA producer: its own A super-tile (rows + halo, all planes) and its half of each weight box,
//     loaded with cp.async.bulk.tensor...cta_group::2 so that the transaction bytes of both CTAs complete on the
//     LEADER's (rank 0) full barriers; the leader's producer posts arrive.expect_tx for the pair's total;
//   * warp 1 of the leader issues every MMA; tcgen05.commit...cta_group::2.multicast::cluster releases the smem stages
//     and publishes the accumulators in both CTAs;
//   * warps 4-11 of both CTAs are the epilogue of their own 128 rows (own TMEM lanes); they hand the accumulator stage
//     back on the leader's barrier (the peer arrives remotely, mbarrier.arrive.shared::cluster);
//   * SPLIT operands (fp32 path): three products per K step, hi*hi into the main accumulator columns [0, n) and
//     lo*hi, hi*lo into the correction columns [n, 2 n): the tensor core truncates when it adds into an fp32
//     accumulator (error ~ steps x 2^-24 x |acc|, tools/conv_precision_diag.py), so the large sum takes one addition
//     per K step and the small terms (2^-8 of it) get their own columns; the epilogue adds the two.
// Every mbarrier wait is bounded (trap instead of hang).
#include <cstring>

#include "conv_umma.cuh"

namespace bvg {

constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // shared::cluster address of the same offset in the pair's even (leader) CTA

namespace ptx2 {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA loads of a CTA pair: data into the executing CTA's shared memory, bytes onto the leader's barrier
__device__ __forceinline__ void tma_load_3d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
// D[tmem, 256 rows over the pair] (+)= A * B, bf16 x bf16 -> fp32; issued by the leader CTA only
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all MMAs issued so far have completed) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma2_commit_both(uint32_t bar) {
  const uint16_t both = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(both) : "memory");
}
// arrive on the leader's barrier at this offset (from either CTA)
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerBitMask) : "memory");
}

}  // namespace ptx2

__device__ __forceinline__ uint32_t make_idesc_m256(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

// MMAs of one weight stage with compile-time K steps (straight-line issue: the lone issuing warp executes a few dependent
// uniform-datapath instructions per tcgen05.mma, which matters once an MMA lasts ~65 cycles)
template <int KS>
__device__ __forceinline__ void issue_whi(bool split, uint32_t tmem_main, uint32_t tmem_corr, uint64_t desc_hi, uint32_t a16, uint32_t a_plane16,
                                          uint32_t b16, uint32_t idesc, uint32_t main_acc, uint32_t corr_acc) {
#pragma unroll
  for (int k = 0; k < KS; ++k) {
    const uint64_t bd = desc_hi | (uint64_t)(b16 + 2u * k);
    if (ptx::elect_one()) ptx2::umma2_f16(tmem_main, desc_hi | (uint64_t)(a16 + 2u * k), bd, idesc, k == 0 ? main_acc : 1u);
    if (split) {
      if (ptx::elect_one()) ptx2::umma2_f16(tmem_corr, desc_hi | (uint64_t)(a16 + a_plane16 + 2u * k), bd, idesc, k == 0 ? corr_acc : 1u);
    }
  }
}
template <int KS>
__device__ __forceinline__ void issue_wlo(uint32_t tmem_corr, uint64_t desc_hi, uint32_t a16, uint32_t b16, uint32_t idesc) {
#pragma unroll
  for (int k = 0; k < KS; ++k)
    if (ptx::elect_one()) ptx2::umma2_f16(tmem_corr, desc_hi | (uint64_t)(a16 + 2u * k), desc_hi | (uint64_t)(b16 + 2u * k), idesc, 1u);
}

// OUT / SBF / RES / ACC: the compile-time epilogue choices of conv_umma_kernel (N % 4 == 0 required: vec_ok)
template <int OUT, bool SBF, bool RES, bool ACC>
__global__ void __launch_bounds__(PR_THREADS, 1) conv_pair_kernel(const __grid_constant__ UmmaParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = smem_base;
  const uint32_t b_base = a_base + p.a_stages * p.a_stage_bytes;
  const uint32_t stg_base = b_base + p.b_stages * p.b_stage_bytes;
  const uint32_t bar_base = stg_base + PR_STAGING_BYTES;
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (UM_MAX_A_STAGES + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (2 * UM_MAX_A_STAGES + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (2 * UM_MAX_A_STAGES + UM_MAX_B_STAGES + s); };
  auto t_full = [&](int s) { return bar_base + 8u * (2 * UM_MAX_A_STAGES + 2 * UM_MAX_B_STAGES + s); };
  auto t_empty = [&](int s) { return bar_base + 8u * (2 * UM_MAX_A_STAGES + 2 * UM_MAX_B_STAGES + UM_MAX_T_STAGES + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * UM_MAX_A_STAGES + 2 * UM_MAX_B_STAGES + 2 * UM_MAX_T_STAGES);
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_gen + (tmem_slot - smem_base));

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx2::cluster_ctarank();
  const bool leader = rank == 0;
  const long long pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  pdl_trigger();  // the next launch of the chain may set itself up while this one runs (common.cuh)

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&p.tm_x[0]);
    ptx::prefetch_tmap(&p.tm_w[0]);
    if (p.planes == 2) {
      ptx::prefetch_tmap(&p.tm_x[1]);
      ptx::prefetch_tmap(&p.tm_w[1]);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.a_stages; ++s) {
      ptx::mbar_init(a_full(s), 1);   // leader's producer: arrive.expect_tx for the pair
      ptx::mbar_init(a_empty(s), 1);  // multicast commit
    }
    for (int s = 0; s < p.b_stages; ++s) {
      ptx::mbar_init(b_full(s), 1);
      ptx::mbar_init(b_empty(s), 1);
    }
    for (int s = 0; s < p.t_stages; ++s) {
      ptx::mbar_init(t_full(s), 1);                   // multicast commit
      ptx::mbar_init(t_empty(s), 2 * PR_EPI_WARPS);  // (leader's copy is the one used) every epilogue warp of both CTAs
    }
    ptx::fence_barrier_init();
    ptx::fence_proxy_async();
  }
  if (warp == 2) {
    ptx2::tmem_alloc2(tmem_slot, 512);
    ptx2::tmem_relinquish2();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx2::cluster_sync();  // both CTAs' barriers are initialised before anything arrives on them remotely
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();  // the previous launch's results are visible from here on

  const int planes = p.planes;
  const int half_rows = p.n_tile >> 1;                                   // weight rows this CTA holds per tap
  // SPLIT: the correction products go to their own columns (p.stacked = 1) or, for the single-tile C = 192 layers
  // (K <= 2112: the main accumulator's extra additions cost 5e-6 there, and two accumulator stages matter more than
  // that on their short tiles), into the main accumulator (p.stacked = 0)
  const bool corr_separate = planes == 2 && p.stacked != 0;
  const uint32_t corr_col = corr_separate ? (uint32_t)p.col_stride : 0u;
  const int stage_cols = (corr_separate ? 2 : 1) * p.col_stride;

  if (warp == 0) {
    // ================================ TMA producer (both CTAs) ================================
    int sa = 0, pa = 0, sb = 0, pb = 0;
    const int box_bytes = p.a_box_rows * 128;
    for (long long tile = pair; tile < p.total_tiles; tile += n_pairs) {
      const int nt = (int)(tile % p.n_tiles);
      const long long mt = tile / p.n_tiles;
      const int b = (int)(mt / p.m_tiles_per_item);
      const int t0 = (int)(mt % p.m_tiles_per_item) * p.tile_rows + (int)rank * UM_BM;
      const int ntaps = p.n_taps[nt];
      const int row0 = t0 + p.min_shift[nt];
      for (int cb = 0; cb < p.n_cb; ++cb) {
        ptx::mbar_wait(a_empty(sa), pa ^ 1, p.err_flag, 1);
        if (ptx::elect_one()) {
          if (leader) ptx::mbar_expect_tx(a_full(sa), 2u * (uint32_t)p.a_stage_bytes);
          for (int pl = 0; pl < planes; ++pl)
            for (int bx = 0; bx < p.a_boxes; ++bx)
              ptx2::tma_load_3d_2sm(a_base + sa * p.a_stage_bytes + pl * p.a_plane_bytes + bx * box_bytes, &p.tm_x[pl], a_full(sa), cb * UM_KB,
                                    row0 + bx * p.a_box_rows, b);
        }
        __syncwarp();
        if (++sa == p.a_stages) { sa = 0; pa ^= 1; }
        for (int tap = 0; tap < ntaps; ++tap) {
          for (int wp = 0; wp < planes; ++wp) {
            ptx::mbar_wait(b_empty(sb), pb ^ 1, p.err_flag, 2);
            if (ptx::elect_one()) {
              if (leader) ptx::mbar_expect_tx(b_full(sb), 2u * (uint32_t)p.b_stage_bytes);
              ptx2::tma_load_2d_2sm(b_base + sb * p.b_stage_bytes, &p.tm_w[wp], b_full(sb), cb * UM_KB,
                                    (nt * p.tap_stride + tap) * p.n_tile + (int)rank * half_rows);
            }
            __syncwarp();
            if (++sb == p.b_stages) { sb = 0; pb ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1 && leader) {
    // ================================ MMA issuer (leader only) ================================
    const uint32_t idesc = make_idesc_m256(p.n_tile);
    const uint64_t desc_hi = (uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29)) << 32;  // SBO, version, SWIZZLE_128B
    const uint32_t a_plane16 = (uint32_t)p.a_plane_bytes >> 4;
    const bool split = planes == 2;
    const int ks_last = (p.cin - (p.n_cb - 1) * UM_KB + 15) >> 4;
    int sa = 0, pa = 0, sb = 0, pb = 0, as = 0, ap = 0;
    for (long long tile = pair; tile < p.total_tiles; tile += n_pairs) {
      const int nt = (int)(tile % p.n_tiles);
      const int ntaps = p.n_taps[nt];
      const int min_shift = p.min_shift[nt];
      ptx::mbar_wait(t_empty(as), ap ^ 1, p.err_flag, 3);
      ptx::tc_fence_after();
      const uint32_t tmem_main = tmem_base + (uint32_t)(as * stage_cols);
      const uint32_t tmem_corr = tmem_main + corr_col;
      const int* shifts = p.shift[nt];
      uint32_t main_acc = 0u, corr_acc = corr_separate ? 0u : 1u;  // 0 => the first product overwrites the accumulator
      for (int cb = 0; cb < p.n_cb; ++cb) {
        const int ksteps = cb + 1 < p.n_cb ? 4 : ks_last;
        ptx::mbar_wait(a_full(sa), pa, p.err_flag, 4);
        const uint32_t a_stage16 = (((a_base + sa * p.a_stage_bytes) & 0x3FFFFu) >> 4) | (1u << 16);
        for (int tap = 0; tap < ntaps; ++tap) {
          const uint32_t a16 = a_stage16 + (uint32_t)(shifts[tap] - min_shift) * 8u;  // 128 B per row
          // weights hi: A hi -> main, A lo -> correction
          ptx::mbar_wait(b_full(sb), pb, p.err_flag, 5);
          ptx::tc_fence_after();
          {
            const uint32_t b16 = (((b_base + sb * p.b_stage_bytes) & 0x3FFFFu) >> 4) | (1u << 16);
            if (ksteps == 4) issue_whi<4>(split, tmem_main, tmem_corr, desc_hi, a16, a_plane16, b16, idesc, main_acc, corr_acc);
            else if (ksteps == 2) issue_whi<2>(split, tmem_main, tmem_corr, desc_hi, a16, a_plane16, b16, idesc, main_acc, corr_acc);
            else if (ksteps == 3) issue_whi<3>(split, tmem_main, tmem_corr, desc_hi, a16, a_plane16, b16, idesc, main_acc, corr_acc);
            else issue_whi<1>(split, tmem_main, tmem_corr, desc_hi, a16, a_plane16, b16, idesc, main_acc, corr_acc);
            main_acc = 1u;
            corr_acc = 1u;
          }
          if (ptx::elect_one()) ptx2::umma2_commit_both(b_empty(sb));
          __syncwarp();
          if (++sb == p.b_stages) { sb = 0; pb ^= 1; }
          if (split) {  // weights lo: A hi -> correction
            ptx::mbar_wait(b_full(sb), pb, p.err_flag, 5);
            ptx::tc_fence_after();
            const uint32_t b16 = (((b_base + sb * p.b_stage_bytes) & 0x3FFFFu) >> 4) | (1u << 16);
            if (ksteps == 4) issue_wlo<4>(tmem_corr, desc_hi, a16, b16, idesc);
            else if (ksteps == 2) issue_wlo<2>(tmem_corr, desc_hi, a16, b16, idesc);
            else if (ksteps == 3) issue_wlo<3>(tmem_corr, desc_hi, a16, b16, idesc);
            else issue_wlo<1>(tmem_corr, desc_hi, a16, b16, idesc);
            if (ptx::elect_one()) ptx2::umma2_commit_both(b_empty(sb));
            __syncwarp();
            if (++sb == p.b_stages) { sb = 0; pb ^= 1; }
          }
        }
        if (ptx::elect_one()) ptx2::umma2_commit_both(a_empty(sa));
        __syncwarp();
        if (++sa == p.a_stages) { sa = 0; pa ^= 1; }
      }
      if (ptx::elect_one()) ptx2::umma2_commit_both(t_full(as));
      __syncwarp();
      if (++as == p.t_stages) { as = 0; ap ^= 1; }
    }
  } else if (warp >= 4 && warp < 4 + PR_EPI_WARPS) {
    // ================================ epilogue (both CTAs, own 128 rows) ================================
    const int e = warp - 4;
    const int q = e & 3;
    const int grp = e >> 2;
    const uint32_t stg = stg_base + (uint32_t)e * 2048u;
    const int n_chunks = p.n_tile >> 4;
    const int rrow = lane >> 2;
    const int g = lane & 3;
    const int n_my = (n_chunks - grp + 1) >> 1;  // chunks grp, grp + 2, ...
    const int N = p.epi.N;
    int as = 0, ap = 0;
    for (long long tile = pair; tile < p.total_tiles; tile += n_pairs) {
      const int nt = (int)(tile % p.n_tiles);
      const long long mt = tile / p.n_tiles;
      const int b = (int)(mt / p.m_tiles_per_item);
      const int t0 = (int)(mt % p.m_tiles_per_item) * p.tile_rows + (int)rank * UM_BM;
      const long long row_base = (long long)b * p.L;
      const uint32_t tmem_q = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * stage_cols);
      const int tbase = t0 + q * 32;
      for (int pb2 = 0; pb2 < n_my; pb2 += 2) {
        uint4 raw[2][4];
        [[maybe_unused]] float ac[2][4][4];
        int it_c0[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          it_c0[u] = (grp + 2 * (pb2 + u)) << 4;
          if (ACC) {
            // running sum (acc_in) of both chunks of this iteration, issued together and -- for the first pair of a tile --
            // before the accumulator-ready wait, like the residuals below (ncu on the DiffSVC dilated layer: tensor pipe
            // 14 % active, the epilogue warps stalled on exactly these loads when they were issued chunk by chunk)
            const int n0 = nt * p.n_tile + it_c0[u] + 4 * g;
            const bool ok = pb2 + u < n_my && n0 < N;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
#pragma unroll
              for (int j = 0; j < 4; ++j) ac[u][i][j] = 0.f;
              const int t = tbase + rrow + 8 * i;
              if (ok && t < p.L) epi_load4(p.epi.acc, SBF ? BVG_BF16 : BVG_F32, (row_base + t) * N + n0, ac[u][i]);
            }
          }
          if (RES) {
            const int n0 = nt * p.n_tile + it_c0[u] + 4 * g;
            const bool ok = pb2 + u < n_my && n0 < N;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              raw[u][i] = make_uint4(0u, 0u, 0u, 0u);
              const int t = tbase + rrow + 8 * i;
              if (ok && t < p.L) {
                const long long off = (row_base + t) * N + n0;
                if (!SBF) {
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
          const int c0 = it_c0[u];
          if (tbase >= p.L) continue;  // whole 32-row slab past the end of the sequence (warp-uniform)
          const int n0 = nt * p.n_tile + c0 + 4 * g;
          const long long off0 = (row_base + tbase + rrow) * N + n0;  // row i adds 8 * i * N
          uint32_t r[16];
          ptx::tmem_ld16(tmem_q + (uint32_t)c0, r);
          if (corr_separate) {
            uint32_t r2[16];
            ptx::tmem_ld16(tmem_q + corr_col + (uint32_t)c0, r2);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
          }
          ptx::tmem_ld_wait();
          const uint32_t wsw = (uint32_t)((lane >> 1) & 3);
#pragma unroll
          for (int gg = 0; gg < 4; ++gg) {
            const uint32_t addr = stg + (uint32_t)lane * 64u + (((uint32_t)gg ^ wsw) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(r[4 * gg]), "r"(r[4 * gg + 1]), "r"(r[4 * gg + 2]), "r"(r[4 * gg + 3]) : "memory");
          }
          __syncwarp();
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
          const float4 bias = *reinterpret_cast<const float4*>(p.epi.bias + n0);
          // per-channel divisor: one load per chunk, next to the bias (inside the row loop it was a dependent global load
          // in front of every division: ncu's top long-scoreboard stall of the DiffSVC output projection)
          const float4 cd = p.epi.coldiv ? __ldg(reinterpret_cast<const float4*>(p.epi.coldiv + n0)) : make_float4(1.f, 1.f, 1.f, 1.f);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (tbase + rrow + 8 * i >= p.L) continue;
            v[i][0] += bias.x; v[i][1] += bias.y; v[i][2] += bias.z; v[i][3] += bias.w;
            if (RES) {
              float rs[4];
              if (!SBF) {
                rs[0] = __uint_as_float(raw[u][i].x); rs[1] = __uint_as_float(raw[u][i].y);
                rs[2] = __uint_as_float(raw[u][i].z); rs[3] = __uint_as_float(raw[u][i].w);
              } else {
                unpack_bf16x2(raw[u][i].x, rs[0], rs[1]);
                unpack_bf16x2(raw[u][i].y, rs[2], rs[3]);
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) v[i][j] += rs[j];
            }
            if (ACC) {
#pragma unroll
              for (int j = 0; j < 4; ++j) v[i][j] += ac[u][i][j];
            }
            if (p.epi.use_div) {
#pragma unroll
              for (int j = 0; j < 4; ++j) v[i][j] = __fdiv_rn(v[i][j], p.epi.div);
            }
            if (p.epi.coldiv) {
              v[i][0] = __fdiv_rn(v[i][0], cd.x); v[i][1] = __fdiv_rn(v[i][1], cd.y);
              v[i][2] = __fdiv_rn(v[i][2], cd.z); v[i][3] = __fdiv_rn(v[i][3], cd.w);
            }
            if (p.epi.relu) {
#pragma unroll
              for (int j = 0; j < 4; ++j) v[i][j] = fmaxf(v[i][j], 0.f);
            }
            const long long off = off0 + (long long)(8 * i) * N;
            if (OUT == BVG_F32) {
              *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.epi.out) + off) = make_float4(v[i][0], v[i][1], v[i][2], v[i][3]);
            } else if (OUT == BVG_BF16) {
              epi_store_bf16x4(p.epi.out, off, v[i]);
            } else {
              float hi[4], lo[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) split_bf16(v[i][j], hi[j], lo[j]);
              epi_store_bf16x4(p.epi.out, off, hi);
              epi_store_bf16x4(p.epi.out_lo, off, lo);
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
      if (lane == 0) ptx2::mbar_arrive_leader(t_empty(as));
      if (++as == p.t_stages) { as = 0; ap ^= 1; }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx2::cluster_sync();  // the peer's MMAs-in-flight and remote arrivals are done before either CTA tears down
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx2::tmem_dealloc2(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct PairLaunch {
  UmmaParams p;
  int grid;
  size_t smem;
};

// Whether this convolution runs on the CTA-pair kernel: tensor-core weights that are not stacked, an N tile of at
// least 128 columns (a multiple of 32, so that each CTA's half is a whole number of 16-row groups), no fused
// Activation1d producer, and an epilogue the specialised kernels cover.
bool conv_pair_eligible(const bvg_conv_desc* d) {
  const bvg_conv_weights* w = d->w;
  if (tune_of(d->tune).umma_pair == 0) return false;
  if (!w || w->backend != BVG_UMMA || d->pre_amp || w->split == 2) return false;
  const int pair_min = tune_of(d->tune).umma_pair_min > 0 ? tune_of(d->tune).umma_pair_min : 128;
  if (w->n_tile < pair_min || w->n_tile % 32 != 0 || w->n_total % 4 != 0) return false;
  if (w->split == 2) return false;
  if (w->split && w->n_tiles == 1 && tune_of(d->tune).umma_pair == 2) return false;  // A/B: C = 192 SPLIT layers on the single-CTA kernel
  const bool res = d->res.d_ptr != nullptr, acc = d->acc_in.d_ptr != nullptr;
  if (res && acc && d->res.dtype != d->acc_in.dtype) return false;
  const bool sbf = (res && d->res.dtype == BVG_BF16) || (acc && d->acc_in.dtype == BVG_BF16);
  const int o = d->out.dtype;
  if (o == BVG_BF16) return (!res && !acc) || (sbf && res);  // bf16 path: residual stream in bf16
  if (sbf) return false;
  if (o == BVG_F32) return true;
  return o == BVG_SPLIT && ((!res && !acc) || (res && acc));
}

int conv_pair_prepare(const bvg_conv_desc* d, PairLaunch* out) {
  const bvg_conv_weights* w = d->w;
  BVG_REQUIRE(conv_pair_eligible(d), "conv_pair: descriptor is not eligible for the CTA-pair kernel");
  BVG_REQUIRE(d->x.dtype == BVG_BF16 || d->x.dtype == BVG_SPLIT, "conv_pair: input must be BF16 or SPLIT");
  const int planes = d->x.dtype == BVG_SPLIT ? 2 : 1;
  BVG_REQUIRE(planes == 1 || (w->split == 1 && w->d_w_lo && d->x.d_lo), "conv_pair: SPLIT input needs split-packed weights and a lo plane");
  BVG_REQUIRE(d->x.d_ptr && w->d_w && d->B > 0 && d->L > 0, "conv_pair: bad descriptor");
  BVG_REQUIRE(w->n_tiles <= BVG_MAX_NTILES && w->x_pitch % 8 == 0, "conv_pair: bad geometry");
  BVG_REQUIRE(((uintptr_t)d->x.d_ptr & 15) == 0 && ((uintptr_t)w->d_w & 15) == 0, "conv_pair: operands must be 16-byte aligned");

  UmmaParams& p = out->p;
  memset(&p, 0, sizeof(p));
  int rc = fill_epilogue(d, p.epi);
  if (rc != BVG_OK) return rc;
  p.planes = planes;
  p.B = d->B;
  p.L = d->L;
  p.N = w->n_total;
  p.n_tile = w->n_tile;
  p.n_tiles = w->n_tiles;
  p.tap_stride = w->tap_stride;
  p.n_cb = w->cin_pad / UM_KB;
  p.cin = w->cin;
  p.vec_ok = 1;
  p.mb = 1;
  p.tap_group = 1;
  int max_span = 0;
  for (int t = 0; t < w->n_tiles; ++t) {
    BVG_REQUIRE(w->n_taps[t] > 0 && w->n_taps[t] <= BVG_MAX_TAPS, "conv_pair: bad tap count");
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
  p.col_stride = (w->n_tile + 31) / 32 * 32;
  // SPLIT: main + correction columns; a layer with a single N tile (C = 192: short tiles, K <= 2112) keeps ONE
  // accumulator for all three products so that two stages fit and its epilogue overlaps the next tile (with main +
  // correction on one stage it measured 12.0 ms per forward against 10.5 ms on the single-CTA kernel)
  const bool corr_separate = planes == 2 && !(w->n_tiles == 1 && 2 * p.col_stride > 256);
  p.stacked = corr_separate ? 1 : 0;
  const int acc_cols = (corr_separate ? 2 : 1) * p.col_stride;
  BVG_REQUIRE(acc_cols <= 512, "conv_pair: accumulators do not fit the tensor memory");
  p.t_stages = 512 / acc_cols;
  if (p.t_stages > UM_MAX_T_STAGES) p.t_stages = UM_MAX_T_STAGES;
  p.tile_rows = 2 * UM_BM;  // per pair
  p.m_tiles_per_item = ceil_div(d->L, p.tile_rows);
  p.total_tiles = (long long)d->B * p.m_tiles_per_item * w->n_tiles;
  // shared memory of one CTA: 2 activation stages (128 rows + halo, all planes), its half of the weight boxes
  const int rows = UM_BM + max_span;
  const int nb = (rows + 255) / 256;
  p.a_box_rows = (((rows + nb - 1) / nb) + 7) / 8 * 8;
  p.a_boxes = nb;
  p.a_plane_bytes = p.a_boxes * p.a_box_rows * 128;
  p.a_stage_bytes = p.a_plane_bytes * planes;
  p.a_stages = 2;
  if (3 * p.a_stage_bytes + 4 * (w->n_tile / 2) * 128 + PR_STAGING_BYTES + UM_BAR_BYTES + 1024 <= UM_SMEM_LIMIT) p.a_stages = 3;
  p.b_stage_bytes = (w->n_tile / 2) * 128;
  int smem_cap = UM_SMEM_LIMIT;
  const int cap_kb = tune_of(d->tune).umma_pair_smem_kb;
  if (cap_kb > 0 && cap_kb * 1024 < smem_cap) smem_cap = cap_kb * 1024;
  if (3 * p.a_stage_bytes + 4 * p.b_stage_bytes + PR_STAGING_BYTES + UM_BAR_BYTES + 1024 > smem_cap) p.a_stages = 2;
  int bs = (smem_cap - 1024 - UM_BAR_BYTES - PR_STAGING_BYTES - p.a_stages * p.a_stage_bytes) / p.b_stage_bytes;
  if (bs > UM_MAX_B_STAGES) bs = UM_MAX_B_STAGES;
  BVG_REQUIRE(bs >= 2, "conv_pair: tile does not fit in shared memory (span %d, n_tile %d, planes %d)", max_span, w->n_tile, planes);
  p.b_stages = bs;
  out->smem = 1024 + (size_t)p.a_stages * p.a_stage_bytes + (size_t)bs * p.b_stage_bytes + PR_STAGING_BYTES + UM_BAR_BYTES;
  if (out->smem < 120 * 1024 && cap_kb == 0) out->smem = 120 * 1024;  // one CTA per SM (each pair owns all 512 TMEM columns of its two SMs)

  for (int pl = 0; pl < planes; ++pl) {
    void* xb = pl == 0 ? d->x.d_ptr : d->x.d_lo;
    cuuint64_t dims[3] = {(cuuint64_t)w->x_pitch, (cuuint64_t)d->L, (cuuint64_t)d->B};
    cuuint64_t strides[2] = {(cuuint64_t)w->x_pitch * 2, (cuuint64_t)w->x_pitch * 2 * (cuuint64_t)d->L};
    cuuint32_t box[3] = {(cuuint32_t)UM_KB, (cuuint32_t)p.a_box_rows, 1};
    rc = umma_encode_bf16_map(&p.tm_x[pl], xb, 3, dims, strides, box, "activation (pair)");
    if (rc != BVG_OK) return rc;
    void* wb = pl == 0 ? w->d_w : w->d_w_lo;
    cuuint64_t wdims[2] = {(cuuint64_t)w->cin_pad, (cuuint64_t)w->n_tiles * w->tap_stride * w->n_tile};
    cuuint64_t wstrides[1] = {(cuuint64_t)w->cin_pad * 2};
    cuuint32_t wbox[2] = {(cuuint32_t)UM_KB, (cuuint32_t)(w->n_tile / 2)};
    rc = umma_encode_bf16_map(&p.tm_w[pl], wb, 2, wdims, wstrides, wbox, "weights (pair)");
    if (rc != BVG_OK) return rc;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long pairs = p.total_tiles < sms / 2 ? p.total_tiles : sms / 2;
  const int cap = tune_of(d->tune).umma_max_ctas;
  if (cap > 0 && pairs > cap / 2) pairs = cap / 2 > 0 ? cap / 2 : 1;
  out->grid = (int)(2 * pairs);
  return BVG_OK;
}

typedef void (*PairKernel)(const UmmaParams);

static PairKernel select_pair_kernel(const UmmaParams& p) {
  const EpiParams& e = p.epi;
  const bool res = e.res != nullptr, acc = e.acc != nullptr;
  const bool sbf = (res && e.res_dtype == BVG_BF16) || (acc && e.acc_dtype == BVG_BF16);
  if (!sbf) {
    if (e.out_dtype == BVG_F32 && !res && !acc) return conv_pair_kernel<BVG_F32, false, false, false>;
    if (e.out_dtype == BVG_F32 && res && !acc) return conv_pair_kernel<BVG_F32, false, true, false>;
    if (e.out_dtype == BVG_F32 && res && acc) return conv_pair_kernel<BVG_F32, false, true, true>;
    if (e.out_dtype == BVG_F32 && !res && acc) return conv_pair_kernel<BVG_F32, false, false, true>;
    if (e.out_dtype == BVG_SPLIT && !res && !acc) return conv_pair_kernel<BVG_SPLIT, false, false, false>;
    if (e.out_dtype == BVG_SPLIT && res && acc) return conv_pair_kernel<BVG_SPLIT, false, true, true>;
  }
  if (e.out_dtype == BVG_BF16 && !res && !acc) return conv_pair_kernel<BVG_BF16, true, false, false>;
  if (e.out_dtype == BVG_BF16 && res && !acc) return conv_pair_kernel<BVG_BF16, true, true, false>;
  if (e.out_dtype == BVG_BF16 && res && acc) return conv_pair_kernel<BVG_BF16, true, true, true>;
  return nullptr;
}

int conv_pair_launch(const PairLaunch* l, cudaStream_t st) {
  if (l->grid <= 0) return BVG_OK;
  PairKernel k = select_pair_kernel(l->p);
  BVG_REQUIRE(k != nullptr, "conv_pair: no kernel for this epilogue");
  if (first_use_on_device(reinterpret_cast<const void*>(k)))
    BVG_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, UM_SMEM_LIMIT));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)l->grid, 1, 1);
  cfg.blockDim = dim3(PR_THREADS, 1, 1);
  cfg.dynamicSmemBytes = l->smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_mode() ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, k, l->p);
  if (e != cudaSuccess) return cuda_fail(e, "conv_pair_kernel launch");
  return BVG_OK;
}

size_t pair_launch_size() { return sizeof(PairLaunch); }

int conv_pair_forward(const bvg_conv_desc* d, cudaStream_t st) {
  PairLaunch l;
  int rc = conv_pair_prepare(d, &l);
  if (rc != BVG_OK) return rc;
  return conv_pair_launch(&l, st);
}

}  // namespace bvg
