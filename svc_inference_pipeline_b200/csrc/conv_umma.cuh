// Shared pieces of the tcgen05 tap-GEMM kernels (conv_umma.cu: one CTA per tile; conv_pair.cu: a CTA pair per tile,
// tcgen05.mma.cta_group::2): tile constants, the kernel parameter block, PTX wrappers, UMMA descriptors.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "epilogue.cuh"

namespace bvg {

constexpr int UM_BM = 128;          // rows per M block (TMEM lanes)
constexpr int UM_KB = 64;           // channels per K slice (one 128-byte swizzle row of bf16)
constexpr int UM_EPI_WARPS = 8;      // epilogue warps of the single-CTA kernel (two per TMEM lane quarter) ...
constexpr int UM_EPI_WARPS_MAX = 16; // ... or four per quarter for bf16 operands (UmmaParams.epi_warps): the narrow layers' epilogue is
                                     // a latency-bound stream of global loads / stores, and with one MMA pass per product it is what
                                     // bounds them (measured per forward, bf16 path: C = 96 / 48 / 24 classes 3.90 / 3.22 / 3.57 ->
                                     // 3.36 / 2.79 / 3.13 ms; with the three passes of the fp32 path it changes nothing)
constexpr int PR_EPI_WARPS = 8;              // epilogue warps per CTA of the pair kernel
constexpr int PR_THREADS = 128 + 32 * PR_EPI_WARPS;
constexpr int PR_STAGING_BYTES = PR_EPI_WARPS * 32 * 64;
constexpr int UM_THREADS = 128 + 32 * UM_EPI_WARPS;
constexpr int UM_AMP_WARPS = 8;      // fused mode: Activation1d producer warps after the epilogue warps
constexpr int UM_THREADS_FUSED = UM_THREADS + 32 * UM_AMP_WARPS;
constexpr int UM_MAX_A_STAGES = 4;   // activation super-tile stages (2..4, chosen per launch)
constexpr int UM_MAX_B_STAGES = 8;
constexpr int UM_MAX_T_STAGES = 4;
constexpr int UM_MAX_MB = 8;
constexpr int UM_STAGING_BYTES = UM_EPI_WARPS * 32 * 64;  // 32 rows x 16 fp32 per epilogue warp
constexpr int UM_SMEM_LIMIT = 227 * 1024;

struct UmmaParams {
  CUtensorMap tm_x[2];  // activation planes (hi, lo)
  CUtensorMap tm_w[2];  // weight planes (hi, lo)
  EpiParams epi;
  int planes;           // 1 (BF16) or 2 (SPLIT)
  int stacked;          // SPLIT with both weight planes stacked along N (rows [0,n) hi, [n,2n) lo per tap)
  int B, L, N;
  int n_tile, n_tiles, tap_stride;
  int n_cb;             // Cin slices
  int cin;              // true input channels (K steps of the last slice)
  int mb;               // M blocks (of 128 rows) per tile: they share every weight box
  int tile_rows;        // mb * 128
  int m_tiles_per_item;
  long long total_tiles;
  int a_box_rows, a_boxes;  // the A super-tile is loaded as a_boxes TMA boxes of a_box_rows rows
  int col_stride;       // TMEM columns per accumulator (n_tile rounded up to 32)
  int t_stages;         // accumulator stages (each mb * col_stride columns)
  int a_stages;
  int b_stages;
  int tap_group;        // taps per weight stage (one TMA box of tap_group * n_tile rows)
  int a_stage_bytes;    // all planes
  int a_plane_bytes;
  int b_stage_bytes;
  int vec_ok;
  int epi_warps;        // epilogue warps of this launch (single-CTA kernel: 8 or 16; block = 128 + 32 * epi_warps threads)
  int n_taps[BVG_MAX_NTILES];
  int min_shift[BVG_MAX_NTILES];
  int shift[BVG_MAX_NTILES][BVG_MAX_TAPS];
  int* err_flag;        // optional device word set before a watchdog trap
  // fused Activation1d producer (bvg_conv_desc.pre_amp): the A operand is computed here from the Activation1d's
  // fp32 input instead of being loaded by TMA
  const float* f_x;     // [B, L, f_C] fp32, channels-last
  const float* f_a;
  const float* f_invb;
  float f_gu[12], f_fd[12], f_fsum;
  int f_C;
  int f_need;           // rows of an A stage the MMAs read: tile_rows + max shift - min shift
  int f_fast_sin;
};

namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one elected lane of a fully converged warp (CUTLASS' elect_one_sync)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// bounded wait: ~2 s at 2 GHz, then flag + trap
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err_flag, int code) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) {
      if (err_flag) atomicExch(err_flag, code);
      __threadfence_system();
      __trap();
    }
  }
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> fp32, cta_group::1
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t addr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(addr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t addr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(addr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

}  // namespace ptx

// UMMA shared-memory matrix descriptor, K-major operand, 128-byte swizzle:
//   [0,14) start address >> 4   [16,30) leading byte offset >> 4 (unused for swizzled K-major)
//   [32,46) stride byte offset >> 4 = 1024 B between 8-row groups   [46,48) version = 1
//   [49,52) base offset   [61,64) layout type = 2 (SWIZZLE_128B)
// The 128B swizzle is a function of the absolute shared-memory address bits (measured on B200:
// profiles/r01_probe_first_contact.log, umma_halo0 vs umma_halo1), so a start address advanced by
// whole 128-byte rows -- not a multiple of the 8-row atom -- with base_offset = 0 addresses the
// rows TMA wrote.  That is what lets every tap of a dilated conv read one shared A halo tile.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// instruction descriptor: D=F32, A=B=BF16, both K-major, N>>3 at [17,23), M>>4 at [24,29)
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(UM_BM >> 4) << 24);
}

// host: 128B-swizzled bf16 tensor map (conv_umma.cu)
int umma_encode_bf16_map(CUtensorMap* map, void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box,
                         const char* what);

constexpr int UM_BAR_BYTES = 8 * (2 * UM_MAX_A_STAGES + 2 * UM_MAX_B_STAGES + 2 * UM_MAX_T_STAGES) + 16;

}  // namespace bvg
