// K-A: fused anti-aliased activation (Activation1d) for channels-last [B, L, C] tensors.
//
// Replaces reference modules/bigvgan.py:251-256 (UpSample1d :278-287 -> Snake/SnakeBeta
// :84-95 / :146-159 -> DownSample1d :304-307 -> LowPassFilter1d :224-231): one pass over HBM,
// the 2x-rate signal never leaves registers.
//
// Math (pinned by tests/golden/activation1d.npz):
//   u[2i]   = sum_m g[2m+1] x[clamp(i+2-m)],  u[2i+1] = sum_m g[2m] x[clamp(i+3-m)],  g = 2*f_up
//   s[j]    = u[j] + invb * sin^2(a * u[j])     (evaluated as u - (invb/2) cos(2 a u) + invb/2, amp_p2.cuh)
//   z[t]    = sum_k f_down[k] * s[clamp(2t + k - 5, 0, 2L-1)]
// Two different replicate clamps: on x (1x rate) and on the *activated* s (2x rate).
//
// Work decomposition: one thread owns VEC adjacent channels and a run of TT consecutive time
// steps.  It slides along time keeping the 6-row x window and the 12-value s window in
// registers; every step consumes one new x row and produces two new s values and one output
// row.  Steps are processed in blocks of 6 with two register sets used ping-pong, so the
// windows rotate with compile-time indices and no register moves.  A warp covers 32*VEC
// adjacent channels of one row => every global access is a fully coalesced row segment.
#include "amp_p2.cuh"

namespace bvg {

struct AmpParams {
  const void* x;
  void* y;
  void* y_lo;
  const float* a;
  const float* invb;
  float gu[12];  // 2 * upsample taps (the ratio gain of bigvgan.py:282 folded in; exact, power of two)
  float fd[12];  // downsample taps
  float fsum;    // their sum (fp32, tap 0 .. 11): what a constant comes out of the low-pass as
  int B, L, C;
  int CG;       // channel groups = C / VEC
  int nchunks;  // time chunks per batch item
  int nblk2;    // chunk length = 12 * nblk2 steps
  long long total_threads;
};

template <bool IN_BF16, int VEC>
__device__ __forceinline__ void load_row(const void* base, long long off, float (&v)[VEC]) {
  if constexpr (!IN_BF16) {
    const float* p = reinterpret_cast<const float*>(base) + off;
    if constexpr (VEC == 4) {
      float4 t = __ldg(reinterpret_cast<const float4*>(p));
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else if constexpr (VEC == 2) {
      float2 t = __ldg(reinterpret_cast<const float2*>(p));
      v[0] = t.x; v[1] = t.y;
    } else {
      v[0] = __ldg(p);
    }
  } else {
    const uint16_t* p = reinterpret_cast<const uint16_t*>(base) + off;
    if constexpr (VEC == 4) {
      uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
      unpack_bf16x2(t.x, v[0], v[1]);
      unpack_bf16x2(t.y, v[2], v[3]);
    } else if constexpr (VEC == 2) {
      uint32_t t = __ldg(reinterpret_cast<const uint32_t*>(p));
      unpack_bf16x2(t, v[0], v[1]);
    } else {
      v[0] = bf16_bits_to_float(__ldg(p));
    }
  }
}

template <int VEC>
__device__ __forceinline__ void store_bf16_row(void* base, long long off, const float (&v)[VEC]) {
  uint16_t* p = reinterpret_cast<uint16_t*>(base) + off;
  if constexpr (VEC == 4) {
    uint2 t;
    t.x = pack_bf16x2(v[0], v[1]);
    t.y = pack_bf16x2(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = t;
  } else if constexpr (VEC == 2) {
    *reinterpret_cast<uint32_t*>(p) = pack_bf16x2(v[0], v[1]);
  } else {
    *p = (uint16_t)float_to_bf16_bits(v[0]);
  }
}

template <int OUT_MODE, int VEC>
__device__ __forceinline__ void store_row(void* y, void* y_lo, long long off, const float (&v)[VEC]) {
  if constexpr (OUT_MODE == BVG_F32) {
    float* p = reinterpret_cast<float*>(y) + off;
    if constexpr (VEC == 4) {
      *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    } else if constexpr (VEC == 2) {
      *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
    } else {
      *p = v[0];
    }
  } else if constexpr (OUT_MODE == BVG_BF16) {
    store_bf16_row<VEC>(y, off, v);
  } else {
    float hi[VEC], lo[VEC];
#pragma unroll
    for (int c = 0; c < VEC; ++c) split_bf16(v[c], hi[c], lo[c]);
    store_bf16_row<VEC>(y, off, hi);
    store_bf16_row<VEC>(y_lo, off, lo);
  }
}

template <bool IN_BF16, int OUT_MODE, int VEC, bool FAST_SIN, bool STORE>
__device__ __forceinline__ void amp_block6(const AmpParams& p, const float (&xa)[6][VEC], float (&xb)[6][VEC],
                                           const float (&sa)[12][VEC], float (&sb)[12][VEC], int tau0, long long in_base,
                                           long long out_base, const float (&apar)[VEC], const float (&invb)[VEC], const float (&zc)[VEC]) {
  const int L = p.L;
  // rows x[tau0+6 .. tau0+11] feed the *next* block; issue them first so they are in flight
  // while this block computes.
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    int r = min(max(tau0 + 6 + j, 0), L - 1);
    load_row<IN_BF16, VEC>(p.x, in_base + (long long)r * p.C, xb[j]);
  }
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    float pa[VEC], pb[VEC];
#pragma unroll
    for (int c = 0; c < VEC; ++c) pa[c] = pb[c] = 0.f;
#pragma unroll
    for (int m = 0; m < 6; ++m) {
      const int w = j + 5 - m;  // window slot of x[tau + 5 - m]
#pragma unroll
      for (int c = 0; c < VEC; ++c) {
        const float xv = (w < 6) ? xa[w < 6 ? w : 0][c] : xb[w >= 6 ? w - 6 : 0][c];
        pb[c] = fmaf(p.gu[2 * m + 1], xv, pb[c]);  // s[2 tau + 6]: even phase, i = tau + 3
        pa[c] = fmaf(p.gu[2 * m], xv, pa[c]);      // s[2 tau + 5]: odd phase,  i = tau + 2
      }
    }
#pragma unroll
    for (int c = 0; c < VEC; ++c) {
      pa[c] = snake_one<FAST_SIN>(pa[c], apar[c], invb[c]);
      pb[c] = snake_one<FAST_SIN>(pb[c], apar[c], invb[c]);
    }
    const int tau = tau0 + j;
    if (tau >= L - 3) {  // right replicate clamp of the activated signal: s[j > 2L-1] = s[2L-1]
#pragma unroll
      for (int c = 0; c < VEC; ++c) {
        const float prev = (j == 0) ? sa[11][c] : sb[j == 0 ? 0 : 2 * j - 1][c];
        if (tau >= L - 2) pa[c] = prev;
        pb[c] = pa[c];
      }
    }
#pragma unroll
    for (int c = 0; c < VEC; ++c) {
      sb[2 * j][c] = pa[c];
      sb[2 * j + 1][c] = pb[c];
    }
    if constexpr (STORE) {
      float z[VEC];
#pragma unroll
      for (int c = 0; c < VEC; ++c) z[c] = zc[c];
#pragma unroll
      for (int k = 0; k < 12; ++k) {
        const int i = 2 * j + 2 + k;  // slot in (sa ++ sb) of s[2 tau - 5 + k]
#pragma unroll
        for (int c = 0; c < VEC; ++c) {
          const float sv = (i < 12) ? sa[i < 12 ? i : 0][c] : sb[i >= 12 ? i - 12 : 0][c];
          z[c] = fmaf(p.fd[k], sv, z[c]);
        }
      }
      if (tau < L) store_row<OUT_MODE, VEC>(p.y, p.y_lo, out_base + (long long)tau * p.C, z);
    }
  }
}

template <bool IN_BF16, int OUT_MODE, int VEC, bool FAST_SIN>
__global__ void __launch_bounds__(128) amp_kernel(const __grid_constant__ AmpParams p) {
  pdl_trigger();  // programmatic dependent launch (common.cuh): no global access before the wait
  pdl_wait();
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= p.total_threads) return;
  const int cg = (int)(tid % p.CG);
  const long long rest = tid / p.CG;
  const int chunk = (int)(rest % p.nchunks);
  const int b = (int)(rest / p.nchunks);
  const int TT = 12 * p.nblk2;
  const int t0 = chunk * TT;
  const int L = p.L;
  const long long base = (long long)b * L * p.C + (long long)cg * VEC;

  float apar[VEC], invb[VEC], zc[VEC];  // cosine-form constants (amp_p2.cuh): 2a | a/pi, -invb/2, (invb/2) sum(taps)
#pragma unroll
  for (int c = 0; c < VEC; ++c) {
    const float ib = __ldg(p.invb + cg * VEC + c);
    apar[c] = snake_apar2<FAST_SIN>(__ldg(p.a + cg * VEC + c));
    invb[c] = snake_hbn(ib);
    zc[c] = snake_zc(ib, p.fsum);
  }

  float xa[6][VEC], xb[6][VEC], sa[12][VEC], sb[12][VEC];
#pragma unroll
  for (int k = 0; k < 12; ++k)
#pragma unroll
    for (int c = 0; c < VEC; ++c) sa[k][c] = 0.f;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    int r = min(max(t0 - 6 + k, 0), L - 1);
    load_row<IN_BF16, VEC>(p.x, base + (long long)r * p.C, xa[k]);
  }
  // warm-up block (tau = t0-6 .. t0-1): fills sb = s[2 t0 - 7 .. 2 t0 + 4], no output
  amp_block6<IN_BF16, OUT_MODE, VEC, FAST_SIN, false>(p, xa, xb, sa, sb, t0 - 6, base, base, apar, invb, zc);
  if (t0 == 0) {  // left replicate clamp of the activated signal: s[j < 0] = s[0] (slot 7)
#pragma unroll
    for (int k = 0; k < 7; ++k)
#pragma unroll
      for (int c = 0; c < VEC; ++c) sb[k][c] = sb[7][c];
  }
  int t = t0;
  for (int i = 0; i < p.nblk2; ++i) {
    if (t >= L) break;
    amp_block6<IN_BF16, OUT_MODE, VEC, FAST_SIN, true>(p, xb, xa, sb, sa, t, base, base, apar, invb, zc);
    t += 6;
    if (t >= L) break;
    amp_block6<IN_BF16, OUT_MODE, VEC, FAST_SIN, true>(p, xa, xb, sa, sb, t, base, base, apar, invb, zc);
    t += 6;
  }
}

// ------------------------------------------------------------------------------------------------
// Packed variant for two channels per thread: Blackwell's FFMA2 / FMUL2 / FADD2 (PTX fma.rn.f32x2 ...)
// execute two fp32 lanes per instruction and take a scalar (broadcast) operand, so every FIR tap is
// ONE instruction for both channels.  ncu on the scalar kernel: 85 % issue-active at 60 instructions
// per element -- the issue slots, not HBM, bound it; this halves the FIR and most of the snake.
// Each lane of an f32x2 operation is an IEEE fp32 operation, so results are bit-identical to the
// scalar kernel above.
// ------------------------------------------------------------------------------------------------
template <bool IN_BF16>
__device__ __forceinline__ P2 load_row2(const void* base, long long off) {
  // plain (coherent) loads, not ld.global.nc: a load that may alias the output rows stays behind the stores
  // before it, which bounds how far ptxas sinks the stores of a block (and the registers they hold)
  if constexpr (!IN_BF16) {
    const float2 t = *reinterpret_cast<const float2*>(reinterpret_cast<const float*>(base) + off);
    return pk2(t.x, t.y);
  } else {
    const uint32_t t = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint16_t*>(base) + off);
    return pk2(__uint_as_float(t << 16), __uint_as_float(t & 0xffff0000u));
  }
}

// INTERIOR: the block touches no sequence end (rows tau0+6 .. tau0+11 exist, tau0+5 < L-3): no clamps, no
// selects, no store predicate.  CT: compile-time channel count (0 = runtime p.C): with it every row of the
// block is an immediate offset from one pointer.  Together they remove ~33 of the 85 instructions per step
// that ncu counted as address arithmetic and edge selects (profiles/r01_ncu_summary_v7.md -> _v13.md).
template <bool IN_BF16, int OUT_MODE, bool FAST_SIN, bool STORE, bool INTERIOR, int CT>
__device__ __forceinline__ void amp_block6_p2(const AmpParams& p, const P2 (&xa)[6], P2 (&xb)[6], const P2 (&sa)[12], P2 (&sb)[12], int tau0,
                                              long long base, P2 apar, P2 invb, P2 zc, bool prefetch = false) {
  const int L = p.L;
  const int Cc = CT ? CT : p.C;
  if constexpr (INTERIOR) {
    const long long row0 = base + (long long)(tau0 + 6) * Cc;
#pragma unroll
    for (int j = 0; j < 6; ++j) xb[j] = load_row2<IN_BF16>(p.x, row0 + j * Cc);
    // L1 prefetch of the rows two blocks ahead (12 rows past the ones just requested): holds no registers, and
    // the demand loads two blocks later hit on chip.  ncu before: 5.9 stall cycles per issue on the global loads
    // (long scoreboard) at 24 resident warps; measured in-program +5..17 % (gpurun_out/ab_pf.txt; 6, 18, 24, 36 rows
    // or an L2-only prefetch are behind).  A chunk prefetches up to 12 rows past its own end (the next chunk's rows,
    // or the next batch item's); `prefetch` is false where that would leave the tensor.
    if (prefetch) {
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const char* a = reinterpret_cast<const char*>(p.x) + (row0 + (long long)(12 + j) * Cc) * (IN_BF16 ? 2 : 4);
        asm volatile("prefetch.global.L1 [%0];" ::"l"(a));
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const int r = min(max(tau0 + 6 + j, 0), L - 1);
      xb[j] = load_row2<IN_BF16>(p.x, base + (long long)r * Cc);
    }
  }
  const long long out0 = base + (long long)tau0 * Cc;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    P2 pa = 0ull, pb = 0ull;  // +0.0f in both lanes
#pragma unroll
    for (int m = 0; m < 6; ++m) {
      const int w = j + 5 - m;  // window slot of x[tau + 5 - m]
      const P2 xv = (w < 6) ? xa[w < 6 ? w : 0] : xb[w >= 6 ? w - 6 : 0];
      pb = fma2(pk2(p.gu[2 * m + 1], p.gu[2 * m + 1]), xv, pb);  // s[2 tau + 6]: even phase, i = tau + 3
      pa = fma2(pk2(p.gu[2 * m], p.gu[2 * m]), xv, pa);          // s[2 tau + 5]: odd phase,  i = tau + 2
    }
    pa = snake_two<FAST_SIN>(pa, apar, invb);
    pb = snake_two<FAST_SIN>(pb, apar, invb);
    const int tau = tau0 + j;
    if constexpr (!INTERIOR) {
      if (tau >= L - 3) {  // right replicate clamp of the activated signal: s[j > 2L-1] = s[2L-1]
        const P2 prev = (j == 0) ? sa[11] : sb[j == 0 ? 0 : 2 * j - 1];
        if (tau >= L - 2) pa = prev;
        pb = pa;
      }
    }
    sb[2 * j] = pa;
    sb[2 * j + 1] = pb;
    if constexpr (STORE) {
      P2 z = zc;
#pragma unroll
      for (int k = 0; k < 12; ++k) {
        const int i = 2 * j + 2 + k;  // slot in (sa ++ sb) of s[2 tau - 5 + k]
        const P2 sv = (i < 12) ? sa[i < 12 ? i : 0] : sb[i >= 12 ? i - 12 : 0];
        z = fma2(pk2(p.fd[k], p.fd[k]), sv, z);
      }
      if (INTERIOR || tau < L) {
        float zz[2];
        upk2(z, zz[0], zz[1]);
        if constexpr (OUT_MODE == BVG_SPLIT) {
          // hi = bf16(z) for both channels in one F2FP, widened back with a shift and a mask, lo = z - hi in
          // one packed operation: the values of split_bf16 in 5 instructions instead of 8
          const uint32_t hi = pack_bf16x2(zz[0], zz[1]);
          const P2 lo = fma2(pk2(__uint_as_float(hi << 16), __uint_as_float(hi & 0xffff0000u)), pk2(-1.0f, -1.0f), z);
          float l0, l1;
          upk2(lo, l0, l1);
          *reinterpret_cast<uint32_t*>(reinterpret_cast<uint16_t*>(p.y) + out0 + j * Cc) = hi;
          *reinterpret_cast<uint32_t*>(reinterpret_cast<uint16_t*>(p.y_lo) + out0 + j * Cc) = pack_bf16x2(l0, l1);
        } else {
          store_row<OUT_MODE, 2>(p.y, p.y_lo, out0 + j * Cc, zz);
        }
      }
    }
  }
}

// __launch_bounds__(128, 6) = 80 registers: the clamp-free loop wants ~96 (5 resident CTAs) and takes a few
// spills at 80; measured on B200 in-program (gpurun_out/ab_p2ct.txt) 80 beats 96 by 5-10 % for C >= 96 and
// ties below, 72 and 64 spill heavily.  The bf16 variants spill at 64 (2.4x slower).
template <bool IN_BF16, int OUT_MODE, bool FAST_SIN, int CT>
__global__ void __launch_bounds__(128, 6) amp_kernel_p2(const __grid_constant__ AmpParams p) {
  pdl_trigger();  // programmatic dependent launch (common.cuh): no global access before the wait
  pdl_wait();
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= p.total_threads) return;
  const int Cc = CT ? CT : p.C;
  const int CG = CT ? CT / 2 : p.CG;
  const int cg = (int)(tid % CG);
  const long long rest = tid / CG;
  const int chunk = (int)(rest % p.nchunks);
  const int b = (int)(rest / p.nchunks);
  const int TT = 12 * p.nblk2;
  const int t0 = chunk * TT;
  const int L = p.L;
  const long long base = (long long)b * L * Cc + (long long)cg * 2;

  // cosine-form constants (amp_p2.cuh): 2a | a/pi, -invb/2, (invb/2) sum(taps)
  const float ib0 = __ldg(p.invb + cg * 2), ib1 = __ldg(p.invb + cg * 2 + 1);
  const P2 apar = pk2(snake_apar2<FAST_SIN>(__ldg(p.a + cg * 2)), snake_apar2<FAST_SIN>(__ldg(p.a + cg * 2 + 1)));
  const P2 invb = pk2(snake_hbn(ib0), snake_hbn(ib1));
  const P2 zc = pk2(snake_zc(ib0, p.fsum), snake_zc(ib1, p.fsum));

  P2 xa[6], xb[6], sa[12], sb[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) sa[k] = 0ull;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const int r = min(max(t0 - 6 + k, 0), L - 1);
    xa[k] = load_row2<IN_BF16>(p.x, base + (long long)r * Cc);
  }
  // A chunk whose last block stays clear of the sequence end (rows up to t0 + TT + 5 exist) runs the
  // clamp-free blocks only; the last chunk of a batch item takes the general path.  Two separate loops:
  // a per-block choice inside one loop costs registers.
  if (t0 + TT + 5 <= L - 1) {
    // warm-up block (tau = t0-6 .. t0-1): fills sb = s[2 t0 - 7 .. 2 t0 + 4], no output
    const bool pf = b + 1 < p.B || t0 + TT + 17 <= L - 1;  // every prefetched row lies inside the tensor
    amp_block6_p2<IN_BF16, OUT_MODE, FAST_SIN, false, true, CT>(p, xa, xb, sa, sb, t0 - 6, base, apar, invb, zc, pf);
    if (t0 == 0) {  // left replicate clamp of the activated signal: s[j < 0] = s[0] (slot 7)
#pragma unroll
      for (int k = 0; k < 7; ++k) sb[k] = sb[7];
    }
    int t = t0;
    for (int i = 0; i < p.nblk2; ++i) {
      amp_block6_p2<IN_BF16, OUT_MODE, FAST_SIN, true, true, CT>(p, xb, xa, sb, sa, t, base, apar, invb, zc, pf);
      amp_block6_p2<IN_BF16, OUT_MODE, FAST_SIN, true, true, CT>(p, xa, xb, sa, sb, t + 6, base, apar, invb, zc, pf);
      t += 12;
    }
    return;
  }
  amp_block6_p2<IN_BF16, OUT_MODE, FAST_SIN, false, false, CT>(p, xa, xb, sa, sb, t0 - 6, base, apar, invb, zc);
  if (t0 == 0) {
#pragma unroll
    for (int k = 0; k < 7; ++k) sb[k] = sb[7];
  }
  int t = t0;
  for (int i = 0; i < p.nblk2; ++i) {
    if (t >= L) break;
    amp_block6_p2<IN_BF16, OUT_MODE, FAST_SIN, true, false, CT>(p, xb, xa, sb, sa, t, base, apar, invb, zc);
    t += 6;
    if (t >= L) break;
    amp_block6_p2<IN_BF16, OUT_MODE, FAST_SIN, true, false, CT>(p, xa, xb, sa, sb, t, base, apar, invb, zc);
    t += 6;
  }
}


template <bool IN_BF16, int OUT_MODE, bool FAST_SIN, int CT>
static cudaError_t launch_amp_p2_ct(const AmpParams& p, cudaStream_t st) {
  const int threads = 128;
  const long long blocks = ceil_div_ll(p.total_threads, threads);
  // Same shared-memory carve-out as the convolution kernel (max shared): an SM can only host kernels of
  // two streams at once when they agree on the L1 / shared split, and this kernel streams through L2 anyway.
  if (first_use_on_device(reinterpret_cast<const void*>(amp_kernel_p2<IN_BF16, OUT_MODE, FAST_SIN, CT>)))
    cudaFuncSetAttribute(amp_kernel_p2<IN_BF16, OUT_MODE, FAST_SIN, CT>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  return launch_k(amp_kernel_p2<IN_BF16, OUT_MODE, FAST_SIN, CT>, dim3((unsigned)blocks), dim3(threads), 0, st, p);
}


// the generator's own operand format (F32 -> SPLIT) gets one instantiation per channel count of the
// repo / v2 generators; everything else runs with the channel count as a kernel parameter
template <bool IN_BF16, int OUT_MODE, bool FAST_SIN>
static cudaError_t launch_amp_p2_sin(const AmpParams& p, cudaStream_t st, bool amp_ct_enable) {
  if constexpr (!IN_BF16 && OUT_MODE == BVG_SPLIT) {
    if (amp_ct_enable) switch (p.C) {
      case 24: return launch_amp_p2_ct<IN_BF16, OUT_MODE, FAST_SIN, 24>(p, st);
      case 48: return launch_amp_p2_ct<IN_BF16, OUT_MODE, FAST_SIN, 48>(p, st);
      case 96: return launch_amp_p2_ct<IN_BF16, OUT_MODE, FAST_SIN, 96>(p, st);
      case 192: return launch_amp_p2_ct<IN_BF16, OUT_MODE, FAST_SIN, 192>(p, st);
      case 384: return launch_amp_p2_ct<IN_BF16, OUT_MODE, FAST_SIN, 384>(p, st);
      case 768: return launch_amp_p2_ct<IN_BF16, OUT_MODE, FAST_SIN, 768>(p, st);
      default: break;
    }
  }
  if constexpr (!IN_BF16 && OUT_MODE == BVG_F32) {
    // activation_post of the repo and v2 generators (F32 -> F32 at the last stage's 24 channels): 0.41 -> ~0.2 ms per forward
    if (amp_ct_enable && p.C == 24) return launch_amp_p2_ct<IN_BF16, OUT_MODE, FAST_SIN, 24>(p, st);
  }
  return launch_amp_p2_ct<IN_BF16, OUT_MODE, FAST_SIN, 0>(p, st);
}

template <bool IN_BF16, int OUT_MODE>
static cudaError_t launch_amp_p2(const AmpParams& p, bool fast, cudaStream_t st, bool ct) {
  return fast ? launch_amp_p2_sin<IN_BF16, OUT_MODE, true>(p, st, ct) : launch_amp_p2_sin<IN_BF16, OUT_MODE, false>(p, st, ct);
}

static cudaError_t launch_amp_packed(const AmpParams& p, bool in_bf16, int out_mode, bool fast, cudaStream_t st, bool ct) {
  if (in_bf16) {
    if (out_mode == BVG_F32) return launch_amp_p2<true, BVG_F32>(p, fast, st, ct);
    if (out_mode == BVG_BF16) return launch_amp_p2<true, BVG_BF16>(p, fast, st, ct);
    return launch_amp_p2<true, BVG_SPLIT>(p, fast, st, ct);
  }
  if (out_mode == BVG_F32) return launch_amp_p2<false, BVG_F32>(p, fast, st, ct);
  if (out_mode == BVG_BF16) return launch_amp_p2<false, BVG_BF16>(p, fast, st, ct);
  return launch_amp_p2<false, BVG_SPLIT>(p, fast, st, ct);
}

template <bool IN_BF16, int OUT_MODE, int VEC, bool FAST_SIN>
static cudaError_t launch_amp(const AmpParams& p, cudaStream_t st) {
  const int threads = 128;
  const long long blocks = ceil_div_ll(p.total_threads, threads);
  return launch_k(amp_kernel<IN_BF16, OUT_MODE, VEC, FAST_SIN>, dim3((unsigned)blocks), dim3(threads), 0, st, p);
}

template <bool IN_BF16, int OUT_MODE, int VEC>
static cudaError_t launch_amp_sin(const AmpParams& p, bool fast, cudaStream_t st) {
  return fast ? launch_amp<IN_BF16, OUT_MODE, VEC, true>(p, st) : launch_amp<IN_BF16, OUT_MODE, VEC, false>(p, st);
}

template <int VEC>
static cudaError_t launch_amp_vec(const AmpParams& p, bool in_bf16, int out_mode, bool fast, cudaStream_t st) {
  if (in_bf16) {
    if (out_mode == BVG_F32) return launch_amp_sin<true, BVG_F32, VEC>(p, fast, st);
    if (out_mode == BVG_BF16) return launch_amp_sin<true, BVG_BF16, VEC>(p, fast, st);
    return launch_amp_sin<true, BVG_SPLIT, VEC>(p, fast, st);
  }
  if (out_mode == BVG_F32) return launch_amp_sin<false, BVG_F32, VEC>(p, fast, st);
  if (out_mode == BVG_BF16) return launch_amp_sin<false, BVG_BF16, VEC>(p, fast, st);
  return launch_amp_sin<false, BVG_SPLIT, VEC>(p, fast, st);
}

bool amp_mma_supported(const bvg_amp_desc* d);               // amp_mma.cu: tensor-core FIR variant
int amp_mma_forward(const bvg_amp_desc* d, cudaStream_t st);
bool amp_stream_supported(const bvg_amp_desc* d);            // amp_stream.cu: per-warp streaming variant (F32 -> SPLIT)
int amp_stream_forward(const bvg_amp_desc* d, cudaStream_t st);


int amp_forward(const bvg_amp_desc* d, cudaStream_t st) {
  BVG_REQUIRE(d != nullptr, "amp: null descriptor");
  BVG_REQUIRE(d->B > 0 && d->L > 0 && d->C > 0, "amp: bad shape B=%d L=%d C=%d", d->B, d->L, d->C);
  BVG_REQUIRE(d->x.d_ptr && d->y.d_ptr && d->d_a && d->d_invb, "amp: null pointer");
  BVG_REQUIRE(d->x.dtype == BVG_F32 || d->x.dtype == BVG_BF16, "amp: input must be F32 or BF16");
  BVG_REQUIRE(d->y.dtype >= BVG_F32 && d->y.dtype <= BVG_SPLIT, "amp: bad output dtype");
  BVG_REQUIRE(d->y.dtype != BVG_SPLIT || d->y.d_lo, "amp: SPLIT output needs a lo plane");
  BVG_REQUIRE((long long)d->L * d->C < (1ll << 31), "amp: L*C too large for one batch item");

  // the two operand formats of the generator (F32 -> SPLIT, BF16 -> BF16) run the FIRs on the
  // tensor cores; every other combination stays on the FFMA kernel below
  const bvg_tuning T = tune_of(d->tune);
  if (amp_stream_supported(d)) return amp_stream_forward(d, st);
  if (amp_mma_supported(d)) return amp_mma_forward(d, st);

  // two channels per thread: measured 15-25 % faster than four on B200 (64 vs 164 registers ->
  // 2.7x the resident warps; profiles/r01_amp_sweep.txt)
  int vec = (d->C % 2 == 0) ? 2 : 1;
  if (T.amp_vec && d->C % T.amp_vec == 0) vec = T.amp_vec;

  AmpParams p;
  p.x = d->x.d_ptr;
  p.y = d->y.d_ptr;
  p.y_lo = d->y.d_lo;
  p.a = d->d_a;
  p.invb = d->d_invb;
  for (int k = 0; k < 12; ++k) {
    p.gu[k] = 2.0f * d->taps_up[k];
    p.fd[k] = d->taps_down[k];
  }
  p.fsum = 0.f;
  for (int k = 0; k < 12; ++k) p.fsum += d->taps_down[k];
  p.B = d->B;
  p.L = d->L;
  p.C = d->C;
  p.CG = d->C / vec;
  // chunk length: as long as possible (the 6 warm-up steps of every chunk are recomputed work:
  // 6 % at 96 steps, 25 % at 24) while still giving every SM a few rounds of resident warps
  int nblk2 = 8;
  const long long want = 148ll * 512 * 3;
  while (nblk2 > 1 && (long long)d->B * p.CG * ceil_div(d->L, 12 * nblk2) < want) nblk2 >>= 1;
  if (T.amp_chunk > 0) nblk2 = T.amp_chunk;
  p.nblk2 = nblk2;
  p.nchunks = ceil_div(d->L, 12 * nblk2);
  p.total_threads = (long long)d->B * p.nchunks * p.CG;
  BVG_REQUIRE(ceil_div_ll(p.total_threads, 128) < (1ll << 31), "amp: grid too large");

  cudaError_t e;
  const bool in_bf16 = d->x.dtype == BVG_BF16;
  if (vec == 4)
    e = launch_amp_vec<4>(p, in_bf16, d->y.dtype, d->fast_sin != 0, st);
  else if (vec == 2 && T.amp_packed)
    e = launch_amp_packed(p, in_bf16, d->y.dtype, d->fast_sin != 0, st, T.amp_ct != 0);
  else if (vec == 2)
    e = launch_amp_vec<2>(p, in_bf16, d->y.dtype, d->fast_sin != 0, st);
  else
    e = launch_amp_vec<1>(p, in_bf16, d->y.dtype, d->fast_sin != 0, st);
  if (e != cudaSuccess) return cuda_fail(e, "amp_kernel launch");
  return BVG_OK;
}

}  // namespace bvg
