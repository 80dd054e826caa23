// Shared pieces of the tensor-core Activation1d kernels (amp_mma.cu, amp_stream.cu): parameters, PTX
// wrappers (ldmatrix / stmatrix / mma.sync), operand splitting, the banded-Toeplitz coefficients and
// the snake nonlinearity.  Index arithmetic pinned by tests/amp_mma_emulation.py.
#pragma once
#include <cuda_fp16.h>

#include <type_traits>

#include "common.cuh"

namespace bvg {

constexpr int AM_NB = 8;                // z-tiles (8 steps each) per staged time tile: 64 steps
constexpr int AM_ROWS = 8 * AM_NB + 16;  // staged x rows per tile (16 = FIR halo of the 17 s-blocks a tile touches)

struct AmpMmaParams {
  const void* x;
  void* y;
  void* y_lo;
  const float* a;
  const float* invb;
  float gu[12];  // 2 * upsample taps
  float fd[12];  // downsample taps
  int B, L, C;
  int n_cg;           // CTAs along channels
  int n_ct;           // CTAs along time
  int n_tiles;        // time tiles per batch item
  int tiles_per_cta;  // consecutive time tiles one CTA walks through (the s fragment is carried across)
  int m_last;         // last z-tile index (z-tiles run from -1)
};

namespace amm {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void stmatrix_x4_trans(uint32_t addr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
  asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
               : "memory");
}
__device__ __forceinline__ void stmatrix_x2_trans(uint32_t addr, uint32_t r0, uint32_t r1) {
  asm volatile("stmatrix.sync.aligned.m8n8.x2.trans.shared.b16 [%0], {%1, %2};" ::"r"(addr), "r"(r0), "r"(r1) : "memory");
}
// The MMA wrappers are plain (non-volatile) asm: an mma.sync is a pure function of its register operands, and volatile
// would pin every one of them in program order against the ldmatrix / stmatrix / cp.async statements, i.e. serialise
// consecutive time blocks inside a warp (ncu r01: 60 % issue-active, stalls = dependent-issue waits of the
// ldmatrix -> HMMA -> MUFU -> HMMA chain).
// d += A(16x16, row) * B(16x8, col), bf16 operands, fp32 accumulate
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// same shape with fp16 operands
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// (a, b) -> packed fp16 pair, round to nearest, saturating at +-65504 instead of overflowing to inf
__device__ __forceinline__ uint32_t pack_f16x2_sat(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %2, %1;" : "=r"(r) : "f"(a), "f"(b));  // first source -> upper half
  return r;
}
// d += A(16x8, row) * B(8x8, col), tf32 operands (fp32 registers), fp32 accumulate
__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

// (a, b) -> packed bf16 pair of the rounded values and of the rounding residuals
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = pack_bf16x2(a, b);
  const float ah = __uint_as_float(hi << 16), bh = __uint_as_float(hi & 0xffff0000u);
  lo = pack_bf16x2(a - ah, b - bh);
}

// Toeplitz coefficients (tests/amp_mma_emulation.py: up_coeff / down_coeff)
__device__ __forceinline__ float up_coeff(const AmpMmaParams& p, int k, int j) {
  const int mm = (j & 1) ? ((j - 1) >> 1) + 6 - k : (j >> 1) + 5 - k;
  if (mm < 0 || mm > 5) return 0.f;
  return p.gu[(j & 1) ? 2 * mm : 2 * mm + 1];
}
__device__ __forceinline__ float down_coeff(const AmpMmaParams& p, int kk, int n) {
  const int t = kk - 2 * n - 1;
  return (t < 0 || t > 11) ? 0.f : p.fd[t];
}

template <bool FAST_SIN>
__device__ __forceinline__ float snake(float u, float apar, float invb) {
  float s;
  if constexpr (FAST_SIN) {
    s = __sinf(u * apar);
  } else {
    const float t = u * apar;                               // half-turns (apar = a / pi)
    const float k = (t + 12582912.0f) - 12582912.0f;        // rint for |t| < 2^22
    s = __sinf((t - k) * 3.14159265358979f);                // sin^2 has period pi; argument in [-pi/2, pi/2]
  }
  return fmaf(invb, s * s, u);
}

// snake on two values of the same channel (an accumulator register pair) with Blackwell's two-lane fp32
// instructions: mul / mul / fma are one instruction each for the pair (the FMA pipe has room in these kernels,
// the issue slots do not); same operation order as snake<> above, so the results are identical.
template <bool FAST_SIN>
__device__ __forceinline__ void snake_pair(float& u0, float& u1, float apar, float invb) {
  unsigned long long u, a, b, arg, ss, out;
  asm("mov.b64 %0, {%1, %2};" : "=l"(u) : "f"(u0), "f"(u1));
  asm("mov.b64 %0, {%1, %1};" : "=l"(a) : "f"(apar));
  asm("mov.b64 %0, {%1, %1};" : "=l"(b) : "f"(invb));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(arg) : "l"(u), "l"(a));
  float t0, t1;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(t0), "=f"(t1) : "l"(arg));
  if constexpr (!FAST_SIN) {
    t0 = (t0 - ((t0 + 12582912.0f) - 12582912.0f)) * 3.14159265358979f;
    t1 = (t1 - ((t1 + 12582912.0f) - 12582912.0f)) * 3.14159265358979f;
  }
  const float s0 = __sinf(t0), s1 = __sinf(t1);
  asm("mov.b64 %0, {%1, %2};" : "=l"(ss) : "f"(s0), "f"(s1));
  asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(ss) : "l"(ss));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(out) : "l"(b), "l"(ss), "l"(u));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(u0), "=f"(u1) : "l"(out));
}

}  // namespace amm

}  // namespace bvg
