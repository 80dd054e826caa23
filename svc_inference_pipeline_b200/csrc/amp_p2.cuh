// Packed fp32 pairs (fma.rn.f32x2 and friends) and the snake nonlinearity on them: shared by the FFMA2
// Activation1d kernel (amp_kernel.cu) and the Activation1d producer fused into the convolution kernel
// (conv_umma.cu).
#pragma once
#include "common.cuh"

namespace bvg {

// Snake in cosine form.  s = u + invb sin^2(a u) = (u - (invb / 2) cos(2 a u)) + invb / 2: the kernels carry
// s' = s - invb / 2 through the low-pass and start its accumulator at (invb / 2) * sum(taps) instead of 0 (replicate
// padding and the FIR both pass a constant through), which saves the squaring: one multiply, one MUFU.COS and one
// FMA per sample.  ncu: the fp32-path kernel is FMA-pipe bound, this is 4 of its 66 pipe cycles per step.
//   apar2 = 2 a (FAST: MUFU on the raw product, phase error ~|2 a u| * 1e-7, like the reference's own fp32
//           rounding of a u) or a / pi (turns of the angle 2 a u, reduced exactly to [-1/2, 1/2] first, which keeps
//           MUFU.COS in its most accurate range);  hbn = -invb / 2.
template <bool FAST_SIN>
__device__ __forceinline__ float snake_one(float u, float apar2, float hbn) {
  float c;
  if constexpr (FAST_SIN) {
    c = __cosf(u * apar2);
  } else {
    float t = u * apar2;                          // turns
    float k = (t + 12582912.0f) - 12582912.0f;    // rint for |t| < 2^22
    float r = t - k;                              // [-0.5, 0.5]
    c = __cosf(r * 6.28318530717958648f);
  }
  return fmaf(c, hbn, u);
}

// per-channel constants of the cosine form from a = exp(alpha) and invb = 1 / (exp(beta) + eps)
template <bool FAST_SIN>
__device__ __forceinline__ float snake_apar2(float a) { return FAST_SIN ? 2.0f * a : a * 0.318309886183790672f; }
__device__ __forceinline__ float snake_hbn(float invb) { return -0.5f * invb; }
__device__ __forceinline__ float snake_zc(float invb, float tap_sum) { return (0.5f * invb) * tap_sum; }

typedef unsigned long long P2;  // two packed fp32 (channel c in the low half, c+1 in the high half)

__device__ __forceinline__ P2 pk2(float a, float b) {
  P2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk2(P2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ P2 fma2(P2 a, P2 b, P2 c) {
  P2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ P2 mul2(P2 a, P2 b) {
  P2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ P2 add2(P2 a, P2 b) {
  P2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// same operation order as snake_one, two channels at a time
template <bool FAST_SIN>
__device__ __forceinline__ P2 snake_two(P2 u, P2 apar2, P2 hbn) {
  P2 arg = mul2(u, apar2);
  if constexpr (!FAST_SIN) {
    const P2 magic = pk2(12582912.0f, 12582912.0f), nmagic = pk2(-12582912.0f, -12582912.0f);
    const P2 k = add2(add2(arg, magic), nmagic);              // rint for |t| < 2^22
    arg = fma2(k, pk2(-1.0f, -1.0f), arg);                     // [-0.5, 0.5] turns (exact)
    arg = mul2(arg, pk2(6.28318530717958648f, 6.28318530717958648f));
  }
  float a0, a1;
  upk2(arg, a0, a1);
  return fma2(pk2(__cosf(a0), __cosf(a1)), hbn, u);
}

}  // namespace bvg
