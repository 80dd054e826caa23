// Packed fp32 pairs (fma.rn.f32x2 and friends) and the snake nonlinearity on them: shared by the FFMA2
// Activation1d kernel (amp_kernel.cu) and the Activation1d producer fused into the convolution kernel
// (conv_umma.cu).
#pragma once
#include "common.cuh"

namespace bvg {

// sin(a*u): FAST = MUFU on the raw product (phase error ~|a*u| * 1e-7, like the reference's own
// fp32 rounding of a*u); otherwise reduce exactly to [-pi/2, pi/2] first (sin^2 has period pi),
// which keeps MUFU.SIN in its most accurate range (abs err 2^-21).  apar = a (FAST) or a/pi.
template <bool FAST_SIN>
__device__ __forceinline__ float snake_one(float u, float apar, float invb) {
  float s;
  if constexpr (FAST_SIN) {
    s = __sinf(u * apar);
  } else {
    float t = u * apar;               // half-turns
    float k = (t + 12582912.0f) - 12582912.0f;  // rint for |t| < 2^22
    float r = t - k;                  // [-0.5, 0.5]
    s = __sinf(r * 3.14159265358979f);
  }
  return fmaf(invb, s * s, u);
}

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

// same operation order as snake_one, two channels at a time (apar = a or a/pi, see snake_one)
template <bool FAST_SIN>
__device__ __forceinline__ P2 snake_two(P2 u, P2 apar, P2 invb) {
  P2 arg = mul2(u, apar);
  if constexpr (!FAST_SIN) {
    const P2 magic = pk2(12582912.0f, 12582912.0f), nmagic = pk2(-12582912.0f, -12582912.0f);
    const P2 k = add2(add2(arg, magic), nmagic);              // rint for |t| < 2^22
    arg = fma2(k, pk2(-1.0f, -1.0f), arg);                     // [-0.5, 0.5] half-turns (exact)
    arg = mul2(arg, pk2(3.14159265358979f, 3.14159265358979f));
  }
  float a0, a1;
  upk2(arg, a0, a1);
  const P2 s = pk2(__sinf(a0), __sinf(a1));
  return fma2(invb, mul2(s, s), u);
}

}  // namespace bvg
