// Shared epilogue of the conv ("tap GEMM") kernels:
//   v = acc + bias[n] (+ res[row, n]) (+ acc_in[row, n]);  v /= div;  store as F32 | BF16 | SPLIT.
// Order of the additions follows the reference: conv output (sum + bias), then "xt + x"
// (modules/bigvgan.py:431), then "xs += ..." (:613-614), then "xs / num_kernels" (:615, a true
// division, kept as a division so the fp32 path differs from torch only by summation order).
#pragma once
#include "common.cuh"

namespace bvg {

struct EpiParams {
  void* out;
  void* out_lo;
  const void* res;
  const void* acc;
  const float* bias;
  const float* coldiv;  // per-column divisor (then use_div = 0)
  float div;
  int out_dtype, res_dtype, acc_dtype;
  int use_div;
  int relu;  // max(v, 0) after everything else
  int N;  // row pitch of out / res / acc (elements)
};

__device__ __forceinline__ void epi_load4(const void* base, int dtype, long long off, float (&v)[4]) {
  if (dtype == BVG_F32) {
    float4 t = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
    uint2 t = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(base) + off);
    unpack_bf16x2(t.x, v[0], v[1]);
    unpack_bf16x2(t.y, v[2], v[3]);
  }
}

__device__ __forceinline__ float epi_load1(const void* base, int dtype, long long off) {
  if (dtype == BVG_F32) return reinterpret_cast<const float*>(base)[off];
  return bf16_bits_to_float(reinterpret_cast<const uint16_t*>(base)[off]);
}

__device__ __forceinline__ void epi_store_bf16x4(void* base, long long off, const float (&v)[4]) {
  uint2 t;
  t.x = pack_bf16x2(v[0], v[1]);
  t.y = pack_bf16x2(v[2], v[3]);
  *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(base) + off) = t;
}

// store 4 finished values at element offset `off` in the output format
__device__ __forceinline__ void epi_store4(const EpiParams& e, long long off, const float (&v)[4]) {
  if (e.out_dtype == BVG_F32) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out) + off) = make_float4(v[0], v[1], v[2], v[3]);
  } else if (e.out_dtype == BVG_BF16) {
    epi_store_bf16x4(e.out, off, v);
  } else {
    float hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_bf16(v[i], hi[i], lo[i]);
    epi_store_bf16x4(e.out, off, hi);
    epi_store_bf16x4(e.out_lo, off, lo);
  }
}

// 4 consecutive outputs n0..n0+3 of one row; requires N % 4 == 0 and n0 % 4 == 0.
__device__ __forceinline__ void epilogue4(const EpiParams& e, long long row, int n0, float (&v)[4]) {
  const long long off = row * e.N + n0;
  const float4 b = *reinterpret_cast<const float4*>(e.bias + n0);
  v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
  if (e.res) {
    float r[4];
    epi_load4(e.res, e.res_dtype, off, r);
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] += r[i];
  }
  if (e.acc) {
    float r[4];
    epi_load4(e.acc, e.acc_dtype, off, r);
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] += r[i];
  }
  if (e.use_div) {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __fdiv_rn(v[i], e.div);
  }
  if (e.coldiv) {
    const float4 cd = *reinterpret_cast<const float4*>(e.coldiv + n0);
    v[0] = __fdiv_rn(v[0], cd.x); v[1] = __fdiv_rn(v[1], cd.y); v[2] = __fdiv_rn(v[2], cd.z); v[3] = __fdiv_rn(v[3], cd.w);
  }
  if (e.relu) {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = fmaxf(v[i], 0.f);
  }
  if (e.out_dtype == BVG_F32) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out) + off) = make_float4(v[0], v[1], v[2], v[3]);
  } else if (e.out_dtype == BVG_BF16) {
    epi_store_bf16x4(e.out, off, v);
  } else {
    float hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_bf16(v[i], hi[i], lo[i]);
    epi_store_bf16x4(e.out, off, hi);
    epi_store_bf16x4(e.out_lo, off, lo);
  }
}

// scalar fallback for ragged N
__device__ __forceinline__ void epilogue1(const EpiParams& e, long long row, int n, float v) {
  const long long off = row * e.N + n;
  v += e.bias[n];
  if (e.res) v += epi_load1(e.res, e.res_dtype, off);
  if (e.acc) v += epi_load1(e.acc, e.acc_dtype, off);
  if (e.use_div) v = __fdiv_rn(v, e.div);
  if (e.coldiv) v = __fdiv_rn(v, e.coldiv[n]);
  if (e.relu) v = fmaxf(v, 0.f);
  if (e.out_dtype == BVG_F32) {
    reinterpret_cast<float*>(e.out)[off] = v;
  } else if (e.out_dtype == BVG_BF16) {
    reinterpret_cast<uint16_t*>(e.out)[off] = (uint16_t)float_to_bf16_bits(v);
  } else {
    float hi, lo;
    split_bf16(v, hi, lo);
    reinterpret_cast<uint16_t*>(e.out)[off] = (uint16_t)float_to_bf16_bits(hi);
    reinterpret_cast<uint16_t*>(e.out_lo)[off] = (uint16_t)float_to_bf16_bits(lo);
  }
}

inline int fill_epilogue(const bvg_conv_desc* d, EpiParams& e) {
  e.out = d->out.d_ptr;
  e.out_lo = d->out.d_lo;
  e.out_dtype = d->out.dtype;
  e.res = d->res.d_ptr;
  e.res_dtype = d->res.dtype;
  e.acc = d->acc_in.d_ptr;
  e.acc_dtype = d->acc_in.dtype;
  e.bias = d->w->d_bias;
  e.div = d->div;
  e.coldiv = d->d_coldiv;
  e.use_div = (d->div != 1.0f && d->div != 0.0f) ? 1 : 0;
  BVG_REQUIRE(!(e.coldiv && e.use_div), "conv: d_coldiv and div != 1 exclude each other");
  BVG_REQUIRE(!e.coldiv || (((uintptr_t)e.coldiv) & 15) == 0, "conv: d_coldiv must be 16-byte aligned");
  e.relu = d->relu ? 1 : 0;
  e.N = d->w->n_total;
  BVG_REQUIRE(e.out != nullptr, "conv: null output");
  BVG_REQUIRE(e.out_dtype >= BVG_F32 && e.out_dtype <= BVG_SPLIT, "conv: bad output dtype");
  BVG_REQUIRE(e.out_dtype != BVG_SPLIT || e.out_lo, "conv: SPLIT output needs a lo plane");
  BVG_REQUIRE(!e.res || e.res_dtype == BVG_F32 || e.res_dtype == BVG_BF16, "conv: residual must be F32 or BF16");
  BVG_REQUIRE(!e.acc || e.acc_dtype == BVG_F32 || e.acc_dtype == BVG_BF16, "conv: acc_in must be F32 or BF16");
  BVG_REQUIRE(e.bias != nullptr, "conv: null bias");
  return BVG_OK;
}

}  // namespace bvg
