// Exact-fp32 implicit-GEMM convolution on CUDA cores (FFMA): the parity anchor of the conv path.
//
// Replaces nn.Conv1d / nn.ConvTranspose1d forward as used at reference modules/bigvgan.py:428-431,
// :602, :607 in the "tap GEMM" formulation of include/bvg_b200.h.  Channels-last activations.
// 64x64 output tile per 256-thread block, 4x4 outputs per thread, K streamed in (tap, 16-channel)
// slices through shared memory.  Not the fast path (that is conv_umma.cu); it exists so that
//  (i) every layer shape has a true-fp32 result on the GPU to compare tcgen05 against, and
//  (ii) configurations the tensor-core kernel does not cover still run on the device.
#include "common.cuh"
#include "epilogue.cuh"

namespace bvg {

constexpr int SM_BM = 64, SM_BN = 64, SM_BK = 16;

struct SimtParams {
  const void* x;
  const void* x_lo;
  int x_dtype;
  int x_pitch;  // channels per row of x
  const float* w;  // [tile][tap][cin_pad][64]
  int cin_pad;
  int B, L, N;
  int tiles_per_item;
  int ntaps;
  int tap_stride;
  int shift[BVG_MAX_TAPS];
  EpiParams epi;
  int vec_ok;  // N % 4 == 0
};

__device__ __forceinline__ void simt_load_x4(const SimtParams& p, long long off, float (&v)[4]) {
  if (p.x_dtype == BVG_F32) {
    float4 t = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.x) + off));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
    uint2 t = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(p.x) + off));
    unpack_bf16x2(t.x, v[0], v[1]);
    unpack_bf16x2(t.y, v[2], v[3]);
    if (p.x_dtype == BVG_SPLIT) {
      uint2 u = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(p.x_lo) + off));
      float l[4];
      unpack_bf16x2(u.x, l[0], l[1]);
      unpack_bf16x2(u.y, l[2], l[3]);
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] += l[i];
    }
  }
}

__global__ void __launch_bounds__(256) conv_simt_kernel(const __grid_constant__ SimtParams p) {
  __shared__ float As[SM_BK][SM_BM + 4];
  __shared__ __align__(16) float Bs[SM_BK][SM_BN];

  const int tid = threadIdx.x;
  const int b = blockIdx.x / p.tiles_per_item;
  const int t0 = (blockIdx.x % p.tiles_per_item) * SM_BM;
  const int ntile = blockIdx.y;
  const int tx = tid % 16, ty = tid / 16;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int a_row = tid / 4;         // 0..63
  const int a_c4 = (tid % 4) * 4;    // channel offset inside the K slice
  const int b_k = tid / 16;          // 0..15
  const int b_n4 = (tid % 16) * 4;

  for (int tap = 0; tap < p.ntaps; ++tap) {
    const int t_in = t0 + a_row + p.shift[tap];
    const bool row_ok = (t_in >= 0) && (t_in < p.L);
    const long long x_row = ((long long)b * p.L + (row_ok ? t_in : 0)) * p.x_pitch;
    const float* wt = p.w + ((long long)(ntile * p.tap_stride + tap) * p.cin_pad) * SM_BN;
    for (int c0 = 0; c0 < p.cin_pad; c0 += SM_BK) {
      float av[4] = {0.f, 0.f, 0.f, 0.f};
      if (row_ok && c0 + a_c4 < p.cin_pad) simt_load_x4(p, x_row + c0 + a_c4, av);
      float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c0 + b_k < p.cin_pad) bv = __ldg(reinterpret_cast<const float4*>(wt + (long long)(c0 + b_k) * SM_BN + b_n4));
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 4; ++i) As[a_c4 + i][a_row] = av[i];
      *reinterpret_cast<float4*>(&Bs[b_k][b_n4]) = bv;
      __syncthreads();
#pragma unroll
      for (int k = 0; k < SM_BK; ++k) {
        float a[4], bb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
        const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        bb[0] = b4.x; bb[1] = b4.y; bb[2] = b4.z; bb[3] = b4.w;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
      }
    }
  }

  const int n0 = ntile * SM_BN + tx * 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int t = t0 + ty * 4 + i;
    if (t >= p.L) continue;
    const long long row = (long long)b * p.L + t;
    if (p.vec_ok) {
      if (n0 < p.N) epilogue4(p.epi, row, n0, acc[i]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n0 + j < p.N) epilogue1(p.epi, row, n0 + j, acc[i][j]);
    }
  }
}

int conv_simt_forward(const bvg_conv_desc* d, cudaStream_t st) {
  const bvg_conv_weights* w = d->w;
  BVG_REQUIRE(w->backend == BVG_SIMT, "conv_simt: weights were packed for another backend");
  BVG_REQUIRE(w->n_tile == SM_BN, "conv_simt: weights must be packed with n_tile = 64");
  BVG_REQUIRE(d->x.d_ptr != nullptr, "conv_simt: null input");
  BVG_REQUIRE(d->x.dtype != BVG_SPLIT || d->x.d_lo, "conv_simt: SPLIT input needs a lo plane");
  BVG_REQUIRE(w->cin_pad % 4 == 0, "conv_simt: packed cin must be a multiple of 4");
  SimtParams p;
  p.x = d->x.d_ptr;
  p.x_lo = d->x.d_lo;
  p.x_dtype = d->x.dtype;
  p.x_pitch = w->x_pitch;
  p.w = reinterpret_cast<const float*>(w->d_w);
  p.cin_pad = w->cin_pad;
  p.B = d->B;
  p.L = d->L;
  p.N = w->n_total;
  p.tiles_per_item = ceil_div(d->L, SM_BM);
  p.ntaps = w->n_taps[0];
  p.tap_stride = w->tap_stride;
  BVG_REQUIRE(p.ntaps > 0 && p.ntaps <= BVG_MAX_TAPS, "conv_simt: bad tap count %d", p.ntaps);
  for (int t = 0; t < p.ntaps; ++t) p.shift[t] = w->shift[0][t];
  int rc = fill_epilogue(d, p.epi);
  if (rc != BVG_OK) return rc;
  p.vec_ok = (w->n_total % 4 == 0) ? 1 : 0;
  const long long gx = (long long)d->B * p.tiles_per_item;
  BVG_REQUIRE(gx < (1ll << 31) && w->n_tiles <= 65535, "conv_simt: grid too large");
  dim3 grid((unsigned)gx, (unsigned)w->n_tiles);
  conv_simt_kernel<<<grid, 256, 0, st>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "conv_simt_kernel launch");
  return BVG_OK;
}

}  // namespace bvg
