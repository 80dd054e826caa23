// Tail of the generator: conv_post (Conv1d C -> 1, k = 7, zero pad 3) + tanh, reading the
// channels-last output of activation_post and writing the waveform [B, L] (== [B, 1, L]).
// Replaces reference modules/bigvgan.py:619-620.  One output sample per thread: a C x k dot
// product (168 FMAs for the repo config) -- bandwidth-bound, not a tensor-core shape.
#include "common.cuh"

namespace bvg {

constexpr int POST_MAX_W = 4096;  // floats of folded weight kept in shared memory

template <bool IN_BF16>
__global__ void __launch_bounds__(256) post_kernel(const void* __restrict__ x, const float* __restrict__ w, float bias,
                                                   float* __restrict__ out, int L, int C, int K, long long total) {
  extern __shared__ float ws[];
  for (int i = threadIdx.x; i < C * K; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int t = (int)(idx % L);
  const long long b = idx / L;
  const int pad = (K - 1) / 2;
  float acc = bias;
  for (int k = 0; k < K; ++k) {
    const int ti = t + k - pad;
    if (ti < 0 || ti >= L) continue;
    const long long row = (b * L + ti) * C;
    const float* wk = ws + k * C;
    if ((C & 3) == 0) {
      for (int c = 0; c < C; c += 4) {
        float v[4];
        if constexpr (IN_BF16) {
          uint2 u = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(x) + row + c));
          unpack_bf16x2(u.x, v[0], v[1]);
          unpack_bf16x2(u.y, v[2], v[3]);
        } else {
          float4 f = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x) + row + c));
          v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) acc = fmaf(wk[c + i], v[i], acc);
      }
    } else {
      for (int c = 0; c < C; ++c) {
        float v;
        if constexpr (IN_BF16)
          v = bf16_bits_to_float(__ldg(reinterpret_cast<const uint16_t*>(x) + row + c));
        else
          v = __ldg(reinterpret_cast<const float*>(x) + row + c);
        acc = fmaf(wk[c], v, acc);
      }
    }
  }
  out[idx] = tanhf(acc);
}

int post_forward(const bvg_post_desc* d, cudaStream_t st) {
  BVG_REQUIRE(d && d->x.d_ptr && d->d_w && d->d_out, "post: null pointer");
  BVG_REQUIRE(d->B > 0 && d->L > 0 && d->C > 0 && d->ksize > 0 && (d->ksize & 1), "post: bad shape");
  BVG_REQUIRE(d->x.dtype == BVG_F32 || d->x.dtype == BVG_BF16, "post: input must be F32 or BF16");
  BVG_REQUIRE(d->C * d->ksize <= POST_MAX_W, "post: weight too large for shared memory");
  const long long total = (long long)d->B * d->L;
  const long long blocks = ceil_div_ll(total, 256);
  BVG_REQUIRE(blocks < (1ll << 31), "post: grid too large");
  const size_t smem = (size_t)d->C * d->ksize * sizeof(float);
  if (d->x.dtype == BVG_BF16)
    post_kernel<true><<<(unsigned)blocks, 256, smem, st>>>(d->x.d_ptr, d->d_w, d->bias, d->d_out, d->L, d->C, d->ksize, total);
  else
    post_kernel<false><<<(unsigned)blocks, 256, smem, st>>>(d->x.d_ptr, d->d_w, d->bias, d->d_out, d->L, d->C, d->ksize, total);
  BVG_CHECK_CUDA(cudaGetLastError());
  return BVG_OK;
}

}  // namespace bvg
