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
  pdl_trigger();  // programmatic dependent launch (common.cuh); the weights were packed at load time
  for (int i = threadIdx.x; i < C * K; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  pdl_wait();
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
    BVG_CHECK_CUDA(launch_k(post_kernel<true>, dim3((unsigned)blocks), dim3(256), smem, st, d->x.d_ptr, d->d_w, d->bias, d->d_out, d->L, d->C, d->ksize, total));
  else
    BVG_CHECK_CUDA(launch_k(post_kernel<false>, dim3((unsigned)blocks), dim3(256), smem, st, d->x.d_ptr, d->d_w, d->bias, d->d_out, d->L, d->C, d->ksize, total));
  return BVG_OK;
}

// ------------------------------------------------------------------------------------------------
// Waveform tail: fade-out, peak normalisation, silence padding, PCM16 (see bvg_tail_desc).
// torch.linspace(1, 0, N) in fp32 (ATen RangeFactories: step = (end - start) / (N - 1); the first
// half counts up from start, the second half counts down from end):
//   i < N / 2 : 1 + step * i        else : 0 - step * (N - 1 - i)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float fade_weight(int i, int n) {
  if (n == 1) return 1.0f;
  const float step = __fdiv_rn(-1.0f, (float)(n - 1));
  return i < n / 2 ? __fadd_rn(1.0f, __fmul_rn(step, (float)i)) : __fsub_rn(0.0f, __fmul_rn(step, (float)(n - 1 - i)));
}

__device__ __forceinline__ float faded(const float* __restrict__ w, long long b, int i, int L, int fade_len) {
  float v = __ldg(w + b * L + i);
  const int k = i - (L - fade_len);
  if (k >= 0) v = __fmul_rn(v, fade_weight(k, fade_len));
  return v;
}

__global__ void __launch_bounds__(256) tail_peak_kernel(const float* __restrict__ w, float* __restrict__ peak, int L, int fade_len, int chunks) {
  const int b = blockIdx.x / chunks, c = blockIdx.x % chunks;
  const long long per = ((long long)L + chunks - 1) / chunks;
  const long long i0 = c * per, i1 = min((long long)L, i0 + per);
  float m = 0.f;
  for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) m = fmaxf(m, fabsf(faded(w, b, (int)i, L, fade_len)));
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; ++k) m = fmaxf(m, red[k]);
    atomicMax(reinterpret_cast<unsigned int*>(peak) + b, __float_as_uint(m));  // m >= 0: uint order == float order
  }
}

__global__ void __launch_bounds__(256) tail_pcm_kernel(const float* __restrict__ w, const float* __restrict__ peak, int16_t* __restrict__ pcm, int L,
                                                       int fade_len, int silence, float volume_peak, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int Lo = L + 2 * silence;
  const long long b = idx / Lo;
  const int i = (int)(idx % Lo) - silence;
  float v = 0.f;
  if (i >= 0 && i < L) {
    v = faded(w, b, i, L, fade_len);
    if (volume_peak > 0.f) {
      const float pk = peak[b];
      v = pk > 0.f ? __fmul_rn(v, __fdiv_rn(volume_peak, pk)) : 0.f;  // waveform * (volume_peak / max |waveform|)
    }
  }
  const float q = rintf(__fmul_rn(v, 32768.0f));
  pcm[idx] = (int16_t)fminf(fmaxf(q, -32768.0f), 32767.0f);
}

int tail_forward(const bvg_tail_desc* d, cudaStream_t st) {
  BVG_REQUIRE(d && d->d_wave && d->d_pcm && d->d_peak, "tail: null pointer");
  BVG_REQUIRE(d->B > 0 && d->L > 0 && d->fade_len >= 0 && d->silence >= 0, "tail: bad shape");
  BVG_REQUIRE(d->fade_len <= d->L, "tail: fade of %d samples does not fit a %d-sample waveform (fewer than 20 frames)", d->fade_len, d->L);
  BVG_CHECK_CUDA(cudaMemsetAsync(d->d_peak, 0, sizeof(float) * d->B, st));
  int chunks = 1;
  while (chunks < 1024 && (long long)d->L / (chunks * 2) >= 4096) chunks *= 2;
  tail_peak_kernel<<<(unsigned)(d->B * chunks), 256, 0, st>>>(d->d_wave, d->d_peak, d->L, d->fade_len, chunks);
  BVG_CHECK_CUDA(cudaGetLastError());
  const long long total = (long long)d->B * (d->L + 2ll * d->silence);
  const long long blocks = ceil_div_ll(total, 256);
  BVG_REQUIRE(blocks < (1ll << 31), "tail: grid too large");
  tail_pcm_kernel<<<(unsigned)blocks, 256, 0, st>>>(d->d_wave, d->d_peak, d->d_pcm, d->L, d->fade_len, d->silence, d->volume_peak, total);
  BVG_CHECK_CUDA(cudaGetLastError());
  return BVG_OK;
}

// ------------------------------------------------------------------------------------------------
// Cross-fade stitch of time chunks (see bvg_stitch_desc): one thread per four samples of a chunk's kept range.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stitch_kernel(const float* __restrict__ wave, long long wave_stride, float* __restrict__ out,
                                                     const long long* __restrict__ table, int first, int step, int blocks_per_chunk) {
  const int j = first + step * (int)(blockIdx.x / blocks_per_chunk);
  const long long* row = table + 5ll * j;
  const long long dst = row[0], src = row[1], n = row[2], fin = row[3], fout = row[4];
  const float* w = wave + (long long)j * wave_stride + src;
  float* o = out + dst;
  const long long per = (long long)blockDim.x * 4;
  for (long long i0 = ((long long)(blockIdx.x % blocks_per_chunk) * blockDim.x + threadIdx.x) * 4; i0 < n; i0 += per * blocks_per_chunk) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long i = i0 + k;
      if (i >= n) break;
      float v = w[i];
      if (i < fin) {
        v = __fmul_rn(v, __fdiv_rn(__fadd_rn((float)i, 0.5f), (float)fin));
        o[i] = __fadd_rn(o[i], v);
      } else if (i >= n - fout) {
        const long long r = i - (n - fout);
        v = __fmul_rn(v, __fsub_rn(1.0f, __fdiv_rn(__fadd_rn((float)r, 0.5f), (float)fout)));
        o[i] = __fadd_rn(o[i], v);
      } else {
        o[i] = v;
      }
    }
  }
}

int stitch_forward(const bvg_stitch_desc* d, cudaStream_t st) {
  BVG_REQUIRE(d && d->d_wave && d->d_out && d->d_table && d->h_table, "stitch: null pointer");
  BVG_REQUIRE(d->n_chunks > 0 && d->wave_stride > 0 && d->out_len > 0, "stitch: bad shape");
  long long max_n = 0;
  for (int j = 0; j < d->n_chunks; ++j) {
    const int64_t* r = d->h_table + 5ll * j;
    BVG_REQUIRE(r[0] >= 0 && r[1] >= 0 && r[2] > 0 && r[3] >= 0 && r[4] >= 0, "stitch: row %d has a negative entry", j);
    BVG_REQUIRE(r[0] + r[2] <= d->out_len && r[1] + r[2] <= d->wave_stride, "stitch: row %d leaves its buffers", j);
    BVG_REQUIRE(r[3] + r[4] <= r[2], "stitch: row %d fades %lld + %lld samples of %lld", j, (long long)r[3], (long long)r[4], (long long)r[2]);
    if (j + 1 < d->n_chunks) {  // rows two apart must not overlap (they run in the same launch)
      const int64_t* q = d->h_table + 5ll * (j + 1);
      // whatever rows j and j+1 share of d_out lies inside j's fade-out window and inside j+1's fade-in window
      BVG_REQUIRE(q[0] >= r[0] + r[2] - r[4] && r[0] + r[2] <= q[0] + q[3],
                  "stitch: rows %d and %d are not consecutive chunks overlapping only in their fade windows", j, j + 1);
      if (j + 2 < d->n_chunks) BVG_REQUIRE((d->h_table + 5ll * (j + 2))[0] >= r[0] + r[2], "stitch: rows %d and %d overlap", j, j + 2);
    }
    if (r[2] > max_n) max_n = r[2];
  }
  int bpc = (int)ceil_div_ll(max_n, 256 * 4 * 4);
  if (bpc < 1) bpc = 1;
  if (bpc > 4096) bpc = 4096;
  for (int first = 0; first < 2 && first < d->n_chunks; ++first) {
    const int count = (d->n_chunks - first + 1) / 2;
    stitch_kernel<<<(unsigned)(count * bpc), 256, 0, st>>>(d->d_wave, d->wave_stride, d->d_out, reinterpret_cast<const long long*>(d->d_table), first, 2, bpc);
    BVG_CHECK_CUDA(cudaGetLastError());
  }
  return BVG_OK;
}

}  // namespace bvg
