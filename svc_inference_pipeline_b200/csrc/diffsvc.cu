// Row-wise kernels of the DiffSVC denoiser step (reference modules/diffsvc.py:192-232,
// :284-321).  The step's dense layers run on the tensor-core tap-GEMM kernel (conv_umma.cu); these are the
// elementwise stages between them, on channels-last [B, L, C] fp32 tensors, writing the operand format the next
// GEMM reads (SPLIT planes on the fp32 path), and the tiny step-embedding MLP.
#include "common.cuh"

namespace bvg {

struct RowopParams {
  const float* x;
  const float* vec;
  void* out;
  void* out_lo;
  float div;
  int kind, out_dtype;
  int L, C, x_pitch, out_pitch;
  long long total;  // B * L * out_pitch / 4
};

// one thread = four consecutive output channels of one row
__global__ void __launch_bounds__(256) rowop_kernel(const __grid_constant__ RowopParams p) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.total) return;
  const int per_row = p.out_pitch >> 2;
  const long long row = idx / per_row;
  const int c0 = (int)(idx % per_row) * 4;
  const int b = (int)(row / p.L);
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (c0 < p.C) {  // C % 4 == 0: a group is either all real channels or all padding
    const float* xr = p.x + row * p.x_pitch + c0;
    const float4 a = *reinterpret_cast<const float4*>(xr);
    if (p.kind == BVG_ROW_ADDVEC) {
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
      if (p.vec) {  // y = x + diffusion_step (modules/diffsvc.py:213)
        const float4 d = *reinterpret_cast<const float4*>(p.vec + (long long)b * p.C + c0);
        v[0] += d.x; v[1] += d.y; v[2] += d.z; v[3] += d.w;
      }
    } else if (p.kind == BVG_ROW_GATE) {  // sigmoid(gate) * tanh(filter), gate = channels [0, C), filter = [C, 2C) (:225-227)
      const float4 f = *reinterpret_cast<const float4*>(xr + p.C);
      const float g[4] = {a.x, a.y, a.z, a.w}, t[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = __fdiv_rn(1.0f, 1.0f + expf(-g[i])) * tanhf(t[i]);
    } else {  // skip / sqrt(n_layers) (:313): a true division, like the reference
      v[0] = __fdiv_rn(a.x, p.div); v[1] = __fdiv_rn(a.y, p.div); v[2] = __fdiv_rn(a.z, p.div); v[3] = __fdiv_rn(a.w, p.div);
    }
  }
  const long long off = row * p.out_pitch + c0;
  if (p.out_dtype == BVG_F32) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + off) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
    float hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_bf16(v[i], hi[i], lo[i]);
    uint2 h, l;
    h.x = pack_bf16x2(hi[0], hi[1]); h.y = pack_bf16x2(hi[2], hi[3]);
    *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p.out) + off) = h;
    if (p.out_dtype == BVG_SPLIT) {
      l.x = pack_bf16x2(lo[0], lo[1]); l.y = pack_bf16x2(lo[2], lo[3]);
      *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p.out_lo) + off) = l;
    }
  }
}

int rowop_forward(const bvg_rowop_desc* d, cudaStream_t st) {
  BVG_REQUIRE(d && d->d_x && d->out.d_ptr, "rowop: null pointer");
  BVG_REQUIRE(d->kind >= BVG_ROW_ADDVEC && d->kind <= BVG_ROW_SCALE, "rowop: unknown kind %d", d->kind);
  BVG_REQUIRE(d->B > 0 && d->L > 0 && d->C > 0 && d->C % 4 == 0 && d->out_pitch % 4 == 0 && d->x_pitch % 4 == 0, "rowop: channel counts must be multiples of 4");
  BVG_REQUIRE(d->out_pitch >= d->C && d->x_pitch >= (d->kind == BVG_ROW_GATE ? 2 * d->C : d->C), "rowop: row pitch smaller than the channels read / written");
  BVG_REQUIRE(d->out.dtype >= BVG_F32 && d->out.dtype <= BVG_SPLIT && (d->out.dtype != BVG_SPLIT || d->out.d_lo), "rowop: bad output tensor");
  BVG_REQUIRE(d->kind != BVG_ROW_SCALE || d->div != 0.f, "rowop: division by zero");
  BVG_REQUIRE((((uintptr_t)d->d_x | (uintptr_t)d->out.d_ptr | (uintptr_t)d->out.d_lo | (uintptr_t)d->d_vec) & 15) == 0, "rowop: pointers must be 16-byte aligned");
  RowopParams p;
  p.x = d->d_x;
  p.vec = d->kind == BVG_ROW_ADDVEC ? d->d_vec : nullptr;
  p.out = d->out.d_ptr;
  p.out_lo = d->out.d_lo;
  p.div = d->div;
  p.kind = d->kind;
  p.out_dtype = d->out.dtype;
  p.L = d->L;
  p.C = d->C;
  p.x_pitch = d->x_pitch;
  p.out_pitch = d->out_pitch;
  p.total = (long long)d->B * d->L * (d->out_pitch / 4);
  const long long blocks = ceil_div_ll(p.total, 256);
  BVG_REQUIRE(blocks < (1ll << 31), "rowop: grid too large");
  rowop_kernel<<<(unsigned)blocks, 256, 0, st>>>(p);
  BVG_CHECK_CUDA(cudaGetLastError());
  return BVG_OK;
}

// ------------------------------------------------------------------------------------------------
// Step encoder + the n_layers diffusion projections (see bvg_diffembed_desc).  One CTA per batch item; every dot
// product is one warp's strided sum + shuffle reduction (the matrices are 128 x 128 and n_layers x C x 128: microseconds).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_dot(const float* __restrict__ w, const float* __restrict__ x, int n, int lane) {
  float s = 0.f;
  for (int k = lane; k < n; k += 32) s = fmaf(__ldg(w + k), x[k], s);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return s;
}
__device__ __forceinline__ float silu(float v) { return __fdiv_rn(v, 1.0f + expf(-v)); }

__global__ void __launch_bounds__(256) diffembed_kernel(const __grid_constant__ bvg_diffembed_desc p) {
  extern __shared__ float sm[];
  float* e = sm;             // [emb]
  float* h1 = e + p.emb;     // [fc]
  float* h2 = h1 + p.fc;     // [fc]
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (p.d_step_f) {  // lerp_embedding (:57-67): low + (high - low) * (t - low_idx)
    const float t = p.d_step_f[b];
    const int lo = min(max((int)floorf(t), 0), p.max_steps - 1), hi = min(max((int)ceilf(t), 0), p.max_steps - 1);
    const float w = t - (float)lo;
    for (int k = threadIdx.x; k < p.emb; k += blockDim.x) {
      const float a = __ldg(p.d_table + (long long)lo * p.emb + k), c = __ldg(p.d_table + (long long)hi * p.emb + k);
      e[k] = a + (c - a) * w;
    }
  } else {
    int step = p.d_step[b];
    step = min(max(step, 0), p.max_steps - 1);
    for (int k = threadIdx.x; k < p.emb; k += blockDim.x) e[k] = __ldg(p.d_table + (long long)step * p.emb + k);
  }
  __syncthreads();
  for (int o = warp; o < p.fc; o += nw) {
    const float s = warp_dot(p.d_w1 + (long long)o * p.emb, e, p.emb, lane);
    if (lane == 0) h1[o] = silu(s + __ldg(p.d_b1 + o));
  }
  __syncthreads();
  for (int o = warp; o < p.fc; o += nw) {
    const float s = warp_dot(p.d_w2 + (long long)o * p.fc, h1, p.fc, lane);
    if (lane == 0) h2[o] = silu(s + __ldg(p.d_b2 + o));
  }
  __syncthreads();
  const int total = p.n_layers * p.C;
  for (int o = warp; o < total; o += nw) {
    const int layer = o / p.C, c = o % p.C;
    const float s = warp_dot(p.d_wd + (long long)o * p.fc, h2, p.fc, lane);
    if (lane == 0) p.d_out[((long long)layer * p.B + b) * p.C + c] = s + __ldg(p.d_bd + o);
  }
}

int diffembed_forward(const bvg_diffembed_desc* d, cudaStream_t st) {
  BVG_REQUIRE(d && (d->d_step || d->d_step_f) && d->d_table && d->d_w1 && d->d_b1 && d->d_w2 && d->d_b2 && d->d_wd && d->d_bd && d->d_out, "diffembed: null pointer");
  BVG_REQUIRE(d->B > 0 && d->emb > 0 && d->fc > 0 && d->C > 0 && d->n_layers > 0 && d->max_steps > 0, "diffembed: bad shape");
  const size_t smem = sizeof(float) * ((size_t)d->emb + 2 * (size_t)d->fc);
  BVG_REQUIRE(smem <= 48 * 1024, "diffembed: embedding / hidden sizes too large");
  diffembed_kernel<<<(unsigned)d->B, 256, smem, st>>>(*d);
  BVG_CHECK_CUDA(cudaGetLastError());
  return BVG_OK;
}

}  // namespace bvg
