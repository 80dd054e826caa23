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
  pdl_trigger();
  pdl_wait();
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
  BVG_CHECK_CUDA(launch_k(rowop_kernel, dim3((unsigned)blocks), dim3(256), 0, st, p));
  return BVG_OK;
}

// ------------------------------------------------------------------------------------------------
// Step encoder + the n_layers diffusion projections (see bvg_diffembed_desc).  Grid (batch item, slice of the
// n_layers x C projection outputs): every CTA recomputes the two small SiLU layers (2 x fc x fc MACs, weights from
// L2) and then its slice of the projections.  A dot product is one warp's float4-strided sum + shuffle reduction,
// four rows at a time so that four independent loads are in flight per lane: the first version (one CTA per item,
// one row at a time) spent 0.87 ms here -- half of a B = 1 step -- waiting on one L2 round trip per row.
// ------------------------------------------------------------------------------------------------
constexpr int DE_WARPS = 8;
constexpr int DE_ROWS_PER_WARP = 8;  // projection outputs per warp -> 64 per CTA

__device__ __forceinline__ float warp_sum(float s) {
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return s;
}
// out[r] = w[r * n .. r * n + n) . x for R consecutive rows (n % 4 == 0, rows 16-byte aligned)
template <int R>
__device__ __forceinline__ void warp_dot_rows(const float* __restrict__ w, const float* x, int n, int lane, float (&out)[R]) {
  float s[R];
#pragma unroll
  for (int r = 0; r < R; ++r) s[r] = 0.f;
  for (int k = lane * 4; k < n; k += 128) {
    const float4 xv = *reinterpret_cast<const float4*>(x + k);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float4 wv = __ldg(reinterpret_cast<const float4*>(w + (long long)r * n + k));
      s[r] = fmaf(wv.x, xv.x, s[r]);
      s[r] = fmaf(wv.y, xv.y, s[r]);
      s[r] = fmaf(wv.z, xv.z, s[r]);
      s[r] = fmaf(wv.w, xv.w, s[r]);
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) out[r] = warp_sum(s[r]);
}
__device__ __forceinline__ float silu(float v) { return __fdiv_rn(v, 1.0f + expf(-v)); }

// y[o] = act(W[o, :] . x + b[o]) for o in [o0, o1), rows dealt to the warps four at a time
template <bool SILU>
__device__ __forceinline__ void dense_rows(const float* __restrict__ W, const float* __restrict__ bias, const float* x, int n, int o0, int o1, float* y,
                                           long long y_off, int warp, int nw, int lane) {
  for (int o = o0 + 4 * warp; o < o1; o += 4 * nw) {
    if (o + 4 <= o1) {
      float s[4];
      warp_dot_rows<4>(W + (long long)o * n, x, n, lane, s);
      if (lane < 4) {
        const float v = (lane == 0 ? s[0] : lane == 1 ? s[1] : lane == 2 ? s[2] : s[3]) + __ldg(bias + o + lane);
        y[y_off + o + lane] = SILU ? silu(v) : v;
      }
    } else {
      for (int r = o; r < o1; ++r) {
        float s[1];
        warp_dot_rows<1>(W + (long long)r * n, x, n, lane, s);
        if (lane == 0) {
          const float v = s[0] + __ldg(bias + r);
          y[y_off + r] = SILU ? silu(v) : v;
        }
      }
    }
  }
}

__global__ void __launch_bounds__(32 * DE_WARPS) diffembed_kernel(const __grid_constant__ bvg_diffembed_desc p) {
  extern __shared__ __align__(16) float sm[];
  float* e = sm;             // [emb]
  float* h1 = e + p.emb;     // [fc]
  float* h2 = h1 + p.fc;     // [fc]
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  pdl_trigger();
  pdl_wait();
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
  dense_rows<true>(p.d_w1, p.d_b1, e, p.emb, 0, p.fc, h1, 0, warp, nw, lane);
  __syncthreads();
  dense_rows<true>(p.d_w2, p.d_b2, h1, p.fc, 0, p.fc, h2, 0, warp, nw, lane);
  __syncthreads();
  // this CTA's slice of the n_layers * C projection outputs; output o = layer * C + c goes to d_out[layer][b][c]
  const int total = p.n_layers * p.C;
  const int per_cta = DE_WARPS * DE_ROWS_PER_WARP;
  const int o0 = blockIdx.y * per_cta, o1 = min(o0 + per_cta, total);
  for (int o = o0 + 4 * warp; o < o1; o += 4 * nw) {
    const int cnt = min(4, o1 - o);
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    if (cnt == 4) {
      warp_dot_rows<4>(p.d_wd + (long long)o * p.fc, h2, p.fc, lane, s);
    } else {
      for (int r = 0; r < cnt; ++r) {
        float s1[1];
        warp_dot_rows<1>(p.d_wd + (long long)(o + r) * p.fc, h2, p.fc, lane, s1);
        s[r] = s1[0];
      }
    }
    if (lane < cnt) {
      const int oo = o + lane;
      const int layer = oo / p.C, c = oo - layer * p.C;
      const float v = lane == 0 ? s[0] : lane == 1 ? s[1] : lane == 2 ? s[2] : s[3];
      p.d_out[((long long)layer * p.B + b) * p.C + c] = v + __ldg(p.d_bd + oo);
    }
  }
}

int diffembed_forward(const bvg_diffembed_desc* d, cudaStream_t st) {
  BVG_REQUIRE(d && (d->d_step || d->d_step_f) && d->d_table && d->d_w1 && d->d_b1 && d->d_w2 && d->d_b2 && d->d_wd && d->d_bd && d->d_out, "diffembed: null pointer");
  BVG_REQUIRE(d->B > 0 && d->emb > 0 && d->fc > 0 && d->C > 0 && d->n_layers > 0 && d->max_steps > 0, "diffembed: bad shape");
  BVG_REQUIRE(d->emb % 4 == 0 && d->fc % 4 == 0, "diffembed: embedding and hidden sizes must be multiples of 4");
  BVG_REQUIRE((((uintptr_t)d->d_w1 | (uintptr_t)d->d_w2 | (uintptr_t)d->d_wd) & 15) == 0, "diffembed: weight matrices must be 16-byte aligned");
  const size_t smem = sizeof(float) * ((size_t)d->emb + 2 * (size_t)d->fc);
  BVG_REQUIRE(smem <= 48 * 1024, "diffembed: embedding / hidden sizes too large");
  const int slices = ceil_div(d->n_layers * d->C, DE_WARPS * DE_ROWS_PER_WARP);
  BVG_REQUIRE(slices <= 65535, "diffembed: too many projection outputs");
  BVG_CHECK_CUDA(launch_k(diffembed_kernel, dim3((unsigned)d->B, (unsigned)slices), dim3(32 * DE_WARPS), smem, st, *d));
  return BVG_OK;
}

// ------------------------------------------------------------------------------------------------
// Sampler update between two denoiser calls (see bvg_sample_desc; reference modules/diffsvcrepo_inference.py
// p_sample :88-97 and p_sample_plms :100-150).  One thread per element of x[B, L, n_mel]; the intrinsics keep every
// operation separately rounded, in the order the reference's tensor expressions evaluate them.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sample_kernel(const __grid_constant__ bvg_sample_desc p) {
  pdl_trigger();
  pdl_wait();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per_item = (long long)p.L * p.n_mel;
  if (idx >= per_item * p.B) return;
  const int b = (int)(idx / per_item);
  const int rem = (int)(idx - b * per_item);
  const int t = min(max(__ldg(p.d_step + b), 0), p.n_steps - 1);
  const float x = p.d_x[idx];
  const float eps = p.d_eps[idx];
  if (p.d_eps_save) p.d_eps_save[idx] = eps;
  float r;
  if (p.mode == BVG_SAMPLE_DDPM) {
    // predict_start_from_noise (:32-36), clamp (:76-77)
    float x0 = __fsub_rn(__fmul_rn(__ldg(p.d_sqrt_recip + t), x), __fmul_rn(__ldg(p.d_sqrt_recipm1 + t), eps));
    if (p.clip) x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
    // q_posterior mean (:40-43)
    r = __fadd_rn(__fmul_rn(__ldg(p.d_coef1 + t), x0), __fmul_rn(__ldg(p.d_coef2 + t), x));
    if (t != 0) {  // nonzero_mask * (0.5 * model_log_variance).exp() * noise (:95-97)
      const int l = rem / p.n_mel, m = rem - l * p.n_mel;
      const float nz = __ldg(p.d_noise + ((long long)b * p.n_mel + m) * p.L + l);
      r = __fadd_rn(r, __fmul_rn(expf(__fmul_rn(0.5f, __ldg(p.d_logvar + t))), nz));
    }
  } else {
    float e = eps;
    if (p.combine == 1) {
      e = __fdiv_rn(__fadd_rn(p.d_hist[0][idx], eps), 2.0f);
    } else if (p.combine == 2) {
      e = __fdiv_rn(__fsub_rn(__fmul_rn(3.0f, eps), p.d_hist[0][idx]), 2.0f);
    } else if (p.combine == 3) {
      e = __fdiv_rn(__fadd_rn(__fsub_rn(__fmul_rn(23.0f, eps), __fmul_rn(16.0f, p.d_hist[0][idx])), __fmul_rn(5.0f, p.d_hist[1][idx])), 12.0f);
    } else if (p.combine == 4) {
      e = __fdiv_rn(__fsub_rn(__fadd_rn(__fsub_rn(__fmul_rn(55.0f, eps), __fmul_rn(59.0f, p.d_hist[0][idx])), __fmul_rn(37.0f, p.d_hist[1][idx])),
                              __fmul_rn(9.0f, p.d_hist[2][idx])),
                    24.0f);
    }
    // get_x_pred (:105-121)
    const float a_t = __ldg(p.d_alphas_cumprod + t), a_prev = __ldg(p.d_alphas_cumprod + max(t - p.interval, 0));
    const float sq_t = __fsqrt_rn(a_t), sq_prev = __fsqrt_rn(a_prev);
    const float cx = __fdiv_rn(1.0f, __fmul_rn(sq_t, __fadd_rn(sq_t, sq_prev)));
    const float ce = __fdiv_rn(1.0f, __fmul_rn(sq_t, __fadd_rn(__fsqrt_rn(__fmul_rn(__fsub_rn(1.0f, a_prev), a_t)), __fsqrt_rn(__fmul_rn(__fsub_rn(1.0f, a_t), a_prev)))));
    const float delta = __fmul_rn(__fsub_rn(a_prev, a_t), __fsub_rn(__fmul_rn(cx, x), __fmul_rn(ce, e)));
    r = __fadd_rn(x, delta);
  }
  p.d_x_out[idx] = r;
}

int sample_forward(const bvg_sample_desc* d, cudaStream_t st) {
  BVG_REQUIRE(d && d->d_x && d->d_x_out && d->d_eps && d->d_step, "sample: null pointer");
  BVG_REQUIRE(d->B > 0 && d->L > 0 && d->n_mel > 0 && d->n_steps > 0, "sample: bad shape");
  if (d->mode == BVG_SAMPLE_DDPM) {
    BVG_REQUIRE(d->d_noise && d->d_sqrt_recip && d->d_sqrt_recipm1 && d->d_coef1 && d->d_coef2 && d->d_logvar, "sample: DDPM needs the noise and the five schedule tables");
  } else {
    BVG_REQUIRE(d->mode == BVG_SAMPLE_PLMS, "sample: unknown mode %d", d->mode);
    BVG_REQUIRE(d->d_alphas_cumprod && d->interval > 0, "sample: PLMS needs alphas_cumprod and a positive interval");
    BVG_REQUIRE(d->combine >= 0 && d->combine <= 4, "sample: unknown combination %d", d->combine);
    const int need = d->combine == 0 ? 0 : d->combine == 1 ? 1 : d->combine - 1;
    for (int k = 0; k < need; ++k) BVG_REQUIRE(d->d_hist[k], "sample: combination %d needs %d earlier predictions", d->combine, need);
  }
  const long long total = (long long)d->B * d->L * d->n_mel;
  const long long blocks = ceil_div_ll(total, 256);
  BVG_REQUIRE(blocks < (1ll << 31), "sample: grid too large");
  BVG_CHECK_CUDA(launch_k(sample_kernel, dim3((unsigned)blocks), dim3(256), 0, st, *d));
  return BVG_OK;
}

}  // namespace bvg
