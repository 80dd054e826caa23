// Load-time kernels: weight-norm fold + weight packing, mel head transpose, format conversion.
//
// Weight norm (torch.nn.utils.weight_norm, dim=0; reference modules/bigvgan.py:319-386, 529, 550,
// 593) is w = g * v / ||v|| with the norm over all dims but dim 0 -- Cout for Conv1d, *Cin* for
// ConvTranspose1d.  The reference recomputes it in a forward pre-hook on every call (the hooks
// are never removed: utils/load_models.py:52-79); here it is folded once into the packed planes.
#include "common.cuh"

namespace bvg {

// scale[i] = g[i] / ||v[i, :]||_2   (1 when g == nullptr, i.e. v is already the folded weight)
__global__ void wn_scale_kernel(const float* __restrict__ v, const float* __restrict__ g, int inner, float* __restrict__ scale) {
  const int i = blockIdx.x;
  if (g == nullptr) {
    if (threadIdx.x == 0) scale[i] = 1.0f;
    return;
  }
  const float* row = v + (long long)i * inner;
  float s = 0.f;
  for (int k = threadIdx.x; k < inner; k += blockDim.x) s = fmaf(row[k], row[k], s);
  __shared__ float red[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) scale[i] = g[i] / sqrtf(s);
  }
}

struct PackParams {
  const float* v;
  const float* scale;
  int transposed, cin, cout, ksize, dilation, stride, padding, fold;
  int backend, n_total, n_tile, n_tiles, cin_pad, tap_stride, stacked;
  int n_taps[BVG_MAX_NTILES];
  int shift[BVG_MAX_NTILES][BVG_MAX_TAPS];
  float* out_f32;
  uint16_t* out_hi;
  uint16_t* out_lo;
  long long total;
};

// value of the folded weight seen by output column n through the tap with input-row shift `sh`
__device__ __forceinline__ float folded_weight(const PackParams& p, int n, int ci, int sh) {
  if (n >= p.n_total || ci >= p.cin * p.fold) return 0.f;
  if (p.fold > 1) {
    // time-folded Conv1d: output row P q + po takes x[P (q + sh) + pi] through original tap j with
    // j d - padding = P sh + pi - po  (block-Toeplitz arrangement of the original taps)
    const int po = n / p.cout, co = n % p.cout;
    const int pi = ci / p.cin, c = ci % p.cin;
    const int num = p.fold * sh + pi - po + p.padding;
    if (num < 0 || num % p.dilation != 0) return 0.f;
    const int j = num / p.dilation;
    if (j >= p.ksize) return 0.f;
    return p.scale[co] * p.v[((long long)co * p.cin + c) * p.ksize + j];
  }
  if (!p.transposed) {
    // out[t] = sum_j w[n, ci, j] x[t - padding + j d]  =>  shift = j d - padding
    const int num = sh + p.padding;
    if (num < 0 || num % p.dilation != 0) return 0.f;
    const int j = num / p.dilation;
    if (j >= p.ksize) return 0.f;
    return p.scale[n] * p.v[((long long)n * p.cin + ci) * p.ksize + j];
  }
  // ConvTranspose1d: output position u q + r receives x[q + delta] through kernel index
  // kk = r + padding - delta u   (from  u q + r = (q + delta) u - padding + kk)
  const int r = n / p.cout, co = n % p.cout;
  const int kk = r + p.padding - sh * p.stride;
  if (kk < 0 || kk >= p.ksize) return 0.f;
  return p.scale[ci] * p.v[((long long)ci * p.cout + co) * p.ksize + kk];
}

__global__ void pack_weights_kernel(const __grid_constant__ PackParams p) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.total) return;
  int tile, slot, nl, ci, plane = 0;
  if (p.backend == BVG_SIMT) {  // [tile][slot][ci][nl]
    nl = (int)(idx % p.n_tile);
    long long r = idx / p.n_tile;
    ci = (int)(r % p.cin_pad);
    r /= p.cin_pad;
    slot = (int)(r % p.tap_stride);
    tile = (int)(r / p.tap_stride);
  } else {  // [tile][slot][nl][ci]   (stacked: [tile][slot][plane][nl][ci])
    ci = (int)(idx % p.cin_pad);
    long long r = idx / p.cin_pad;
    nl = (int)(r % p.n_tile);
    r /= p.n_tile;
    if (p.stacked) {
      plane = (int)(r & 1);
      r >>= 1;
    }
    slot = (int)(r % p.tap_stride);
    tile = (int)(r / p.tap_stride);
  }
  float w = 0.f;
  const int tt = (p.backend == BVG_SIMT) ? 0 : tile;  // SIMT: one tap table for all tiles
  if (slot < p.n_taps[tt]) w = folded_weight(p, tile * p.n_tile + nl, ci, p.shift[tt][slot]);
  if (p.backend == BVG_SIMT) {
    p.out_f32[idx] = w;
  } else {
    float hi, lo;
    split_bf16(w, hi, lo);
    if (p.stacked) {
      p.out_hi[idx] = (uint16_t)float_to_bf16_bits(plane ? lo : hi);
    } else {
      p.out_hi[idx] = (uint16_t)float_to_bf16_bits(hi);
      if (p.out_lo) p.out_lo[idx] = (uint16_t)float_to_bf16_bits(lo);
    }
  }
}

__global__ void pack_bias_kernel(const float* __restrict__ bias, int cout, int n_total, int n_padded, float* __restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_padded) return;
  out[n] = (n < n_total && bias != nullptr) ? bias[n % cout] : 0.f;
}

// bvg_tuning knobs read here (bvg_conv_geom.tune): umma_ntile_cap = widest UMMA N tile chosen by conv_geometry (multiple
// of 16), umma_stack = widest n_tile whose (hi, lo) weight planes are stacked along N (0 = never)

static int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
static int round_up(int a, int b) { return (a + b - 1) / b * b; }

// Tap tables and tiling of one layer (host).  Fills everything in `w` except the pointers.
int conv_geometry(const bvg_conv_geom* g, bvg_conv_weights* w) {
  BVG_REQUIRE(g && w, "conv geometry: null argument");
  BVG_REQUIRE(g->cin > 0 && g->cout > 0 && g->ksize > 0, "conv geometry: bad sizes");
  BVG_REQUIRE(g->backend == BVG_SIMT || g->backend == BVG_UMMA, "conv geometry: bad backend");
  const bool tr = g->transposed != 0;
  BVG_REQUIRE(tr ? g->stride > 0 : g->dilation > 0, "conv geometry: bad stride/dilation");
  const int fold = g->fold > 1 ? g->fold : 1;
  BVG_REQUIRE(fold == 1 || (!tr && g->backend == BVG_UMMA), "conv geometry: time folding is for Conv1d on the UMMA backend");
  const int umma_ntile_cap = tune_of(g->tune).umma_ntile_cap, umma_stack = tune_of(g->tune).umma_stack, umma_pair = tune_of(g->tune).umma_pair;
  w->backend = g->backend;
  w->cin = g->cin * fold;
  w->split = (g->backend == BVG_UMMA && g->split) ? 1 : 0;
  const int n_total = tr ? g->stride * g->cout : fold * g->cout;
  w->n_total = n_total;

  int n_tile;
  if (g->backend == BVG_SIMT) {
    n_tile = 64;
    w->cin_pad = round_up(g->cin, 4);
    w->x_pitch = round_up(g->cin, 4);
  } else {
    w->cin_pad = round_up(g->cin * fold, 64);
    w->x_pitch = round_up(g->cin * fold, 8);
    if (g->n_tile > 0) {
      n_tile = g->n_tile;
    } else {
      // fewest tiles of at most 256 columns; a transposed conv keeps each tile inside one phase
      // when the per-phase width allows (then every tile has exactly the taps of its phase)
      // split operands, more than one tile's worth of columns (C >= 384: K = Cin x taps in the thousands, where the
      // tensor core's truncating fp32 accumulation is the largest error of the fp32 path): at most 128 columns, so
      // that the layer runs "stacked" (below) with the correction products in their own accumulator columns.
      // Measured on B200 (gpurun_out/r02_ops_fp32_b.txt): C = 768 same speed as 3 x 256, C = 384 9 % faster than 2 x 192;
      // C = 192 keeps its single 192-column tile (2 x 96 stacked is 8 % slower and K <= 2112 there)
      const int cap_max = (umma_ntile_cap >= 16 && umma_ntile_cap <= 256) ? umma_ntile_cap / 16 * 16 : 256;
      // (with the CTA-pair kernel, conv_pair.cu, the wide layers keep 256 / 192-column tiles and two separate weight
      // planes: there the correction products have their own accumulator columns without stacking)
      int cap = (w->split && !umma_pair && umma_stack >= 128 && cap_max > 128 && n_total > 256) ? 128 : cap_max;
      // SPLIT layers with several N tiles (C >= 384) take 128-column tiles on the pair kernel too: main + correction
      // columns are then 256 per stage and TWO accumulator stages fit, so the epilogue overlaps the next tile (ncu on
      // the one-stage 256 / 192-column tiles: 93 % / 87 % tensor-pipe active; measured per forward C = 384 20.6 -> 18.4 ms,
      // C = 768 19.6 -> 19.1 ms).  umma_pair == 4 keeps the wide tiles (A/B).
      if (w->split && umma_pair && umma_pair != 4 && n_total >= 384 && cap > 128) cap = 128;
      for (;;) {
        const int base = (tr && g->cout % 16 == 0) ? g->cout : n_total;
        const int t = ceil_div(base, cap);
        n_tile = round_up(ceil_div(base, t), 16);
        if (tr && g->cout % 16 == 0 && g->cout % n_tile != 0) n_tile = round_up(ceil_div(n_total, ceil_div(n_total, cap)), 16);
        if (ceil_div(n_total, n_tile) <= BVG_MAX_NTILES || cap >= 256) break;
        // a very wide layer (a transposed conv with 8 x 768 output columns; 2C = 2048 at a 32-column cap): the cap is
        // a preference, the per-tile tap tables (BVG_MAX_NTILES) are a limit -- widen until the tile count fits
        cap = cap < cap_max ? cap_max : (cap * 2 > 256 ? 256 : cap * 2);
      }
    }
    BVG_REQUIRE(n_tile % 16 == 0 && n_tile >= 16 && n_tile <= 256, "conv geometry: UMMA n_tile %d must be a multiple of 16 in [16, 256]", n_tile);
  }
  w->n_tile = n_tile;
  w->n_tiles = ceil_div(n_total, n_tile);
  // Narrow split layers: both weight planes in ONE operand, rows [0, n_tile) = hi, [n_tile, 2 n_tile) = lo
  // per tap ("stacked", split = 2).  A_hi x [W_hi; W_lo] is then a single MMA of width 2 n_tile, so a product
  // costs two reads of the A tile from shared memory instead of three (these layers are bound by exactly that).
  // (a 128-column tile that the CTA-pair kernel takes keeps two separate planes: each CTA of the pair loads half of a box)
  const int pair_min = tune_of(g->tune).umma_pair_min > 0 ? tune_of(g->tune).umma_pair_min : 128;
  const bool for_pair = umma_pair && n_tile >= pair_min && n_tile % 32 == 0 && n_total % 4 == 0;
  if (w->split && n_tile <= umma_stack && n_tile <= 128 && !for_pair) w->split = 2;
  const int n_tables = (g->backend == BVG_SIMT) ? 1 : w->n_tiles;
  BVG_REQUIRE(n_tables <= BVG_MAX_NTILES, "conv geometry: %d N tiles exceed BVG_MAX_NTILES", w->n_tiles);
  for (int t = 0; t < BVG_MAX_NTILES; ++t) {
    w->n_taps[t] = 0;
    for (int k = 0; k < BVG_MAX_TAPS; ++k) w->shift[t][k] = 0;
  }

  int max_taps = 0;
  for (int t = 0; t < n_tables; ++t) {
    int nt = 0;
    if (fold > 1) {
      // super-row shifts sh with some (p_in, p_out, j):  P sh = j d - padding - (p_in - p_out),  |p_in - p_out| < P
      const int s_lo = -floor_div(g->padding + fold - 1, fold);
      const int s_hi = floor_div((g->ksize - 1) * g->dilation - g->padding + fold - 1, fold);
      BVG_REQUIRE(s_hi >= s_lo && s_hi - s_lo + 1 <= BVG_MAX_TAPS, "conv geometry: folded conv needs %d taps", s_hi - s_lo + 1);
      for (int sh = s_lo; sh <= s_hi; ++sh) w->shift[t][nt++] = sh;
    } else if (!tr) {
      BVG_REQUIRE(g->ksize <= BVG_MAX_TAPS, "conv geometry: kernel size %d exceeds BVG_MAX_TAPS", g->ksize);
      for (int j = 0; j < g->ksize; ++j) w->shift[t][nt++] = j * g->dilation - g->padding;
    } else {
      int r_lo = 0, r_hi = g->stride - 1;
      if (g->backend == BVG_UMMA) {  // phases this tile touches
        r_lo = (t * n_tile) / g->cout;
        r_hi = (((t + 1) * n_tile < n_total ? (t + 1) * n_tile : n_total) - 1) / g->cout;
      }
      // delta valid for phase r  <=>  0 <= r + padding - delta u < k
      const int d_hi = floor_div(r_hi + g->padding, g->stride);
      const int d_lo = -floor_div(g->ksize - 1 - g->padding - r_lo, g->stride);
      BVG_REQUIRE(d_hi - d_lo + 1 <= BVG_MAX_TAPS && d_hi >= d_lo, "conv geometry: transposed conv needs %d taps", d_hi - d_lo + 1);
      for (int dlt = d_lo; dlt <= d_hi; ++dlt) w->shift[t][nt++] = dlt;
    }
    w->n_taps[t] = nt;
    if (nt > max_taps) max_taps = nt;
  }
  w->tap_stride = max_taps;
  return BVG_OK;
}

size_t conv_plane_elems(const bvg_conv_weights* w) {
  return (size_t)w->n_tiles * w->tap_stride * w->n_tile * w->cin_pad * (w->split == 2 ? 2 : 1);
}

int pack_conv_weights(const bvg_conv_geom* g, const float* d_v, const float* d_g, const float* d_bias, bvg_conv_weights* w,
                      float* d_bias_out, float* d_scale_scratch, cudaStream_t st) {
  int rc = conv_geometry(g, w);
  if (rc != BVG_OK) return rc;
  BVG_REQUIRE(d_v && w->d_w && d_bias_out && d_scale_scratch, "pack weights: null pointer");
  BVG_REQUIRE(w->split != 1 || w->d_w_lo, "pack weights: split packing needs d_w_lo");
  const int dim0 = g->transposed ? g->cin : g->cout;
  const int inner = (g->transposed ? g->cout : g->cin) * g->ksize;
  wn_scale_kernel<<<dim0, 256, 0, st>>>(d_v, d_g, inner, d_scale_scratch);
  BVG_CHECK_CUDA(cudaGetLastError());

  PackParams p;
  p.v = d_v;
  p.scale = d_scale_scratch;
  p.transposed = g->transposed;
  p.cin = g->cin;
  p.cout = g->cout;
  p.ksize = g->ksize;
  p.dilation = g->dilation > 0 ? g->dilation : 1;
  p.stride = g->stride > 0 ? g->stride : 1;
  p.padding = g->padding;
  p.fold = g->fold > 1 ? g->fold : 1;
  p.backend = w->backend;
  p.n_total = w->n_total;
  p.n_tile = w->n_tile;
  p.n_tiles = w->n_tiles;
  p.cin_pad = w->cin_pad;
  p.tap_stride = w->tap_stride;
  p.stacked = w->split == 2;
  for (int t = 0; t < BVG_MAX_NTILES; ++t) {
    p.n_taps[t] = w->n_taps[t];
    for (int k = 0; k < BVG_MAX_TAPS; ++k) p.shift[t][k] = w->shift[t][k];
  }
  p.out_f32 = (w->backend == BVG_SIMT) ? reinterpret_cast<float*>(w->d_w) : nullptr;
  p.out_hi = (w->backend == BVG_UMMA) ? reinterpret_cast<uint16_t*>(w->d_w) : nullptr;
  p.out_lo = (w->backend == BVG_UMMA && w->split == 1) ? reinterpret_cast<uint16_t*>(w->d_w_lo) : nullptr;
  p.total = (long long)conv_plane_elems(w);
  const long long blocks = ceil_div_ll(p.total, 256);
  BVG_REQUIRE(blocks < (1ll << 31), "pack weights: too large");
  pack_weights_kernel<<<(unsigned)blocks, 256, 0, st>>>(p);
  BVG_CHECK_CUDA(cudaGetLastError());
  const int n_padded = w->n_tiles * w->n_tile;
  pack_bias_kernel<<<ceil_div(n_padded, 256), 256, 0, st>>>(d_bias, g->cout, w->n_total, n_padded, d_bias_out);
  BVG_CHECK_CUDA(cudaGetLastError());
  w->d_bias = d_bias_out;
  return BVG_OK;
}

// conv_post weights: w_out[j][c] = g v[0, c, j] / ||v||
__global__ void pack_post_kernel(const float* __restrict__ v, const float* __restrict__ scale, int cin, int ksize, float* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= cin * ksize) return;
  const int c = idx % cin, j = idx / cin;
  out[idx] = scale[0] * v[c * ksize + j];
}

int pack_post_weights(const float* d_v, const float* d_g, int cin, int ksize, float* d_w_out, float* d_scale_scratch, cudaStream_t st) {
  BVG_REQUIRE(d_v && d_w_out && d_scale_scratch && cin > 0 && ksize > 0, "pack post weights: bad argument");
  wn_scale_kernel<<<1, 256, 0, st>>>(d_v, d_g, cin * ksize, d_scale_scratch);
  BVG_CHECK_CUDA(cudaGetLastError());
  pack_post_kernel<<<ceil_div(cin * ksize, 128), 128, 0, st>>>(d_v, d_scale_scratch, cin, ksize, d_w_out);
  BVG_CHECK_CUDA(cudaGetLastError());
  return BVG_OK;
}

// ---- mel head: [B, C, T] float -> channels-last [B, T, c_pad] ------------------------------------
template <int OUT_MODE>
__global__ void pack_mel_kernel(const float* __restrict__ mel, void* out, void* out_lo, int C, int T, int c_pad, const float* __restrict__ range,
                                const float* __restrict__ mn) {
  __shared__ float tile[32][33];
  pdl_trigger();  // programmatic dependent launch (common.cuh): no global access before the wait
  pdl_wait();
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, t = t0 + tx;
    float v = 0.f;
    if (c < C && t < T) {
      v = mel[((long long)b * C + c) * T + t];
      // fused denormalize_mel_channel: ((mel + 1) / 2) * range + min, each operation rounded to fp32
      if (range) v = __fadd_rn(__fmul_rn(__fmul_rn(__fadd_rn(v, 1.0f), 0.5f), range[c]), mn[c]);
    }
    tile[i][tx] = v;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int t = t0 + i, c = c0 + tx;
    if (t < T && c < c_pad) {
      const float v = tile[tx][i];
      const long long off = ((long long)b * T + t) * c_pad + c;
      if (OUT_MODE == BVG_F32) {
        reinterpret_cast<float*>(out)[off] = v;
      } else {
        float hi, lo;
        split_bf16(v, hi, lo);
        reinterpret_cast<uint16_t*>(out)[off] = (uint16_t)float_to_bf16_bits(OUT_MODE == BVG_BF16 ? v : hi);
        if (OUT_MODE == BVG_SPLIT) reinterpret_cast<uint16_t*>(out_lo)[off] = (uint16_t)float_to_bf16_bits(lo);
      }
    }
  }
}

int pack_mel(const bvg_pack_desc* d, cudaStream_t st) {
  BVG_REQUIRE(d && d->d_mel && d->out.d_ptr, "pack_mel: null pointer");
  BVG_REQUIRE(d->B > 0 && d->C > 0 && d->T > 0 && d->c_pad >= d->C, "pack_mel: bad shape");
  BVG_REQUIRE(d->out.dtype != BVG_SPLIT || d->out.d_lo, "pack_mel: SPLIT output needs a lo plane");
  BVG_REQUIRE(d->B <= 65535, "pack_mel: batch too large");
  BVG_REQUIRE((d->d_range == nullptr) == (d->d_min == nullptr), "pack_mel: d_range and d_min go together");
  dim3 grid(ceil_div(d->T, 32), ceil_div(d->c_pad, 32), d->B), block(32, 8);
  if (d->out.dtype == BVG_F32)
    BVG_CHECK_CUDA(launch_k(pack_mel_kernel<BVG_F32>, grid, block, 0, st, d->d_mel, d->out.d_ptr, (void*)nullptr, d->C, d->T, d->c_pad, d->d_range, d->d_min));
  else if (d->out.dtype == BVG_BF16)
    BVG_CHECK_CUDA(launch_k(pack_mel_kernel<BVG_BF16>, grid, block, 0, st, d->d_mel, d->out.d_ptr, (void*)nullptr, d->C, d->T, d->c_pad, d->d_range, d->d_min));
  else if (d->out.dtype == BVG_SPLIT)
    BVG_CHECK_CUDA(launch_k(pack_mel_kernel<BVG_SPLIT>, grid, block, 0, st, d->d_mel, d->out.d_ptr, d->out.d_lo, d->C, d->T, d->c_pad, d->d_range, d->d_min));
  else
    BVG_REQUIRE(false, "pack_mel: bad dtype");
  return BVG_OK;
}

// ---- element format conversion -----------------------------------------------------------------
__global__ void convert_kernel(const void* src, const void* src_lo, int sdt, void* dst, void* dst_lo, int ddt, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v;
  if (sdt == BVG_F32) {
    v = reinterpret_cast<const float*>(src)[i];
  } else {
    v = bf16_bits_to_float(reinterpret_cast<const uint16_t*>(src)[i]);
    if (sdt == BVG_SPLIT) v += bf16_bits_to_float(reinterpret_cast<const uint16_t*>(src_lo)[i]);
  }
  if (ddt == BVG_F32) {
    reinterpret_cast<float*>(dst)[i] = v;
  } else if (ddt == BVG_BF16) {
    reinterpret_cast<uint16_t*>(dst)[i] = (uint16_t)float_to_bf16_bits(v);
  } else {
    float hi, lo;
    split_bf16(v, hi, lo);
    reinterpret_cast<uint16_t*>(dst)[i] = (uint16_t)float_to_bf16_bits(hi);
    reinterpret_cast<uint16_t*>(dst_lo)[i] = (uint16_t)float_to_bf16_bits(lo);
  }
}

int convert(const bvg_tensor* src, const bvg_tensor* dst, size_t n, cudaStream_t st) {
  BVG_REQUIRE(src && dst && src->d_ptr && dst->d_ptr, "convert: null pointer");
  BVG_REQUIRE(src->dtype != BVG_SPLIT || src->d_lo, "convert: SPLIT source needs a lo plane");
  BVG_REQUIRE(dst->dtype != BVG_SPLIT || dst->d_lo, "convert: SPLIT destination needs a lo plane");
  if (n == 0) return BVG_OK;
  const long long blocks = ceil_div_ll((long long)n, 256);
  BVG_REQUIRE(blocks < (1ll << 31), "convert: too large");
  convert_kernel<<<(unsigned)blocks, 256, 0, st>>>(src->d_ptr, src->d_lo, src->dtype, dst->d_ptr, dst->d_lo, dst->dtype, (long long)n);
  BVG_CHECK_CUDA(cudaGetLastError());
  return BVG_OK;
}

}  // namespace bvg
