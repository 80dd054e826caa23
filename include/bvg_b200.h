/*
 * bvg_b200.h -- C ABI of libbvg_b200.so: the B200 (sm_100a) BigVGAN vocoder kernels.
 *
 * The reference (WallaceRao/svc_inference_pipeline) has no FFI / plugin interface on this path:
 * its boundary is three Python call sites (Generator(cfg.vocoder) modules/bigvgan.py:521,
 * vocoder_model_loader utils/load_models.py:52-79, synthesis_audios modules/bigvgan_inference.py:29)
 * and every arithmetic step is a PyTorch library call.  This header is therefore the interface the
 * Python drop-in (svc_inference_pipeline_b200/modules/bigvgan.py) binds with ctypes; each entry point
 * below cites the reference call it replaces.  See INTEGRATION.md for the binding stub.
 *
 * Conventions
 *  - plain C types only; every pointer named d_* is a DEVICE pointer borrowed from the caller
 *    (torch owns all memory; the library never allocates device memory);
 *  - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream);
 *  - no process-global mutable state: test / tuning knobs are a struct the caller attaches to a descriptor
 *    (bvg_tuning); the only caches are keyed by (device, kernel) and guarded by a mutex;
 *  - every function returns 0 on success, a negative BVG_E* code otherwise, and never throws;
 *    bvg_last_error() returns a thread-local human-readable message for the last failure;
 *  - activations are CHANNELS-LAST inside the library: [B, L, C] with C contiguous
 *    (the reference's external layout [B, C, L] is converted by bvg_pack_mel / produced by
 *    bvg_post_fwd);
 *  - element formats (bvg_dtype): F32, BF16, or SPLIT = two bf16 planes (hi, lo) with
 *    hi = bf16(x), lo = bf16(x - hi); SPLIT operands feed the 3-MMA error-compensated
 *    tensor-core product (hi*hi + hi*lo + lo*hi, fp32 accumulate) of the fp32-parity path.
 */
#ifndef BVG_B200_H_
#define BVG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BVG_ABI_VERSION 5  /* 2: bvg_conv_geom.fold, bvg_conv_desc.pre_amp; 3: bvg_stitch_fwd; 4: bvg_tuning in the descriptors (bvg_set_tuning removed), bvg_logmel_fwd, bvg_rowop_fwd, bvg_diffembed_fwd, bvg_conv_desc.relu; 5: bvg_sample_fwd, bvg_program_set_pdl, bvg_conv_desc.d_coldiv */

enum bvg_status {
  BVG_OK = 0,
  BVG_EINVAL = -1,   /* bad argument / unsupported shape */
  BVG_ECUDA = -2,    /* CUDA runtime or driver error (message has the CUDA string) */
  BVG_EARCH = -3,    /* device is not sm_100 */
  BVG_ENOMEM = -4    /* caller-provided buffer too small */
};

enum bvg_dtype { BVG_F32 = 0, BVG_BF16 = 1, BVG_SPLIT = 2 };
enum bvg_act { BVG_SNAKE = 0, BVG_SNAKEBETA = 1 };
enum bvg_backend {
  BVG_SIMT = 0, /* fp32 FFMA implicit GEMM on CUDA cores: exact-fp32 anchor path */
  BVG_UMMA = 1  /* tcgen05.mma + TMEM accumulators fed by TMA (bf16 operands, fp32 accumulate) */
};

/* Test / A-B knobs.  They travel WITH the descriptor they apply to (bvg_amp_desc.tune, bvg_conv_desc.tune,
 * bvg_conv_geom.tune; NULL = the library's own choices), so the library keeps no process-global mutable state
 * and two threads can run differently tuned calls side by side.  Fill with bvg_tuning_defaults() first. */
typedef struct bvg_tuning {
  int32_t amp_vec;         /* 0 = choose; else force channels per thread of the FFMA Activation1d kernel (1, 2, 4) */
  int32_t amp_chunk;       /* 0 = choose; else 12-step block pairs per thread */
  int32_t amp_mma;         /* default 1: tensor-core Activation1d for BF16 -> BF16; 2: also F32 -> SPLIT; 0: never */
  int32_t amp_mma_tiles;   /* 0 = choose; time tiles per CTA of the tensor-core Activation1d */
  int32_t amp_packed;      /* default 1: FFMA2 kernel for two channels per thread; 0: scalar FFMA */
  int32_t amp_stream;      /* default 0: per-warp streaming tensor-core Activation1d (F32 -> SPLIT) */
  int32_t amp_stream_bf16; /* default 0: the same for BF16 -> BF16 */
  int32_t amp_ct;          /* default 1: per-channel-count instantiations of the FFMA2 kernel */
  int32_t umma_mb;         /* 0 = choose; force M blocks per tile */
  int32_t umma_wide_mb2;   /* 0 = SPLIT only; 1 = every tile that fits; -1 = never: two row blocks on one TMEM stage */
  int32_t umma_max_ctas;   /* 0 = all SMs; cap of the persistent grid */
  int32_t umma_ntile_cap;  /* default 256: widest N tile bvg_conv_geometry chooses */
  int32_t umma_tap_group;  /* 0 = choose; taps per weight stage */
  int32_t umma_a_stages;   /* 0 = choose; activation stages (2..4) */
  int32_t umma_stack;      /* default 128: widest n_tile with stacked (hi, lo) weight planes (0 = never) */
  int32_t umma_pair;       /* default 1: wide convolutions (N tile >= 128) on the CTA-pair kernel (tcgen05.mma.cta_group::2);
                              0: single-CTA kernel, wide SPLIT layers then pack as stacked 128-column tiles;
                              2: single-tile SPLIT layers (C = 192) on the single-CTA kernel; 4: wide SPLIT layers keep 256 / 192-column tiles */
  int32_t umma_pair_smem_kb; /* 0 = default; cap (KB) on the pair kernel's dynamic shared memory: fewer weight stages leave
                              room for Activation1d CTAs of another stream on the same SM */
  int32_t umma_pair_min;     /* 0 = default (128): narrowest N tile that runs on the CTA-pair kernel (a multiple of 32) */
  int32_t _reserved[2];
} bvg_tuning;

void bvg_tuning_defaults(bvg_tuning* t);

/* A tensor handed to a kernel: base pointer(s) + element format.  `lo` is only read for SPLIT. */
typedef struct bvg_tensor {
  void* d_ptr;  /* F32: float*, BF16: bf16*, SPLIT: bf16* hi plane */
  void* d_lo;   /* SPLIT: bf16* lo plane, else NULL */
  int32_t dtype; /* bvg_dtype */
  int32_t _pad;
} bvg_tensor;

/* ------------------------------------------------------------------------------------------
 * Fused anti-aliased activation  (replaces Activation1d.forward, modules/bigvgan.py:251-256:
 * UpSample1d :278-287 -> Snake/SnakeBeta :84-95/:146-159 -> DownSample1d :304-307 / :224-231).
 * One kernel; the 2x-rate signal lives only in registers.
 *   y[b,t,c] = sum_k f[k] * s[clamp(2t+k-5)],  s = snake(u),  u = 2x polyphase upsample of x.
 * d_a[c]   = exp(alpha[c]) (or alpha[c] for linear scale);  d_invb[c] = 1/(b[c] + 1e-9)
 * with b = beta (snakebeta) or alpha (snake), exponentiated likewise: precomputed at load.
 * x: F32 or BF16 [B,L,C];  y: F32, BF16 or SPLIT [B,L,C].  taps: the 12 filter taps (host).
 * ------------------------------------------------------------------------------------------ */
typedef struct bvg_amp_desc {
  bvg_tensor x;
  bvg_tensor y;
  const float* d_a;
  const float* d_invb;
  float taps_up[12];   /* upsample.filter   (state_dict buffer, modules/bigvgan.py:273-276) */
  float taps_down[12]; /* downsample.lowpass.filter (modules/bigvgan.py:220-221) */
  int32_t B, L, C;
  int32_t fast_sin;    /* 1: MUFU.COS / MUFU.SIN on the raw argument 2 a u (default of both paths: measured equal to the
                          reduced form on every golden, tests/test_gpu_ops.py::test_amp_large_argument); 0: the argument is
                          first reduced exactly to [-1/2, 1/2] turns (Generator(precise_sin=True)) */
  const bvg_tuning* tune; /* NULL = defaults */
} bvg_amp_desc;

int bvg_amp_fwd(const bvg_amp_desc* d, void* stream);

/* ------------------------------------------------------------------------------------------
 * Dense convolution as a "tap GEMM"  (replaces the weight-normed Conv1d of AMPBlock1/2,
 * modules/bigvgan.py:319-386 / :428-431, conv_pre :529-537 / :602, and the ConvTranspose1d
 * upsamplers :547-561 / :607):
 *   out[b, t, n] = epi( bias[n] + sum_{tap} sum_{ci} x[b, t + shift[tile(n)][tap], ci] * W[n][tap][ci] )
 * rows of x outside [0, L) read as zero (Conv1d zero padding; ConvTranspose1d has no such taps).
 * Conv1d(k, dilation d): N = Cout, shifts (j - (k-1)/2) * d.  ConvTranspose1d(k, stride u):
 * N = u*Cout (phase-major), out viewed as [B, L, u*Cout] == [B, L*u, Cout]; each phase uses the
 * input shifts its taps touch.  epi: + res, + acc_in, / div, store in out.dtype.
 * Weights are pre-folded (w = g*v/||v||, torch.nn.utils.weight_norm) and pre-packed by
 * bvg_pack_conv_weights into the layout of the chosen backend.
 * ------------------------------------------------------------------------------------------ */
#define BVG_MAX_TAPS 16
#define BVG_MAX_NTILES 32  /* UMMA tiles with their own tap table; SIMT uses entry 0 for all tiles */

typedef struct bvg_conv_weights {
  int32_t backend;     /* bvg_backend */
  int32_t cin;         /* input channels (GEMM K per tap) */
  int32_t n_total;     /* GEMM N: Cout (conv) or u*Cout (transposed conv) */
  int32_t n_tile;      /* N-tile the packing was made for (UMMA: multiple of 16, <= 256) */
  int32_t n_tiles;
  int32_t cin_pad;     /* K extent of the packed weights: SIMT round_up(cin,4), UMMA round_up(cin,64) */
  int32_t x_pitch;     /* channel pitch the x tensor must have: SIMT round_up(cin,4), UMMA round_up(cin,8) */
  int32_t tap_stride;  /* tap slots reserved per N tile in the packed planes (>= max n_taps) */
  int32_t split;       /* UMMA: 1 = hi and lo planes packed (fp32-parity path); 2 = both planes stacked along N in d_w
                          (rows [0, n_tile) hi, [n_tile, 2 n_tile) lo per tap; chosen by bvg_conv_geometry for n_tile <= 128) */
  int32_t n_taps[BVG_MAX_NTILES];                  /* taps of each N tile */
  int32_t shift[BVG_MAX_NTILES][BVG_MAX_TAPS];    /* input row shift of each tap */
  void* d_w;           /* SIMT: float [tile][tap][cin_pad][n_tile]; UMMA: bf16 [tile][tap][n_tile][cin_pad] */
  void* d_w_lo;        /* UMMA split: lo plane, same layout */
  const float* d_bias; /* [n_total] */
} bvg_conv_weights;

typedef struct bvg_conv_desc {
  bvg_tensor x;        /* [B, L, cin]   SIMT: F32;  UMMA: BF16 or SPLIT */
  bvg_tensor out;      /* [B, L, n_total] */
  bvg_tensor res;      /* optional residual, same shape as out (d_ptr NULL = none); F32 or BF16 */
  bvg_tensor acc_in;   /* optional running sum to add (xs += ...), same shape; F32 or BF16 */
  float div;           /* out /= div when != 1 (the xs / num_kernels of modules/bigvgan.py:615) */
  int32_t B, L;
  const bvg_conv_weights* w;
  const struct bvg_amp_desc* pre_amp; /* optional (UMMA, 8 <= Cin <= 64): fuse this Activation1d in front of the convolution --
                                         the kernel computes its operand from pre_amp->x (F32 [B, L, Cin]) and x above is
                                         ignored; pre_amp->y is not written.  The AMP-into-conv fusion of the narrow stages:
                                         modules/bigvgan.py:428-431  xt = c1(a1(x)); xt = c2(a2(xt)) */
  const bvg_tuning* tune; /* NULL = defaults */
  int32_t relu;        /* 1: out = max(out, 0) as the last epilogue step (the F.relu after the 1x1 projections of the
                          DiffSVC denoiser, modules/diffsvc.py:126, :316) */
  int32_t _pad;
  const float* d_coldiv; /* optional [n_total]: out[.., n] /= d_coldiv[n] (a true division per output channel) instead of
                            the scalar div, which must then be 1.  One launch for the DiffSVC residual layer's
                            output_projection (modules/diffsvc.py:229-232, :307): channels [0, C) are "(x + residual) /
                            sqrt(2)", channels [C, 2C) the skip sum (divisor 1), with res = out = the [x | skip] buffer. */
} bvg_conv_desc;

int bvg_conv_fwd(const bvg_conv_desc* d, void* stream);

/* Weight-norm fold + pack (replaces the per-forward torch._weight_norm pre-hook, i.e.
 * torch.nn.utils.weight_norm at modules/bigvgan.py:319-386,529,550,593: done once at load).
 *   d_v: weight_v  Conv1d [Cout, Cin, K]   / ConvTranspose1d [Cin, Cout, K]
 *   d_g: weight_g  [dim0,1,1] or NULL when d_v already holds the folded weight
 * Fills w->n_taps/shift/cin_pad/n_tiles and writes the packed planes into w->d_w (/d_w_lo)
 * and the phase-replicated bias into d_bias_out.  Query sizes first with bvg_conv_pack_bytes. */
typedef struct bvg_conv_geom {
  int32_t transposed;  /* 0: Conv1d, 1: ConvTranspose1d */
  int32_t cin, cout, ksize;
  int32_t dilation;    /* Conv1d */
  int32_t stride;      /* ConvTranspose1d upsampling rate u */
  int32_t padding;     /* Conv1d: get_padding(k, d); ConvTranspose1d: (k - u) / 2 */
  int32_t backend;     /* bvg_backend */
  int32_t split;       /* UMMA only */
  int32_t n_tile;      /* 0 = choose */
  int32_t fold;        /* Conv1d on UMMA only; 0 / 1 = none.  P > 1: "time folding" for narrow layers -- P consecutive
                          rows of [B, L, cin] are read as one row of [B, L/P, P*cin] (same memory, L % P == 0) and the
                          layer becomes a Conv1d with P*cin -> P*cout channels whose taps are the block-Toeplitz
                          arrangement of the original ones (row p_out of a block of outputs takes original tap j from
                          input row P*shift + p_in with j*d - padding = P*shift + p_in - p_out).  The caller passes
                          L/P as bvg_conv_desc.L; w->cin, n_total, x_pitch describe the folded layer. */
  int32_t _pad;
  const bvg_tuning* tune; /* NULL = defaults (umma_ntile_cap, umma_stack) */
} bvg_conv_geom;

int bvg_conv_pack_bytes(const bvg_conv_geom* g, size_t* weight_plane_bytes, size_t* bias_bytes);
int bvg_conv_geometry(const bvg_conv_geom* g, bvg_conv_weights* w); /* host only: tiling + tap tables */
/* d_scratch: >= max(cin, cout) floats of device scratch (per-slice g/||v||). */
int bvg_pack_conv_weights(const bvg_conv_geom* g, const float* d_v, const float* d_g, const float* d_bias,
                          bvg_conv_weights* w, float* d_bias_out, float* d_scratch, void* stream);

/* ------------------------------------------------------------------------------------------
 * Tail: conv_post (Conv1d C->1, k=7, pad 3) + tanh  (modules/bigvgan.py:593, :619-620).
 * x: F32 or BF16 [B, L, C];  d_w: float [7][C] folded;  d_out: float [B, L] (== [B,1,L]).
 * ------------------------------------------------------------------------------------------ */
typedef struct bvg_post_desc {
  bvg_tensor x;
  const float* d_w;
  float bias;
  float* d_out;
  int32_t B, L, C, ksize;
} bvg_post_desc;

int bvg_post_fwd(const bvg_post_desc* d, void* stream);
int bvg_pack_post_weights(const float* d_v, const float* d_g, int32_t cin, int32_t ksize, float* d_w_out,
                          float* d_scratch /* >= 1 float */, void* stream);

/* Head: mel [B, C, T] float (reference layout, modules/bigvgan.py:600) -> channels-last operand
 * [B, T, c_pad] (F32, BF16 or SPLIT), zero-filling channels C..c_pad-1.
 * Optional fused de-normalisation (replaces denormalize_mel_channel,
 * utils/acoustic_feature_extraction.py:83-97, the step infer.py:80 runs on the host between the
 * acoustic model and the vocoder): when d_range != NULL the kernel reads a mel normalised to
 * [-1, 1] and forms ((mel + 1) / 2) * d_range[c] + d_min[c] in fp32, one rounding per operation like
 * the reference's numpy expression, with d_range[c] = mel_max[c] - mel_min[c] + 1e-12. */
typedef struct bvg_pack_desc {
  const float* d_mel;
  bvg_tensor out;
  int32_t B, C, T, c_pad;
  const float* d_range; /* [C] or NULL */
  const float* d_min;   /* [C] or NULL */
} bvg_pack_desc;

int bvg_pack_mel(const bvg_pack_desc* d, void* stream);

/* ------------------------------------------------------------------------------------------
 * Waveform tail: what infer.py:86-90 does on the host after the
 * vocoder -- synthesis_audios' fade-out (modules/bigvgan_inference.py:37-42: the last fade_len =
 * 20 * hop_length samples times torch.linspace(1, 0, fade_len)) and save_audio's peak
 * normalisation to volume_peak, 50 ms of leading / trailing silence and 16-bit PCM quantisation
 * (utils/util.py:20-37) -- fused into two kernels, so the device -> host copy is int16.
 *   d_wave: float [B, L] generator output (not modified);  d_pcm: int16 [B, silence + L + silence];
 *   d_peak: float [B] scratch, receives max |faded wave| per item.
 * PCM convention: clamp(rint(v * 32768), -32768, 32767) (torchaudio's encoder is not importable in
 * this image, so the quantiser is stated, not pinned).  A silent item (peak 0) stays silent.
 * ------------------------------------------------------------------------------------------ */
typedef struct bvg_tail_desc {
  const float* d_wave;
  int16_t* d_pcm;
  float* d_peak;
  int32_t B, L;
  int32_t fade_len;   /* 0 = no fade; must be <= L (the reference fails for T < 20 frames) */
  int32_t silence;    /* samples of silence on each side (fs / 20 in save_audio; 0 = none) */
  float volume_peak;  /* 0.9 in save_audio; <= 0 = no peak normalisation ("turn_up=False") */
  int32_t _pad;
} bvg_tail_desc;

int bvg_tail_fwd(const bvg_tail_desc* d, void* stream);

/* ------------------------------------------------------------------------------------------
 * Cross-fade stitch of time chunks (the reference vocodes one whole utterance per
 * forward, modules/bigvgan_inference.py:34-36, and has no chunking -- this is the "cross-fades the chunks"
 * step of the long-form path).  d_wave holds the generator outputs of n_chunks equal-length input
 * windows (one batch, row stride wave_stride floats).  Row j of d_table = {dst, src, n, fade_in,
 * fade_out} (int64): samples src .. src+n-1 of chunk j go to d_out[dst ..], the first fade_in of them
 * ramped up with weight (i + 0.5) / fade_in, the last fade_out ramped down with 1 - (i + 0.5) / fade_out.
 * The body is stored, the two fade windows are ADDED to what d_out holds (the neighbour's half of the
 * blend, or zero: d_out must be zero-filled before the first chunk of a span).  Neighbouring chunks of
 * one call overlap in their fade windows, so the call runs the even and the odd rows as two launches.
 * ------------------------------------------------------------------------------------------ */
typedef struct bvg_stitch_desc {
  const float* d_wave;
  int64_t wave_stride;
  float* d_out;
  int64_t out_len;        /* bound check: dst + n <= out_len for every row */
  const int64_t* d_table; /* DEVICE [n_chunks][5] */
  const int64_t* h_table; /* HOST copy of the same table (validated here; the device is not read back) */
  int32_t n_chunks;
  int32_t _pad;
} bvg_stitch_desc;

int bvg_stitch_fwd(const bvg_stitch_desc* d, void* stream);

/* ------------------------------------------------------------------------------------------
 * DiffSVC denoiser step (modules/diffsvc.py:192-232, :284-321 -- the function the
 * sampler calls 1000 times per utterance, modules/diffsvcrepo_inference.py:234-235).  Its dense layers
 * (dilated k=3 convolutions C -> 2C and the 1x1 projections) are tap GEMMs on bvg_conv_fwd; what is
 * left are row-wise operations on channels-last [B, L, C] tensors and the step-embedding MLP:
 *   BVG_ROW_ADDVEC: out[b,l,c] = x[b,l,c] + vec[b,c]      "y = x + diffusion_step" (:213), vec NULL = copy;
 *                   channels C .. out_pitch-1 of out are zero-filled (operand padding)
 *   BVG_ROW_GATE  : out[b,l,c] = sigmoid(x[b,l,c]) * tanh(x[b,l,C+c])   (:225-227; x has 2C channels)
 *   BVG_ROW_SCALE : out[b,l,c] = x[b,l,c] / div           "skip / sqrt(n_layers)" (:313)
 * x is F32 with row pitch x_pitch; out is F32, BF16 or SPLIT with row pitch out_pitch.
 * ------------------------------------------------------------------------------------------ */
enum bvg_rowop_kind { BVG_ROW_ADDVEC = 0, BVG_ROW_GATE = 1, BVG_ROW_SCALE = 2 };

typedef struct bvg_rowop_desc {
  int32_t kind; /* bvg_rowop_kind */
  int32_t _pad;
  const float* d_x;
  const float* d_vec; /* ADDVEC: [B, C] or NULL */
  bvg_tensor out;
  float div;
  int32_t B, L, C, x_pitch, out_pitch;
} bvg_rowop_desc;

int bvg_rowop_fwd(const bvg_rowop_desc* d, void* stream);

/* Step encoder + every residual layer's diffusion projection in one launch (modules/diffsvc.py:69-93 StepEncoder.forward
 * with an integer step, :205 ResidualBlock.diffusion_projection):
 *   e = d_table[step[b]]  (the sin / cos lookup table, built on the host like the reference builds its buffer; a
 *       fractional step interpolates between the two neighbouring rows);
 *   h = silu(W2 silu(W1 e + b1) + b2);   d_out[i][b][:] = Wd[i] h + bd[i]   for the n_layers layers.
 * Weights are torch Linear layouts ([out, in] row-major); d_wd = [n_layers][C][fc], d_bd = [n_layers][C]. */
typedef struct bvg_diffembed_desc {
  const int32_t* d_step; /* [B] integer steps, or NULL when d_step_f is given */
  const float* d_table;  /* [max_steps][emb] */
  const float* d_w1;     /* [fc][emb] */
  const float* d_b1;
  const float* d_w2;     /* [fc][fc] */
  const float* d_b2;
  const float* d_wd;
  const float* d_bd;
  float* d_out;          /* [n_layers][B][C] */
  int32_t B, emb, fc, C, n_layers, max_steps;
  const float* d_step_f; /* [B] fractional steps (StepEncoder.lerp_embedding, modules/diffsvc.py:57-67): e = low + (high - low) * (t - floor t) */
} bvg_diffembed_desc;

int bvg_diffembed_fwd(const bvg_diffembed_desc* d, void* stream);

/* One update of the diffusion sampler that drives the denoiser (modules/diffsvcrepo_inference.py): everything the
 * reference does between two denoiser calls, as one elementwise launch over the sample x[B, L, n_mel] (the memory
 * of the reference's x[B, 1, n_mel, T] transposed, i.e. the layout the denoiser reads).  Every product, sum and
 * quotient is rounded separately in fp32 in the reference's order (no contraction).
 *   BVG_SAMPLE_DDPM (p_sample :88-97 = predict_start_from_noise :32-36, clamp :76-77, q_posterior :39-50):
 *       x0 = clamp(sqrt_recip[t] x - sqrt_recipm1[t] eps, -1, 1)               (clamp iff clip)
 *       x' = (coef1[t] x0 + coef2[t] x) + [t != 0] exp(0.5 logvar[t]) noise
 *     noise is the reference's torch.randn(x.shape): [B, n_mel, L]; the five tables have n_steps entries.
 *   BVG_SAMPLE_PLMS (p_sample_plms :100-150): e' = the linear multistep combination of eps and the history
 *       combine 0: eps | 1: (h0 + eps) / 2 | 2: (3 eps - h0) / 2 | 3: (23 eps - 16 h0 + 5 h1) / 12
 *             | 4: (55 eps - 59 h0 + 37 h1 - 9 h2) / 24        (h0 = the most recent earlier prediction)
 *       x' = x + (a_prev - a_t) ((1 / (sqrt a_t (sqrt a_t + sqrt a_prev))) x
 *                 - (1 / (sqrt a_t (sqrt((1 - a_prev) a_t) + sqrt((1 - a_t) a_prev)))) e')       (get_x_pred :105-121)
 *     with a_t = alphas_cumprod[t], a_prev = alphas_cumprod[max(t - interval, 0)].
 * d_x_out may alias d_x.  d_eps_save (optional) receives a copy of eps (the sampler's noise_list entry). */
enum bvg_sample_mode { BVG_SAMPLE_DDPM = 0, BVG_SAMPLE_PLMS = 1 };

typedef struct bvg_sample_desc {
  int32_t mode; /* bvg_sample_mode */
  int32_t clip; /* DDPM: clamp the predicted x0 to [-1, 1] (clip_denoised) */
  const float* d_x;
  float* d_x_out;
  const float* d_eps;
  const float* d_noise;  /* DDPM: [B, n_mel, L] */
  const int32_t* d_step; /* [B] */
  const float* d_sqrt_recip;   /* DDPM tables, [n_steps] each */
  const float* d_sqrt_recipm1;
  const float* d_coef1;
  const float* d_coef2;
  const float* d_logvar;
  const float* d_alphas_cumprod; /* PLMS, [n_steps] */
  const float* d_hist[3];        /* PLMS: earlier predictions, most recent first */
  float* d_eps_save;
  int32_t interval, combine;
  int32_t B, L, n_mel, n_steps;
} bvg_sample_desc;

int bvg_sample_fwd(const bvg_sample_desc* d, void* stream);

/* ------------------------------------------------------------------------------------------
 * Log-mel front end (replaces mel_spectrogram, utils/mel.py:130-174, which the
 * reference runs with torch.stft + a librosa mel basis on the host: reflect pad (n_fft - hop) / 2, frames
 * every hop with center = False, periodic hann window of `win` samples, sqrt(re^2 + im^2 + 1e-9), basis
 * matmul, log(max(., clip))).
 *   d_wave : float [B, n] (row stride wave_stride);   d_out : float [B, n_mels, frames],
 *   frames = 1 + (n + 2 * ((n_fft - hop) / 2) - n_fft) / hop  (checked);
 *   d_basis: float [n_mels, n_fft / 2 + 1], the caller's mel basis (librosa.filters.mel for the reference);
 *   d_band : int32 [n_mels][2] = first and one-past-last non-zero bin of every band.
 * ------------------------------------------------------------------------------------------ */
typedef struct bvg_logmel_desc {
  const float* d_wave;
  int64_t wave_stride;
  float* d_out;
  const float* d_basis;
  const int32_t* d_band;
  int32_t B, n, n_fft, hop, win, n_mels, frames;
  float clip; /* 1e-5 in the reference (dynamic_range_compression_torch, utils/mel.py:25-26) */
} bvg_logmel_desc;

int bvg_logmel_fwd(const bvg_logmel_desc* d, void* stream);

/* Layout/format helpers used by tests and by the sharded stitch. */
int bvg_convert(const bvg_tensor* src, const bvg_tensor* dst, size_t n_elems, void* stream);

/* ------------------------------------------------------------------------------------------
 * Programs: a whole Generator.forward (modules/bigvgan.py:600-622) as one pre-validated launch
 * list, so the per-call host work is one C call (and the list can be captured in a CUDA graph).
 * ------------------------------------------------------------------------------------------ */
enum bvg_op_kind { BVG_OP_PACK = 0, BVG_OP_AMP = 1, BVG_OP_CONV = 2, BVG_OP_POST = 3, BVG_OP_ROWOP = 4, BVG_OP_DIFFEMBED = 5, BVG_OP_SAMPLE = 6 };
#define BVG_N_OP_KINDS 7

typedef struct bvg_op {
  int32_t kind; /* bvg_op_kind */
  int32_t _pad;
  union {
    bvg_pack_desc pack;
    bvg_amp_desc amp;
    bvg_conv_desc conv;
    bvg_post_desc post;
    bvg_rowop_desc rowop;
    bvg_diffembed_desc diffembed;
    bvg_sample_desc sample;
  } u;
} bvg_op;

typedef struct bvg_program bvg_program;

int bvg_program_create(const bvg_op* ops, int32_t n_ops, bvg_program** out);
int bvg_program_run(bvg_program* p, void* stream);
/* Same launches with a CUDA event between consecutive ops; synchronises the stream and returns the
 * device time (ms) and launch count per bvg_op_kind (arrays of BVG_N_OP_KINDS).  Measurement aid for bench.py. */
int bvg_program_run_timed(bvg_program* p, void* stream, float* ms_by_kind, int32_t* n_by_kind,
                          float* ms_per_op /* optional, n_ops entries */);
/* Issue n programs with the same op sequence op by op, program k on streams[k].  Each stream keeps its own
 * order; launching the ops alternately lets one half-batch's tensor-core convolutions (one persistent CTA per
 * SM, almost no issue slots) share the SMs with the other half-batch's FFMA-bound Activation1d kernels. */
int bvg_program_run_interleaved(bvg_program* const* progs, void* const* streams, int32_t n);
/* Chain the program's launches with programmatic dependent launch: every kernel is launched with
 * cudaLaunchAttributeProgrammaticStreamSerialization, signals `griddepcontrol.launch_dependents` on entry and
 * executes `griddepcontrol.wait` after its set-up and before its first global access, so that launch latency and
 * prologue (barrier init, TMEM allocation, tensor-map prefetch) of launch i + 1 overlap launch i.  Pays on
 * latency-bound programs (one utterance; a DiffSVC step); off by default. */
int bvg_program_set_pdl(bvg_program* p, int on);
int bvg_program_num_launches(const bvg_program* p);
void bvg_program_destroy(bvg_program* p);

/* Misc */
int bvg_abi_version(void);
const char* bvg_last_error(void);
int bvg_device_check(int device); /* BVG_OK iff `device` is compute capability 10.x */
size_t bvg_sizeof_op(void);       /* ABI self-check for the ctypes mirror */
size_t bvg_sizeof_conv_weights(void);

#ifdef __cplusplus
}
#endif
#endif /* BVG_B200_H_ */
