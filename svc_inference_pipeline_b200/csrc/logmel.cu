// Log-mel front end on the device: the analysis the reference runs on the host before the
// acoustic model and whose output range the vocoder consumes -- reference utils/mel.py:130-174 mel_spectrogram:
//   reflect-pad (n_fft - hop) / 2 samples each side (:148-153), frames of n_fft samples every hop (center = False),
//   periodic hann window (:146), |rFFT| as sqrt(re^2 + im^2 + 1e-9) (:156-169), mel basis matmul (:171),
//   log(clamp(., 1e-5)) (:172, :25-26).
// One CTA transforms TWO frames per complex FFT (frame A in the real part, frame B in the imaginary part; the two
// spectra separate as X_A[k] = (Z[k] + conj Z[N-k]) / 2, X_B[k] = (Z[k] - conj Z[N-k]) / 2i), radix-2 in shared
// memory with a twiddle table built once per CTA, then each thread owns one mel band and sums its triangle (the basis
// is the caller's dense [n_mels, n_fft/2 + 1] matrix -- librosa's slaney bank for the reference -- with the non-zero
// range of every band passed alongside, so the sum skips the zeros but uses exactly the caller's coefficients).
// Persistent grid: a CTA walks frame pairs with stride gridDim.x.  This is a bandwidth-trivial stage (2 KB in, 400 B
// out per frame); it exists so that the log-mel parity metric and analysis -> synthesis round trips stay on the GPU.
#include "common.cuh"

namespace bvg {

struct LogmelParams {
  const float* wave;
  long long wave_stride;
  float* out;
  const float* basis;
  const int* band;
  int B, n, n_fft, log2n, hop, win, n_mels, frames, pad;
  float clip;
  long long pairs_per_item, total_pairs;
};

__device__ __forceinline__ int reflect_index(int i, int n) {
  // torch "reflect" padding (no edge repeat); pad < n is checked on the host
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

__global__ void __launch_bounds__(256) logmel_kernel(const __grid_constant__ LogmelParams p) {
  extern __shared__ float sm[];
  const int N = p.n_fft, H = N >> 1, NB = H + 1;
  float* re = sm;            // [N]
  float* im = re + N;        // [N]
  float* twc = im + N;       // [H] cos(2 pi k / N)
  float* tws = twc + H;      // [H] -sin(2 pi k / N)
  float* win = tws + H;      // [N] window, zero outside the centred win samples
  float* magA = win + N;     // [NB]
  float* magB = magA + NB;   // [NB]
  const int tid = threadIdx.x, nt = blockDim.x;

  for (int k = tid; k < H; k += nt) {
    float s, c;
    sincospif(2.0f * (float)k / (float)N, &s, &c);
    twc[k] = c;
    tws[k] = -s;
  }
  const int wl = (N - p.win) >> 1;  // torch.stft centres a shorter window inside n_fft
  for (int k = tid; k < N; k += nt) {
    const int j = k - wl;
    win[k] = (j >= 0 && j < p.win) ? 0.5f - 0.5f * cospif(2.0f * (float)j / (float)p.win) : 0.f;
  }
  __syncthreads();

  for (long long pair = blockIdx.x; pair < p.total_pairs; pair += gridDim.x) {
    const int b = (int)(pair / p.pairs_per_item);
    const int fA = (int)(pair % p.pairs_per_item) * 2, fB = fA + 1;
    const bool hasB = fB < p.frames;
    const float* w = p.wave + (long long)b * p.wave_stride;
    // windowed frames into bit-reversed positions
    for (int k = tid; k < N; k += nt) {
      const int r = (int)(__brev((unsigned)k) >> (32 - p.log2n));
      const float wk = win[k];
      re[r] = wk * __ldg(w + reflect_index(fA * p.hop + k - p.pad, p.n));
      im[r] = hasB ? wk * __ldg(w + reflect_index(fB * p.hop + k - p.pad, p.n)) : 0.f;
    }
    __syncthreads();
    for (int s = 0; s < p.log2n; ++s) {
      const int half = 1 << s;
      const int tstep = H >> s;  // twiddle index stride: W_{2 half}^pos = W_N^{pos * N / (2 half)}
      for (int j = tid; j < H; j += nt) {
        const int pos = j & (half - 1);
        const int i0 = ((j >> s) << (s + 1)) + pos, i1 = i0 + half;
        const float c = twc[pos * tstep], sn = tws[pos * tstep];
        const float xr = re[i1], xi = im[i1];
        const float tr = xr * c - xi * sn, ti = xr * sn + xi * c;
        const float ur = re[i0], ui = im[i0];
        re[i0] = ur + tr;
        im[i0] = ui + ti;
        re[i1] = ur - tr;
        im[i1] = ui - ti;
      }
      __syncthreads();
    }
    // separate the two real spectra and take magnitudes
    for (int k = tid; k < NB; k += nt) {
      const int nk = (N - k) & (N - 1);
      const float zr = re[k], zi = im[k], yr = re[nk], yi = im[nk];
      const float ar = 0.5f * (zr + yr), ai = 0.5f * (zi - yi);   // X_A[k]
      const float br = 0.5f * (zi + yi), bi = 0.5f * (yr - zr);   // X_B[k]
      magA[k] = sqrtf(ar * ar + ai * ai + 1e-9f);
      magB[k] = sqrtf(br * br + bi * bi + 1e-9f);
    }
    __syncthreads();
    for (int m = tid; m < p.n_mels; m += nt) {
      const int lo = __ldg(p.band + 2 * m), hi = __ldg(p.band + 2 * m + 1);
      const float* row = p.basis + (long long)m * NB;
      float sa = 0.f, sb = 0.f;
      for (int k = lo; k < hi; ++k) {
        const float c = __ldg(row + k);
        sa = fmaf(c, magA[k], sa);
        sb = fmaf(c, magB[k], sb);
      }
      float* o = p.out + ((long long)b * p.n_mels + m) * p.frames;
      o[fA] = logf(fmaxf(sa, p.clip));
      if (hasB) o[fB] = logf(fmaxf(sb, p.clip));
    }
    __syncthreads();
  }
}

int logmel_forward(const bvg_logmel_desc* d, cudaStream_t st) {
  BVG_REQUIRE(d && d->d_wave && d->d_out && d->d_basis && d->d_band, "logmel: null pointer");
  BVG_REQUIRE(d->B > 0 && d->n > 0 && d->n_mels > 0 && d->hop > 0, "logmel: bad shape");
  int log2n = 0;
  while ((1 << log2n) < d->n_fft) ++log2n;
  BVG_REQUIRE((1 << log2n) == d->n_fft && d->n_fft >= 64 && d->n_fft <= 4096, "logmel: n_fft %d must be a power of two in [64, 4096]", d->n_fft);
  BVG_REQUIRE(d->win > 0 && d->win <= d->n_fft, "logmel: window %d does not fit n_fft %d", d->win, d->n_fft);
  const int pad = (d->n_fft - d->hop) / 2;  // int((n_fft - hop_size) / 2), utils/mel.py:150
  BVG_REQUIRE(d->n_fft >= d->hop, "logmel: hop %d exceeds n_fft %d", d->hop, d->n_fft);
  BVG_REQUIRE(pad < d->n, "logmel: reflect padding of %d samples needs a longer waveform (%d samples)", pad, d->n);
  const int frames = 1 + (d->n + 2 * pad - d->n_fft) / d->hop;
  BVG_REQUIRE(d->n + 2 * pad >= d->n_fft && frames == d->frames, "logmel: %d samples give %d frames, the descriptor says %d", d->n, frames, d->frames);
  BVG_REQUIRE(d->wave_stride >= d->n, "logmel: wave_stride smaller than the waveform");
  LogmelParams p;
  p.wave = d->d_wave;
  p.wave_stride = d->wave_stride;
  p.out = d->d_out;
  p.basis = d->d_basis;
  p.band = d->d_band;
  p.B = d->B;
  p.n = d->n;
  p.n_fft = d->n_fft;
  p.log2n = log2n;
  p.hop = d->hop;
  p.win = d->win;
  p.n_mels = d->n_mels;
  p.frames = frames;
  p.pad = pad;
  p.clip = d->clip;
  p.pairs_per_item = (frames + 1) / 2;
  p.total_pairs = p.pairs_per_item * d->B;
  const size_t smem = sizeof(float) * ((size_t)d->n_fft * 4 + 2 * (d->n_fft / 2 + 1));
  if (smem > 48 * 1024 && first_use_on_device(reinterpret_cast<const void*>(logmel_kernel)))
    BVG_CHECK_CUDA(cudaFuncSetAttribute(logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long grid = p.total_pairs < (long long)sms * 8 ? p.total_pairs : (long long)sms * 8;
  logmel_kernel<<<(unsigned)grid, 256, smem, st>>>(p);
  BVG_CHECK_CUDA(cudaGetLastError());
  return BVG_OK;
}

}  // namespace bvg
