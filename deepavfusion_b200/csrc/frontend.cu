// Input stage in front of the hot path: the reference builds its log-mel spectrograms and normalised
// frames on CPU data-loader workers (train.py:44-54: torchaudio MelSpectrogram(16 kHz, n_fft 800, hop 250, 128 mels) ->
// log10(x + 1e-7); torchvision ToTensor + Normalize) and ships fp32 tensors over PCIe.  Here int16 PCM and uint8 frames
// cross PCIe (4x / 4x fewer bytes) and both transforms run on the GPU.
//
// log-mel, one CTA per (clip, group of 8 frames), all f32:
//   frame t = samples [250 t - 400, 250 t + 400) of the reflect-padded clip (torch.stft centre = True), times the periodic
//   Hann window; 800-point real DFT as a 25 x 32 Cooley-Tukey step in shared memory
//       n = 32 n1 + n2,  k = k1 + 25 k2:   X[k] = sum_n2 W800^(n2 k1) W32^(n2 k2) [ sum_n1 x[32 n1 + n2] W25^(n1 k1) ]
//   (95 k FMA per frame instead of 641 k for the direct transform; only k <= 400 is produced), power, the 128 HTK
//   triangular mel filters (each thread owns one filter and walks its own frequency span), log10(. + 1e-7).
// The twiddle tables, window and filter bank live in a caller-provided workspace filled once by davf_logmel_init.
#include <math.h>
#include "common.cuh"

namespace davf {

constexpr int kNfft = 800, kHop = 250, kN1 = 25, kN2 = 32, kBins = 401, kFramesPerCta = 8, kMaxMels = 128;

struct LogmelTables {           // layout of the workspace
  float2 w25[kN1 * kN1];        // [k1][n1]  exp(-2 pi i n1 k1 / 25)
  float2 tw[kN2 * kN1];         // [n2][k1]  exp(-2 pi i n2 k1 / 800)
  float2 w32[17 * kN2];         // [k2][n2]  exp(-2 pi i n2 k2 / 32), k2 <= 16
  float window[kNfft];
  int2 span[kMaxMels];          // first / last frequency bin with a non-zero weight
  float fb[kBins * kMaxMels];   // [bin][mel]
};

__global__ void logmel_init_kernel(LogmelTables* t, int sample_rate, int n_mels) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const double two_pi = 6.283185307179586476925286766559;
  if (i < kN1 * kN1) { const int k1 = i / kN1, n1 = i % kN1; double s, c; sincos(-two_pi * ((n1 * k1) % kN1) / kN1, &s, &c); t->w25[i] = make_float2((float)c, (float)s); }
  if (i < kN2 * kN1) { const int n2 = i / kN1, k1 = i % kN1; double s, c; sincos(-two_pi * (n2 * k1) / kNfft, &s, &c); t->tw[i] = make_float2((float)c, (float)s); }
  if (i < 17 * kN2) { const int k2 = i / kN2, n2 = i % kN2; double s, c; sincos(-two_pi * ((n2 * k2) % kN2) / kN2, &s, &c); t->w32[i] = make_float2((float)c, (float)s); }
  if (i < kNfft) t->window[i] = (float)(0.5 - 0.5 * cos(two_pi * i / kNfft));
  // torchaudio.functional.melscale_fbanks(n_freqs = 401, f_min = 0, f_max = sr / 2, norm = None, mel_scale = 'htk')
  const double f_max = sample_rate / 2.0, m_max = 2595.0 * log10(1.0 + f_max / 700.0);
  auto f_pt = [&](int j) { return 700.0 * (pow(10.0, (m_max * j / (n_mels + 1)) / 2595.0) - 1.0); };
  if (i < kBins * kMaxMels) {
    const int bin = i / kMaxMels, m = i % kMaxMels;
    float v = 0.f;
    if (m < n_mels) {
      const double f = (double)(sample_rate / 2) * bin / (kBins - 1);
      const double lo = f_pt(m), mid = f_pt(m + 1), hi = f_pt(m + 2);
      const double down = (f - lo) / (mid - lo), up = (hi - f) / (hi - mid);
      v = (float)fmax(0.0, fmin(down, up));
    }
    t->fb[i] = v;
  }
  if (i < kMaxMels) {
    int lo = kBins, hi = -1;
    if (i < n_mels) {
      const double flo = f_pt(i), fhi = f_pt(i + 2), df = (double)(sample_rate / 2) / (kBins - 1);
      lo = max(0, (int)floor(flo / df));
      hi = min(kBins - 1, (int)ceil(fhi / df));
    }
    t->span[i] = make_int2(lo, hi);
  }
}

// wave: f32 [B, T] or i16 [B, T] PCM; gain_db: per-clip gain (RandomVol, audio_transforms.py:8-18) or NULL;
// out f32 [B, 1, n_mels, frames] with frames <= T / hop + 1 (the reference keeps T / hop: datasets.py:242)
__global__ void __launch_bounds__(256) logmel_kernel(const LogmelTables* __restrict__ t, const float* __restrict__ wave_f32,
                                                     const int16_t* __restrict__ wave_i16, const float* __restrict__ gain_db,
                                                     int T, int n_mels, int frames, float eps, float* __restrict__ out) {
  __shared__ float xs[kNfft];
  __shared__ float2 ys[kN2 * kN1];
  __shared__ float pw[kBins + 3];
  __shared__ float mel[kFramesPerCta][kMaxMels];
  const int b = blockIdx.y, f0 = blockIdx.x * kFramesPerCta, tid = threadIdx.x;
  const float gain = gain_db ? exp10f(gain_db[b] * 0.05f) : 1.0f;
  for (int fi = 0; fi < kFramesPerCta; ++fi) {
    const int f = f0 + fi;
    if (f >= frames) break;
    for (int n = tid; n < kNfft; n += blockDim.x) {
      int j = f * kHop - kNfft / 2 + n;
      j = j < 0 ? -j : (j >= T ? 2 * (T - 1) - j : j);               // reflect padding
      float v = wave_i16 ? (float)wave_i16[(int64_t)b * T + j] * (1.0f / 32768.0f) : wave_f32[(int64_t)b * T + j];
      if (gain_db) v = fminf(1.0f, fmaxf(-1.0f, v * gain));
      xs[n] = v * t->window[n];
    }
    __syncthreads();
    for (int o = tid; o < kN2 * kN1; o += blockDim.x) {               // stage 1 + twiddle: Y[n2][k1]
      const int n2 = o / kN1, k1 = o % kN1;
      float re = 0.f, im = 0.f;
#pragma unroll 5
      for (int n1 = 0; n1 < kN1; ++n1) {
        const float2 w = t->w25[k1 * kN1 + n1];
        const float x = xs[kN2 * n1 + n2];
        re = fmaf(x, w.x, re);
        im = fmaf(x, w.y, im);
      }
      const float2 w = t->tw[o];
      ys[o] = make_float2(re * w.x - im * w.y, re * w.y + im * w.x);
    }
    __syncthreads();
    for (int k = tid; k < kBins; k += blockDim.x) {                   // stage 2: X[k1 + 25 k2], power
      const int k2 = k / kN1, k1 = k % kN1;
      float re = 0.f, im = 0.f;
#pragma unroll 8
      for (int n2 = 0; n2 < kN2; ++n2) {
        const float2 y = ys[n2 * kN1 + k1], w = t->w32[k2 * kN2 + n2];
        re += y.x * w.x - y.y * w.y;
        im += y.x * w.y + y.y * w.x;
      }
      pw[k] = re * re + im * im;
    }
    __syncthreads();
    if (tid < n_mels) {
      const int2 sp = t->span[tid];
      float acc = 0.f;
      for (int k = sp.x; k <= sp.y; ++k) acc = fmaf(pw[k], t->fb[k * kMaxMels + tid], acc);
      mel[fi][tid] = log10f(acc + eps);
    }
    __syncthreads();
  }
  const int nf = min(kFramesPerCta, frames - f0);
  for (int i = tid; i < n_mels * kFramesPerCta; i += blockDim.x) {
    const int m = i / kFramesPerCta, fi = i % kFramesPerCta;
    if (fi < nf) out[((int64_t)b * n_mels + m) * frames + f0 + fi] = mel[fi][m];
  }
}

// uint8 [B, H, W, C] frames -> f32 [B, C, H, W] normalised (ToTensor + Normalize, train.py:48-49); C <= 4
__global__ void image_normalize_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, int64_t pixels_per_image, int C,
                                       float4 mean, float4 inv_std, int64_t total) {
  const float mu[4] = {mean.x, mean.y, mean.z, mean.w}, is[4] = {inv_std.x, inv_std.y, inv_std.z, inv_std.w};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / pixels_per_image, px = i - b * pixels_per_image;       // one thread per pixel: C reads, C coalesced writes
    for (int c = 0; c < C; ++c)
      dst[(b * C + c) * pixels_per_image + px] = ((float)src[i * C + c] * (1.0f / 255.0f) - mu[c]) * is[c];
  }
}

}  // namespace davf

using namespace davf;

extern "C" int64_t davf_logmel_workspace_bytes(void) { return (int64_t)sizeof(LogmelTables); }

extern "C" int davf_logmel_init(void* workspace, int sample_rate, int n_fft, int hop, int n_mels, davf_stream_t s) {
  DAVF_CHECK_ARG(workspace && ((uintptr_t)workspace & 15) == 0, "logmel_init: workspace must be 16-byte aligned");
  DAVF_CHECK_ARG(n_fft == kNfft && hop == kHop, "logmel: only n_fft = 800 / hop = 250 (16 kHz, train.py:53) is built, got %d / %d", n_fft, hop);
  DAVF_CHECK_ARG(n_mels >= 1 && n_mels <= kMaxMels && sample_rate > 0, "logmel: n_mels=%d sample_rate=%d", n_mels, sample_rate);
  const int n = kBins * kMaxMels;
  logmel_init_kernel<<<(n + 255) / 256, 256, 0, as_stream(s)>>>(reinterpret_cast<LogmelTables*>(workspace), sample_rate, n_mels);
  DAVF_LAUNCH_OK();
  return DAVF_OK;
}

extern "C" int davf_logmel_fwd(const void* workspace, const float* wave_f32, const int16_t* wave_i16, const float* gain_db, int B, int T,
                               int n_mels, int frames, float eps, float* out, davf_stream_t s) {
  DAVF_CHECK_ARG(workspace && out && ((wave_f32 != nullptr) != (wave_i16 != nullptr)), "logmel: pass exactly one of wave_f32 / wave_i16");
  DAVF_CHECK_ARG(T > kNfft / 2 && frames >= 1 && frames <= T / kHop + 1 && n_mels >= 1 && n_mels <= kMaxMels, "logmel: T=%d frames=%d n_mels=%d", T, frames, n_mels);
  if (B == 0) return DAVF_OK;
  dim3 grid((frames + kFramesPerCta - 1) / kFramesPerCta, B);
  logmel_kernel<<<grid, 256, 0, as_stream(s)>>>(reinterpret_cast<const LogmelTables*>(workspace), wave_f32, wave_i16, gain_db, T, n_mels, frames, eps, out);
  DAVF_LAUNCH_OK();
  return DAVF_OK;
}

extern "C" int davf_image_normalize_u8(const uint8_t* src, float* dst, int B, int H, int W, int C, const float* mean, const float* std_,
                                       davf_stream_t s) {
  DAVF_CHECK_ARG(src && dst && mean && std_ && C >= 1 && C <= 4, "image_normalize: null pointer or C=%d", C);
  if (B == 0) return DAVF_OK;
  float mu[4] = {0, 0, 0, 0}, is[4] = {1, 1, 1, 1};
  for (int c = 0; c < C; ++c) { mu[c] = mean[c]; is[c] = 1.0f / std_[c]; }
  const int64_t ppi = (int64_t)H * W, total = ppi * B;
  const int grid = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  image_normalize_kernel<<<grid, 256, 0, as_stream(s)>>>(src, dst, ppi, C, make_float4(mu[0], mu[1], mu[2], mu[3]), make_float4(is[0], is[1], is[2], is[3]), total);
  DAVF_LAUNCH_OK();
  return DAVF_OK;
}
