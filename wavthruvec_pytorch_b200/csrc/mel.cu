// Log-mel spectrogram of the generated waveform (vec2wav/dataset.py:53-77).
//
//   mel_frame_kernel : wav fp32 [B, L] -> reflect-padded hop-sample rows, fp16 high/low split
//                      [B, frames + taps - 1, 3 * hop] = (hi | 2^11 lo | hi)
//   v2w_tapgemm      : the STFT as one tcgen05 tap-GEMM; weights (W_hi | 2^-11 W | W_lo) per tap, so that
//                      x_hi W_hi + x_lo W + x_hi W_lo accumulates in TMEM (fp32)
//   mel_log_kernel   : sqrt(re^2 + im^2 + 1e-9) -> filterbank (only the non-zero span of each
//                      triangle) -> log(max(., 1e-5)), output [B, num_mels, frames]
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>
#include <vector>

#include "common.h"

namespace v2w {
namespace mel {

#define V2W_CUDA_M(expr)                                                                  \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return V2W_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

constexpr float kLoScale = 2048.f;   // the low part of a sample is carried as lo * 2^11, its weight block as W * 2^-11

// One thread per 4 samples of a padded row.  Padded sample s = row * hop + c maps to the original
// index i = s - pad, reflected at both ends (torch.nn.functional.pad mode='reflect').
__global__ void __launch_bounds__(256) mel_frame_kernel(const float* __restrict__ wav, __half* __restrict__ a,
                                                        int L, int hop, int rows, int pad) {
  const int b = blockIdx.y;
  const long long i4 = (long long)blockIdx.x * 256 + threadIdx.x;   // float4 index in the padded signal
  const long long total4 = (long long)rows * hop / 4;
  if (i4 >= total4) return;
  const long long s0 = i4 * 4;
  const int row = (int)(s0 / hop), c = (int)(s0 % hop);
  const float* w = wav + (size_t)b * L;
  __half hi[4], lo[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    long long i = s0 + k - pad;
    if (i < 0) i = -i;
    if (i >= L) i = 2LL * (L - 1) - i;
    float x = 0.f;
    if (i >= 0 && i < L) x = __ldg(w + i);
    hi[k] = __float2half_rn(x);
    lo[k] = __float2half_rn((x - __half2float(hi[k])) * kLoScale);   // exact power-of-two scaling keeps it out of the subnormals
  }
  __half* dst = a + ((size_t)b * rows + row) * (3 * hop) + c;
  *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<uint2*>(hi);
  *reinterpret_cast<uint2*>(dst + hop) = *reinterpret_cast<uint2*>(lo);
  *reinterpret_cast<uint2*>(dst + 2 * hop) = *reinterpret_cast<uint2*>(hi);
}

constexpr int kMelFrames = 8;     // frames per CTA (cfg2: 16 x 32 = 512 CTAs)
constexpr int kMelThreads = 256;

// spec fp32 [B, rows, ncols]: columns [0, nbins) = Re, [nbins, 2 nbins) = Im of frame `row`.
__global__ void __launch_bounds__(kMelThreads) mel_log_kernel(const float* __restrict__ spec,
                                                              const float* __restrict__ fb,      // [num_mels][nbins]
                                                              const int* __restrict__ span,      // [num_mels][2] non-zero bins
                                                              float* __restrict__ out, int rows, int frames,
                                                              int ncols, int nbins, int num_mels) {
  extern __shared__ float mag[];   // [nbins][kMelFrames + 1]
  const int b = blockIdx.y;
  const int f0 = blockIdx.x * kMelFrames;
  const int nf = min(kMelFrames, frames - f0);
  // one warp per frame, lanes stride the bins; kMelUnroll (re, im) pairs are in flight per thread
  // before the first use (the first version issued one dependent pair at a time: 0.6 TB/s)
  constexpr int kMelUnroll = 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int f = warp; f < kMelFrames; f += kMelThreads / 32) {
    if (f >= nf) {
      for (int k = lane; k < nbins; k += 32) mag[k * (kMelFrames + 1) + f] = 0.f;
      continue;
    }
    const float* row = spec + ((size_t)b * rows + f0 + f) * ncols;
    for (int k0 = lane; k0 < nbins; k0 += 32 * kMelUnroll) {
      float re[kMelUnroll], im[kMelUnroll];
#pragma unroll
      for (int u = 0; u < kMelUnroll; ++u) {
        const int k = k0 + 32 * u;
        re[u] = k < nbins ? __ldg(row + k) : 0.f;
        im[u] = k < nbins ? __ldg(row + nbins + k) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < kMelUnroll; ++u) {
        const int k = k0 + 32 * u;
        if (k < nbins) mag[k * (kMelFrames + 1) + f] = sqrtf(re[u] * re[u] + im[u] * im[u] + 1e-9f);
      }
    }
  }
  __syncthreads();
  const int f = threadIdx.x % kMelFrames;
  for (int m = threadIdx.x / kMelFrames; m < num_mels; m += kMelThreads / kMelFrames) {
    const int k0 = span[2 * m], k1 = span[2 * m + 1];
    float acc = 0.f;
    for (int k = k0; k < k1; ++k) acc = fmaf(__ldg(fb + (size_t)m * nbins + k), mag[k * (kMelFrames + 1) + f], acc);
    if (f < nf) out[((size_t)b * num_mels + m) * frames + f0 + f] = logf(fmaxf(acc, 1e-5f));
  }
}

// librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) with its defaults (htk=False, norm='slaney'), in
// float64 like librosa, cast to float32 at the end.
static double hz_to_mel(double f) {
  const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp;
  const double logstep = std::log(6.4) / 27.0;
  return f >= min_log_hz ? min_log_mel + std::log(f / min_log_hz) / logstep : f / f_sp;
}
static double mel_to_hz(double m) {
  const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp;
  const double logstep = std::log(6.4) / 27.0;
  return m >= min_log_mel ? min_log_hz * std::exp(logstep * (m - min_log_mel)) : f_sp * m;
}
static void slaney_filterbank(const v2w_mel_config& c, double fmax, std::vector<float>* fb) {
  const int nb = c.n_fft / 2 + 1, nm = c.num_mels;
  std::vector<double> mel_f(nm + 2);
  const double m0 = hz_to_mel(c.fmin), m1 = hz_to_mel(fmax);
  for (int i = 0; i < nm + 2; ++i) mel_f[i] = mel_to_hz(m0 + (m1 - m0) * i / (nm + 1));
  fb->assign((size_t)nm * nb, 0.f);
  for (int i = 0; i < nm; ++i) {
    const double enorm = 2.0 / (mel_f[i + 2] - mel_f[i]);
    for (int k = 0; k < nb; ++k) {
      const double fr = (double)c.sampling_rate / 2.0 * k / (nb - 1);
      const double lower = (fr - mel_f[i]) / (mel_f[i + 1] - mel_f[i]);
      const double upper = (mel_f[i + 2] - fr) / (mel_f[i + 2] - mel_f[i + 1]);
      (*fb)[(size_t)i * nb + k] = (float)(std::max(0.0, std::min(lower, upper)) * enorm);
    }
  }
}

}  // namespace mel
}  // namespace v2w

struct v2w_mel {
  v2w_mel_config cfg;
  int batch = 0, samples = 0, frames = 0, taps = 0, rows = 0, nbins = 0, ncols = 0, pad = 0;
  std::vector<float> fb;        // [num_mels][nbins]
  std::vector<int> span;        // [num_mels][2]
  int64_t off_a = 0, off_spec = 0, off_w = 0, off_fb = 0, off_span = 0, ws_bytes = 0;
  uint8_t* ws = nullptr;
  bool bound = false;
};

using namespace v2w;

extern "C" {

int32_t v2w_mel_create(const v2w_mel_config* cfg, int32_t batch, int32_t samples, v2w_mel** out) {
  if (!cfg || !out) {
    set_error("v2w_mel_create: null argument");
    return V2W_ERR_INVALID;
  }
  *out = nullptr;
  const v2w_mel_config& c = *cfg;
  if (c.n_fft < 32 || c.hop_size < 16 || c.hop_size % 16 || c.n_fft % c.hop_size || c.n_fft / c.hop_size > V2W_MAX_TAPS ||
      c.win_size < 1 || c.win_size > c.n_fft || c.num_mels < 1 || c.num_mels > 1024 || c.sampling_rate < 1 ||
      (c.n_fft - c.hop_size) % 2) {
    set_error("v2w_mel_create: unsupported n_fft=%d hop=%d win=%d mels=%d", c.n_fft, c.hop_size, c.win_size, c.num_mels);
    return V2W_ERR_INVALID;
  }
  const int pad = (c.n_fft - c.hop_size) / 2;
  if (batch < 1 || batch > 65535 || samples <= pad || samples < c.hop_size) {
    set_error("v2w_mel_create: batch=%d samples=%d out of range (reflect padding needs more than %d samples)",
              batch, samples, pad);
    return V2W_ERR_INVALID;
  }
  const double fmax = c.fmax > 0.f ? (double)c.fmax : c.sampling_rate / 2.0;
  if (c.fmin < 0.f || fmax <= c.fmin) {
    set_error("v2w_mel_create: fmin=%g fmax=%g", (double)c.fmin, fmax);
    return V2W_ERR_INVALID;
  }
  v2w_mel* M = new (std::nothrow) v2w_mel();
  if (!M) {
    set_error("out of memory");
    return V2W_ERR_NOMEM;
  }
  M->cfg = c;
  M->batch = batch; M->samples = samples; M->pad = pad;
  M->frames = (samples + 2 * pad - c.n_fft) / c.hop_size + 1;
  M->taps = c.n_fft / c.hop_size;
  M->rows = M->frames + M->taps - 1;
  M->nbins = c.n_fft / 2 + 1;
  M->ncols = (int)align_up(2 * M->nbins, 128);
  if (M->ncols / 128 > V2W_MAX_TILES) {
    delete M;
    set_error("v2w_mel_create: n_fft=%d needs more than %d column tiles", c.n_fft, V2W_MAX_TILES);
    return V2W_ERR_INVALID;
  }
  mel::slaney_filterbank(c, fmax, &M->fb);
  M->span.assign((size_t)2 * c.num_mels, 0);
  for (int m = 0; m < c.num_mels; ++m) {
    int k0 = M->nbins, k1 = 0;
    for (int k = 0; k < M->nbins; ++k)
      if (M->fb[(size_t)m * M->nbins + k] != 0.f) { k0 = std::min(k0, k); k1 = std::max(k1, k + 1); }
    if (k1 <= k0) k0 = k1 = 0;
    M->span[2 * m] = k0; M->span[2 * m + 1] = k1;
  }
  int64_t off = 0;
  auto take = [&](int64_t bytes) { const int64_t o = off; off = align_up(off + bytes, 256); return o; };
  M->off_a = take((int64_t)batch * M->rows * 3 * c.hop_size * 2);
  M->off_spec = take((int64_t)batch * M->rows * M->ncols * 4);
  M->off_w = take((int64_t)M->taps * M->ncols * 3 * c.hop_size * 2);
  M->off_fb = take((int64_t)c.num_mels * M->nbins * 4);
  M->off_span = take((int64_t)2 * c.num_mels * 4);
  M->ws_bytes = off;
  *out = M;
  return V2W_OK;
}

void v2w_mel_destroy(v2w_mel* m) { delete m; }
int32_t v2w_mel_num_frames(const v2w_mel* m) { return m ? m->frames : 0; }
int64_t v2w_mel_workspace_bytes(const v2w_mel* m) { return m ? m->ws_bytes : 0; }

int32_t v2w_mel_get_filterbank(const v2w_mel* m, float* out_host) {
  if (!m || !out_host) {
    set_error("v2w_mel_get_filterbank: null argument");
    return V2W_ERR_INVALID;
  }
  std::memcpy(out_host, m->fb.data(), m->fb.size() * 4);
  return V2W_OK;
}

int32_t v2w_mel_bind(v2w_mel* m, void* workspace_dev) {
  if (!m || !workspace_dev || (reinterpret_cast<uintptr_t>(workspace_dev) & 255)) {
    set_error("v2w_mel_bind: null or unaligned workspace");
    return V2W_ERR_INVALID;
  }
  const v2w_mel_config& c = m->cfg;
  const int hop = c.hop_size, K = 3 * hop;
  // windowed one-sided DFT basis, fp16 high / low parts: w16[tap][n][(hi | hi | lo)]
  std::vector<double> win(c.n_fft, 0.0);
  const int left = (c.n_fft - c.win_size) / 2;     // torch.stft centres a short window
  const double kPi = 3.14159265358979323846;
  for (int t = 0; t < c.win_size; ++t) win[left + t] = 0.5 - 0.5 * std::cos(2.0 * kPi * t / c.win_size);  // periodic Hann
  std::vector<uint16_t> w((size_t)m->taps * m->ncols * K, 0);
  for (int j = 0; j < m->taps; ++j)
    for (int n = 0; n < 2 * m->nbins; ++n) {
      const int k = n < m->nbins ? n : n - m->nbins;
      uint16_t* dst = w.data() + ((size_t)j * m->ncols + n) * K;
      for (int cc = 0; cc < hop; ++cc) {
        const int t = j * hop + cc;
        const long long ph = ((long long)k * t) % c.n_fft;        // exact phase reduction
        const double ang = 2.0 * kPi * (double)ph / c.n_fft;
        const double v = win[t] * (n < m->nbins ? std::cos(ang) : -std::sin(ang));
        const uint16_t hi = fp16_bits((float)v);
        const float hif = __half2float(__ushort_as_half(hi));
        const uint16_t lo = fp16_bits((float)(v - (double)hif));
        dst[cc] = hi; dst[hop + cc] = fp16_bits((float)(v / (double)mel::kLoScale)); dst[2 * hop + cc] = lo;
      }
    }
  uint8_t* ws = static_cast<uint8_t*>(workspace_dev);
  V2W_CUDA_M(cudaMemcpy(ws + m->off_w, w.data(), w.size() * 2, cudaMemcpyHostToDevice));
  V2W_CUDA_M(cudaMemcpy(ws + m->off_fb, m->fb.data(), m->fb.size() * 4, cudaMemcpyHostToDevice));
  V2W_CUDA_M(cudaMemcpy(ws + m->off_span, m->span.data(), m->span.size() * 4, cudaMemcpyHostToDevice));
  m->ws = ws;
  m->bound = true;
  return V2W_OK;
}

int32_t v2w_mel_forward(v2w_mel* m, const float* wav_dev, float* out_dev, void* stream, int32_t* gpu_launches) {
  if (!m || !wav_dev || !out_dev) {
    set_error("v2w_mel_forward: null argument");
    return V2W_ERR_INVALID;
  }
  if (!m->bound) {
    set_error("v2w_mel_forward: not bound to device memory");
    return V2W_ERR_STATE;
  }
  const v2w_mel_config& c = m->cfg;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int hop = c.hop_size;
  {
    const long long total4 = (long long)m->rows * hop / 4;
    dim3 grid((unsigned)((total4 + 255) / 256), m->batch);
    mel::mel_frame_kernel<<<grid, 256, 0, st>>>(wav_dev, reinterpret_cast<__half*>(m->ws + m->off_a), m->samples,
                                                hop, m->rows, m->pad);
    V2W_CUDA_M(cudaGetLastError());
  }
  v2w_tapgemm_args g;
  std::memset(&g, 0, sizeof(g));
  g.engine = V2W_ENGINE_TC;
  g.batch = m->batch; g.rows = m->rows; g.cin = 3 * hop; g.n_total = m->ncols; g.cmod = m->ncols;
  g.n_tile = 128; g.n_tiles = m->ncols / 128;
  for (int t = 0; t < g.n_tiles; ++t) {
    g.num_taps[t] = m->taps;
    for (int j = 0; j < m->taps; ++j) { g.tap_off[t][j] = j; g.tap_row[t][j] = j; }
  }
  g.in_act = m->ws + m->off_a; g.in_is_bf16 = 1; g.in_slope = 1.f;
  g.w16 = m->ws + m->off_w;
  g.out_scale = 1.f; g.out_slope = 1.f;
  g.out_f32 = reinterpret_cast<float*>(m->ws + m->off_spec);
  g.tc_flags = 0x100;   // fp16 operands
  int rc = v2w_tapgemm(&g, stream);
  if (rc) return rc;
  {
    const int smem = m->nbins * (mel::kMelFrames + 1) * 4;
    if (smem > 48 * 1024)   // per device, and only n_fft > 2048 needs it: not worth caching
      V2W_CUDA_M(cudaFuncSetAttribute(mel::mel_log_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    if (smem > 200 * 1024) {
      set_error("v2w_mel_forward: n_fft=%d too large for the filterbank kernel", c.n_fft);
      return V2W_ERR_INVALID;
    }
    dim3 grid((m->frames + mel::kMelFrames - 1) / mel::kMelFrames, m->batch);
    mel::mel_log_kernel<<<grid, mel::kMelThreads, smem, st>>>(
        reinterpret_cast<const float*>(m->ws + m->off_spec), reinterpret_cast<const float*>(m->ws + m->off_fb),
        reinterpret_cast<const int*>(m->ws + m->off_span), out_dev, m->rows, m->frames, m->ncols, m->nbins, c.num_mels);
    V2W_CUDA_M(cudaGetLastError());
  }
  if (gpu_launches) *gpu_launches = 3;
  return V2W_OK;
}

}  // extern "C"
