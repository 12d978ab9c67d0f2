// FFMA tap-GEMM (fp32 accumulate from fp32 or bf16 activations) and the small memory-bound
// kernels of the generator: input transpose, conditioning affine, conv_post + tanh.
//
// The FFMA tap-GEMM has exactly the semantics of the tcgen05 one (kernels_tc.cuh); it is the
// engine of V2W_PREC_FP32 plans (parity bar max-abs <= 1e-4, which a single-pass tensor-core
// format cannot meet: SURVEY.md section 7 "Hard parts 1") and the on-device cross-check for
// the tensor-core engine.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "v2w.h"

namespace v2w {
namespace simt {

constexpr int kBM = 64;        // rows per CTA
constexpr int kKC = 16;        // channels per K-chunk
constexpr int kThreads = 128;
constexpr int kMaxHalo = 64;   // (num_taps-1)*dilation upper bound supported

struct Params {
  const int* row_limit;   // variable-length batches: frames utterance b needs (nullptr = all)
  int rows_per_frame;
  int batch, rows, cin, n_total, cmod, n_tile;
  int num_taps[V2W_MAX_TILES];
  int off_min[V2W_MAX_TILES];
  int halo[V2W_MAX_TILES];
  int tap_off[V2W_MAX_TILES][V2W_MAX_TAPS];
  int tap_row[V2W_MAX_TILES][V2W_MAX_TAPS];
  const void* in_act;
  float in_slope;
  const float* w32;        // [taps][cin][n_total]
  const float* bias;
  const float* scale;
  const float* shift;
  const float* residual;   // may alias out_f32
  const float* accum_in;   // may alias out_f32
  float out_scale;
  float* out_f32;
  __nv_bfloat16* out_bf16;
  float out_slope;
};

__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }

__device__ __forceinline__ float4 load4(const float* p) {
  return *reinterpret_cast<const float4*>(p);
}
__device__ __forceinline__ float4 load4(const __nv_bfloat16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
  const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
  const float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}

// BN: columns per CTA (16, 32 or 64).  Thread layout: BN/4 column groups x 128/(BN/4) row
// groups; each thread owns TM x 4 outputs.
template <typename TIn, int BN>
__global__ void __launch_bounds__(kThreads) tapgemm_simt_kernel(const Params p) {
  constexpr int CG = BN / 4;
  constexpr int RG = kThreads / CG;
  constexpr int TM = kBM / RG;
  constexpr int AS = kKC + 1;  // padded row stride: conflict-free column reads
  __shared__ float As[(kBM + kMaxHalo) * AS];
  __shared__ __align__(16) float Ws[kKC * BN];

  const int tid = threadIdx.x;
  const int cg = tid % CG, rg = tid / CG;
  const int q0 = blockIdx.x * kBM;
  const int n0 = blockIdx.y * BN;
  const int b = blockIdx.z;
  if (p.row_limit != nullptr && q0 >= __ldg(p.row_limit + b) * p.rows_per_frame) return;   // whole CTA
  const int nt = n0 / p.n_tile;
  const int ntaps = p.num_taps[nt];
  const int off_min = p.off_min[nt];
  const int a_rows = kBM + p.halo[nt];
  const TIn* in = reinterpret_cast<const TIn*>(p.in_act) + static_cast<size_t>(b) * p.rows * p.cin;

  float acc[TM][4];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[i][e] = 0.f;

  for (int c0 = 0; c0 < p.cin; c0 += kKC) {
    __syncthreads();
    // A halo tile: a_rows x 16 channels, 4 channels per thread-load, leaky-relu on load
    for (int idx = tid; idx < a_rows * (kKC / 4); idx += kThreads) {
      const int r = idx / (kKC / 4), c4 = (idx % (kKC / 4)) * 4;
      const int row = q0 + off_min + r;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row >= 0 && row < p.rows) {
        v = load4(in + static_cast<size_t>(row) * p.cin + c0 + c4);
        v.x = lrelu(v.x, p.in_slope); v.y = lrelu(v.y, p.in_slope);
        v.z = lrelu(v.z, p.in_slope); v.w = lrelu(v.w, p.in_slope);
      }
      float* d = As + r * AS + c4;
      d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
    for (int j = 0; j < ntaps; ++j) {
      __syncthreads();
      const float* wsrc =
          p.w32 + (static_cast<size_t>(p.tap_row[nt][j]) * p.cin + c0) * p.n_total + n0;
      for (int idx = tid; idx < kKC * (BN / 4); idx += kThreads) {
        const int kk = idx / (BN / 4), c4 = (idx % (BN / 4)) * 4;
        *reinterpret_cast<float4*>(Ws + kk * BN + c4) =
            __ldg(reinterpret_cast<const float4*>(wsrc + static_cast<size_t>(kk) * p.n_total + c4));
      }
      __syncthreads();
      const int shift = p.tap_off[nt][j] - off_min;
      const float* arow = As + (rg * TM + shift) * AS;
#pragma unroll
      for (int kk = 0; kk < kKC; ++kk) {
        const float4 w = *reinterpret_cast<const float4*>(Ws + kk * BN + cg * 4);
#pragma unroll
        for (int i = 0; i < TM; ++i) {
          const float a = arow[i * AS + kk];
          acc[i][0] = fmaf(a, w.x, acc[i][0]);
          acc[i][1] = fmaf(a, w.y, acc[i][1]);
          acc[i][2] = fmaf(a, w.z, acc[i][2]);
          acc[i][3] = fmaf(a, w.w, acc[i][3]);
        }
      }
    }
  }

  const int col = n0 + cg * 4;
  const int cb = col % p.cmod;
  float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p.bias) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + cb));
  float4 sc4 = make_float4(1.f, 1.f, 1.f, 1.f), sh4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p.scale) {
    sc4 = __ldg(reinterpret_cast<const float4*>(p.scale + static_cast<size_t>(b) * p.cmod + cb));
    sh4 = __ldg(reinterpret_cast<const float4*>(p.shift + static_cast<size_t>(b) * p.cmod + cb));
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int row = q0 + rg * TM + i;
    if (row >= p.rows) continue;
    const size_t o = (static_cast<size_t>(b) * p.rows + row) * p.n_total + col;
    float v0 = acc[i][0] + bias4.x, v1 = acc[i][1] + bias4.y, v2 = acc[i][2] + bias4.z,
          v3 = acc[i][3] + bias4.w;
    if (p.scale) {
      v0 = fmaf(v0, sc4.x, sh4.x); v1 = fmaf(v1, sc4.y, sh4.y);
      v2 = fmaf(v2, sc4.z, sh4.z); v3 = fmaf(v3, sc4.w, sh4.w);
    }
    if (p.residual) {
      const float4 t = *reinterpret_cast<const float4*>(p.residual + o);
      v0 += t.x; v1 += t.y; v2 += t.z; v3 += t.w;
    }
    if (p.accum_in) {
      const float4 t = *reinterpret_cast<const float4*>(p.accum_in + o);
      v0 += t.x; v1 += t.y; v2 += t.z; v3 += t.w;
    }
    if (p.out_scale != 1.0f) {
      v0 *= p.out_scale; v1 *= p.out_scale; v2 *= p.out_scale; v3 *= p.out_scale;
    }
    if (p.out_f32) *reinterpret_cast<float4*>(p.out_f32 + o) = make_float4(v0, v1, v2, v3);
    if (p.out_bf16) {
      const __nv_bfloat162 h0 =
          __floats2bfloat162_rn(lrelu(v0, p.out_slope), lrelu(v1, p.out_slope));
      const __nv_bfloat162 h1 =
          __floats2bfloat162_rn(lrelu(v2, p.out_slope), lrelu(v3, p.out_slope));
      uint2 u;
      u.x = *reinterpret_cast<const uint32_t*>(&h0);
      u.y = *reinterpret_cast<const uint32_t*>(&h1);
      *reinterpret_cast<uint2*>(p.out_bf16 + o) = u;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// x[B, C, T] fp32 (channels-first, what collate_fn_tensor emits: vec2wav/dataset.py:212-213)
//   -> [B, T, C] channels-last, fp32 or bf16.  32x32 tiles through shared memory so that both
// the global read (along T) and the global write (along C) are coalesced.
// ---------------------------------------------------------------------------------------------
template <typename TOut> __device__ __forceinline__ TOut cvt_out(float v) { return static_cast<TOut>(v); }
template <> __device__ __forceinline__ __half cvt_out<__half>(float v) {
  return __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
}

constexpr int kMaxRaggedBatch = 512;
struct RaggedOffsets {          // by-value kernel parameter (constant bank): no upload, no sync
  int off[kMaxRaggedBatch + 1];
};

// Variable-length input: features arrive as packed time-major rows [sum_b T_b, C] -- the
// layout a dataset item already has (vec2wav/dataset.py:181-187) -- and utterance b owns rows
// [off[b], off[b + 1]).  Writes the padded channels-last activation y[b][t][:] for t < T with zeros
// past the utterance's end: the zero padding of collate_fn_tensor (dataset.py:212) without the
// [B, C, T] detour and its two transposes.
template <typename TOut>
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ x,
                                                          const RaggedOffsets ro,
                                                          TOut* __restrict__ y, int C, int T,
                                                          int* __restrict__ row_limit, int halo, int lo_off = 0) {
  const int b = blockIdx.y;
  const int first = ro.off[b];
  const int len = ro.off[b + 1] - first;
  // frames of utterance b that later kernels have to compute: its own plus the generator's right
  // receptive field (rows past that cannot reach a valid sample; their tiles are skipped)
  if (blockIdx.x == 0 && threadIdx.x == 0) row_limit[b] = min(len + halo, T);
  const int c4n = C >> 2;
  const int t0 = blockIdx.x * 8;
  for (int i = threadIdx.x; i < 8 * c4n; i += 256) {
    const int t = t0 + i / c4n, c4 = i % c4n;
    if (t >= T) break;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t < len) v = __ldg(reinterpret_cast<const float4*>(x + (static_cast<size_t>(first) + t) * C) + c4);
    // lo_off > 0 (split-fp16 operands): rows of [hi(C) | lo(C)], lo = the part of the value hi does not hold
    TOut* dst = y + (static_cast<size_t>(b) * T + t) * (lo_off ? 2 * C : C) + 4 * c4;
    const TOut h0 = cvt_out<TOut>(v.x), h1 = cvt_out<TOut>(v.y), h2 = cvt_out<TOut>(v.z), h3 = cvt_out<TOut>(v.w);
    dst[0] = h0; dst[1] = h1; dst[2] = h2; dst[3] = h3;
    if (lo_off) {
      dst[lo_off] = cvt_out<TOut>(v.x - static_cast<float>(h0));
      dst[lo_off + 1] = cvt_out<TOut>(v.y - static_cast<float>(h1));
      dst[lo_off + 2] = cvt_out<TOut>(v.z - static_cast<float>(h2));
      dst[lo_off + 3] = cvt_out<TOut>(v.w - static_cast<float>(h3));
    }
  }
}

template <typename TOut>
__global__ void __launch_bounds__(256) transpose_in_kernel(const float* __restrict__ x,
                                                           TOut* __restrict__ y, int C, int T, int lo_off = 0) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const float* xb = x + static_cast<size_t>(b) * C * T;
  const int pitch = lo_off ? 2 * C : C;     // lo_off > 0 (split-fp16 operands): rows of [hi(C) | lo(C)]
  TOut* yb = y + static_cast<size_t>(b) * pitch * T;
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int c = c0 + ty + i, t = t0 + tx;
    tile[ty + i][tx] = (c < C && t < T) ? xb[static_cast<size_t>(c) * T + t] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int t = t0 + ty + i, c = c0 + tx;
    if (t < T && c < C) {
      const float v = tile[tx][ty + i];
      const TOut hi = cvt_out<TOut>(v);
      yb[static_cast<size_t>(t) * pitch + c] = hi;
      if (lo_off) yb[static_cast<size_t>(t) * pitch + lo_off + c] = cvt_out<TOut>(v - static_cast<float>(hi));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Conditioning affine, one CTA per (stage, utterance)   (models.py:120,131-134; modules.py:23-28)
//   z      = fcs[i](cat(spk, noise))                       [128]
//   g|b    = (W_orig / sigma) z + bias                     [2C]   (sigma folded at pack time)
//   scale  = g * rsqrt(var + eps)     shift = b - scale * mean
// so that CBN(x) = scale * x + shift, applied in the epilogue of ups[i].
// ---------------------------------------------------------------------------------------------
struct CondStage {
  const float* fcw; const float* fcb; const float* snw; const float* snb;
  const float* rstd; const float* mean;
  float* scale; float* shift;   // [B][C]
  int C;
  const float* inv_sigma;       // train() mode: snw is the raw weight_orig, *inv_sigma = 1/sigma of this forward; else null
};
struct CondParams {
  CondStage st[V2W_MAX_STAGES];
  const float* spk; const float* noise;
  int spk_dim, noise_dim;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr int kCondThreads = 512;

// One warp per output row, lanes stride the (contiguous) weight row: coalesced reads + shuffle
// reduction, four rows in flight per warp.  The op is pure latency (0.4 MFLOP per utterance, and
// in the benchmark its weights come from DRAM after the L2 flush): the first version walked one
// row per thread (uncoalesced, 70 us), the second one row per warp at a time (still 70 us: one
// dependent DRAM round trip per row).
__global__ void __launch_bounds__(kCondThreads) cond_affine_kernel(const CondParams p) {
  extern __shared__ float sm[];
  const CondStage& S = p.st[blockIdx.x];
  const int b = blockIdx.y;
  const int D = p.spk_dim + p.noise_dim;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kWarps = kCondThreads / 32;
  constexpr int kRows = 4;   // rows per warp pass
  float* sn = sm;            // [D]
  float* z = sm + D;         // [128]
  float* gb = z + 128;       // [2C]
  for (int i = threadIdx.x; i < D; i += kCondThreads)
    sn[i] = i < p.spk_dim ? p.spk[static_cast<size_t>(b) * p.spk_dim + i]
                          : p.noise[static_cast<size_t>(b) * p.noise_dim + (i - p.spk_dim)];
  __syncthreads();
  for (int t0 = warp * kRows; t0 < 128; t0 += kWarps * kRows) {
    float a[kRows] = {0.f, 0.f, 0.f, 0.f};
    for (int d = lane; d < D; d += 32) {
      float w[kRows];
#pragma unroll
      for (int r = 0; r < kRows; ++r) w[r] = __ldg(S.fcw + static_cast<size_t>(t0 + r) * D + d);
      const float x = sn[d];
#pragma unroll
      for (int r = 0; r < kRows; ++r) a[r] = fmaf(w[r], x, a[r]);
    }
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      const float v = warp_sum(a[r]);
      if (lane == 0) z[t0 + r] = v + S.fcb[t0 + r];
    }
  }
  __syncthreads();
  const float4 z4 = *reinterpret_cast<const float4*>(z + 4 * lane);
  const float isg = S.inv_sigma ? __ldg(S.inv_sigma) : 1.f;
  for (int r0 = warp * kRows; r0 < 2 * S.C; r0 += kWarps * kRows) {
    float4 w4[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r)
      w4[r] = __ldg(reinterpret_cast<const float4*>(S.snw + static_cast<size_t>(r0 + r) * 128) + lane);
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      float v = w4[r].x * z4.x + w4[r].y * z4.y + w4[r].z * z4.z + w4[r].w * z4.w;
      v = warp_sum(v);
      if (lane == 0) gb[r0 + r] = v * isg + S.snb[r0 + r];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < S.C; c += kCondThreads) {
    const float sc = gb[c] * S.rstd[c];
    S.scale[static_cast<size_t>(b) * S.C + c] = sc;
    S.shift[static_cast<size_t>(b) * S.C + c] = gb[S.C + c] - sc * S.mean[c];
  }
}

// ---------------------------------------------------------------------------------------------
// train()-mode conditional batch norm (vec2wav/modules.py:14,16,20-30)
// ---------------------------------------------------------------------------------------------
// Per-channel sum and sum of squares of a channels-last fp32 tensor y[R][C] (R = batch x time rows,
// C a power of two, 4 <= C <= 1024) -> sums[0][c] += sum, sums[1][c] += sum of squares (double; the
// caller zeroes them).  HBM-bound: every thread reads float4s of consecutive channels, a CTA walks
// whole rows so a warp's loads are contiguous; fp32 partial sums are flushed into double accumulators
// every kStatRun rows, so the variance E[y^2] - E[y]^2 is formed from double sums.
constexpr int kStatThreads = 256;
constexpr int kStatRun = 16;

__global__ void __launch_bounds__(kStatThreads) bn_stats_kernel(const float* __restrict__ y, long long R,
                                                               int C, double* __restrict__ sums) {
  __shared__ double red[kStatThreads][8];
  const int c4n = C >> 2;                       // float4s per row
  const int groups = kStatThreads / c4n > 0 ? kStatThreads / c4n : 1;
  const int c4 = threadIdx.x % c4n, g = threadIdx.x / c4n;
  const bool active = g < groups && c4n <= kStatThreads;
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (active) {
    // c4n <= 256 here: one float4 column per thread, `groups` rows in flight per CTA
    for (long long r0 = (long long)blockIdx.x * groups * kStatRun; r0 < R;
         r0 += (long long)gridDim.x * groups * kStatRun) {
      float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
      // all kStatRun loads are issued before the first use: ~16 x 16 B in flight per thread is what it
      // takes to cover HBM latency with one pass over the data (4 in flight ran at 1.2-1.7 TB/s)
      float4 v[kStatRun];
#pragma unroll
      for (int i = 0; i < kStatRun; ++i) {
        const long long r = r0 + (long long)i * groups + g;
        v[i] = r < R ? __ldg(reinterpret_cast<const float4*>(y + r * C) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int i = 0; i < kStatRun; ++i) {
        s[0] += v[i].x; s[1] += v[i].y; s[2] += v[i].z; s[3] += v[i].w;
        q[0] = fmaf(v[i].x, v[i].x, q[0]); q[1] = fmaf(v[i].y, v[i].y, q[1]);
        q[2] = fmaf(v[i].z, v[i].z, q[2]); q[3] = fmaf(v[i].w, v[i].w, q[3]);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) { acc[k] += s[k]; acc[4 + k] += q[k]; }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[threadIdx.x][k] = acc[k];
  __syncthreads();
  if (g == 0 && active) {
    for (int gg = 1; gg < groups; ++gg)
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += red[gg * c4n + c4][k];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      atomicAdd(sums + 4 * c4 + k, acc[k]);
      atomicAdd(sums + C + 4 * c4 + k, acc[4 + k]);
    }
  }
}

__device__ __forceinline__ float block_sum(float v, float* scratch /* >= 33 floats */) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  v = warp_sum(v);
  __syncthreads();                 // scratch free (previous use read)
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  if (warp == 0) {
    float t = lane < (int)(blockDim.x >> 5) ? scratch[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) scratch[32] = t;
  }
  __syncthreads();
  return scratch[32];
}

// One CTA per stage (grid = stages): one spectral-norm power iteration exactly as
// torch.nn.utils.spectral_norm does in training mode: v <- normalize(W^T u), u <- normalize(W v)
// (eps 1e-12), sigma = u . (W v); u, v stored in place, 1/sigma to the workspace.  Independent of the
// activations, so all stages run up front on the side stream, followed by cond_affine_kernel with
// rstd = 1, mean = 0 (its scale / shift outputs are then gamma / beta).
struct SnStage {
  const float* W;        // [2C][128]
  float* u; float* v;    // [2C], [128]
  float* inv_sigma;      // workspace output
  int C;
};
struct SnTrainParams {
  SnStage st[V2W_MAX_STAGES];
};
constexpr int kSnThreads = 512;

__global__ void __launch_bounds__(kSnThreads) sn_train_kernel(const SnTrainParams prm) {
  __shared__ float us[2 * 1024 / 2];   // [2C], C <= 512
  __shared__ float vs[128];
  __shared__ float part[4][128];
  __shared__ float ts[2 * 1024 / 2];   // W v
  __shared__ float scratch[33];
  const SnStage& p = prm.st[blockIdx.x];
  const int R = 2 * p.C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = threadIdx.x; r < R; r += kSnThreads) us[r] = p.u[r];
  __syncthreads();
  {  // v = W^T u : thread (slice, k), coalesced over k
    const int k = threadIdx.x & 127, sl = threadIdx.x >> 7;
    float a = 0.f;
    for (int r = sl; r < R; r += 4) a = fmaf(__ldg(p.W + (size_t)r * 128 + k), us[r], a);
    part[sl][k] = a;
  }
  __syncthreads();
  float vk = 0.f;
  if (threadIdx.x < 128) vk = part[0][threadIdx.x] + part[1][threadIdx.x] + part[2][threadIdx.x] + part[3][threadIdx.x];
  const float vn = sqrtf(block_sum(vk * vk, scratch));
  if (threadIdx.x < 128) {
    vk = vk / fmaxf(vn, 1e-12f);
    vs[threadIdx.x] = vk;
    p.v[threadIdx.x] = vk;
  }
  __syncthreads();
  const float4 v4 = *reinterpret_cast<const float4*>(vs + 4 * lane);
  for (int r = warp; r < R; r += kSnThreads / 32) {   // t = W v : one warp per row
    const float4 w4 = __ldg(reinterpret_cast<const float4*>(p.W + (size_t)r * 128) + lane);
    float a = w4.x * v4.x + w4.y * v4.y + w4.z * v4.z + w4.w * v4.w;
    a = warp_sum(a);
    if (lane == 0) ts[r] = a;
  }
  __syncthreads();
  float tt = 0.f;
  for (int r = threadIdx.x; r < R; r += kSnThreads) tt = fmaf(ts[r], ts[r], tt);
  const float tn = fmaxf(sqrtf(block_sum(tt, scratch)), 1e-12f);
  float su = 0.f;
  for (int r = threadIdx.x; r < R; r += kSnThreads) {
    const float un = ts[r] / tn;
    p.u[r] = un;
    su = fmaf(un, ts[r], su);
  }
  const float sigma = block_sum(su, scratch);
  if (threadIdx.x == 0) *p.inv_sigma = 1.f / sigma;
}

// Batch statistics into the affine, one thread per (utterance, channel): mean and biased variance from
// the double sums, scale = gamma * rsqrt(var + eps), shift = beta - mean * scale written over gamma /
// beta; the threads of utterance 0 also move the running statistics by `momentum` (unbiased variance),
// as nn.BatchNorm1d does in training mode.
struct BnFinalizeParams {
  const double* sums;    // [2][C]
  float* scale; float* shift;          // [B][C]: gamma / beta in, scale / shift out
  float* run_mean; float* run_var;     // [C] in/out
  int C, batch;
  long long count;       // values per channel (batch x time)
  float momentum, eps;
};

__global__ void __launch_bounds__(256) bn_finalize_kernel(const BnFinalizeParams p) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= p.batch * p.C) return;
  const int c = idx % p.C;
  const double n = (double)p.count;
  const double m = p.sums[c] / n;
  double var = p.sums[p.C + c] / n - m * m;
  if (var < 0) var = 0;
  const float rstd = (float)(1.0 / sqrt(var + (double)p.eps));
  const float sc = p.scale[idx] * rstd;
  p.scale[idx] = sc;
  p.shift[idx] = p.shift[idx] - sc * (float)m;
  if (idx < p.C) {
    const double unb = p.count > 1 ? var * n / (n - 1.0) : var;
    p.run_mean[c] = (1.f - p.momentum) * p.run_mean[c] + p.momentum * (float)m;
    p.run_var[c] = (1.f - p.momentum) * p.run_var[c] + p.momentum * (float)unb;
  }
}

// ---------------------------------------------------------------------------------------------
// leaky_relu(0.01) -> Conv1d(C -> 1, k = 7, pad 3) -> tanh      (models.py:143-145)
// in: fp32 channels-last [B, L, C] (C = 16 in the reference config); out: fp32 [B, 1, L].
// Purely memory-bound.
// ---------------------------------------------------------------------------------------------
constexpr int kPostThreads = 256;
constexpr int kPostK = 7;
constexpr int kPostOut = kPostThreads - (kPostK - 1);   // outputs per CTA (250)

template <int C>
struct PostWeights {      // by-value kernel parameter: FFMA reads it straight from the constant bank
  float w[kPostK][C];
  float bias;
};

// Output addressing of conv_post.  Padded: y[b][t] fp32, every t < L.  Ragged (variable-length
// batches): utterance b owns samples [off[b] * up, off[b + 1] * up) of ONE packed buffer;
// samples of the padded tail are computed but not stored.  pcm16: round(clamp(y * 32768))
// (MAX_WAV_VALUE, vec2wav/dataset.py:13).
struct PostOut {
  float* y;
  int ragged;          // 0 = padded layout
  int up;              // samples per input frame
  int pcm16;
};

// One thread per input row: 64 B (C = 16) contiguous per thread => a warp reads 2 KB contiguous.
// Each thread forms the 7 tap partial sums of its row (weights are constant-bank operands, no
// shared-memory traffic), the partials meet through shared memory (7 floats per row instead of
// the 112 LDS per output of the first version) and thread t adds the 7 partials of output t.
template <int C>
__global__ void __launch_bounds__(kPostThreads) conv_post_kernel(const float* __restrict__ x,
                                                                 const PostWeights<C> pw,
                                                                 const PostOut po,
                                                                 const RaggedOffsets ro, int L,
                                                                 float in_slope) {
  __shared__ float part[kPostK][kPostThreads + 1];
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * kPostOut;
  if (po.ragged && t0 >= (ro.off[b + 1] - ro.off[b]) * po.up) return;   // nothing valid in this block
  const int row = t0 - (kPostK / 2) + threadIdx.x;   // input row owned by this thread
  float a[C];
  if (row >= 0 && row < L) {
    const float4* src = reinterpret_cast<const float4*>(x + (static_cast<size_t>(b) * L + row) * C);
#pragma unroll
    for (int c4 = 0; c4 < C / 4; ++c4) {
      const float4 v = __ldg(src + c4);
      a[4 * c4] = lrelu(v.x, in_slope); a[4 * c4 + 1] = lrelu(v.y, in_slope);
      a[4 * c4 + 2] = lrelu(v.z, in_slope); a[4 * c4 + 3] = lrelu(v.w, in_slope);
    }
  } else {
#pragma unroll
    for (int c = 0; c < C; ++c) a[c] = 0.f;
  }
#pragma unroll
  for (int j = 0; j < kPostK; ++j) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) s = fmaf(a[c], pw.w[j][c], s);
    part[j][threadIdx.x] = s;
  }
  __syncthreads();
  const int t = t0 + threadIdx.x;
  if (threadIdx.x < kPostOut && t < L) {
    // y[t] = bias + sum_j w[j] . x[t + j - 3]; input row t + j - 3 is owned by thread tid + j
    float s = pw.bias;
#pragma unroll
    for (int j = 0; j < kPostK; ++j) s += part[j][threadIdx.x + j];
    const float v = tanhf(s);
    if (!po.ragged) {
      po.y[static_cast<size_t>(b) * L + t] = v;
    } else {
      const int first = ro.off[b];
      const int valid = (ro.off[b + 1] - first) * po.up;
      if (t < valid) {
        const size_t o = static_cast<size_t>(first) * po.up + t;
        if (po.pcm16)
          reinterpret_cast<short*>(po.y)[o] =
              static_cast<short>(__float2int_rn(fminf(fmaxf(v * 32768.f, -32768.f), 32767.f)));
        else
          po.y[o] = v;
      }
    }
  }
}

}  // namespace simt
}  // namespace v2w
