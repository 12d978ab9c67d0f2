// Device-side weight fold: v2w_packer_* (include/v2w.h).
//
// What the reference does in forward-pre hooks on every call -- torch.nn.utils.weight_norm
// (w = g * v / ||v||, vec2wav/models.py:5,83-100) and spectral_norm in eval() (W / sigma, sigma = u.(W v),
// vec2wav/modules.py:16) -- the host packer (model.cpp v2w_model_pack_half) does once per parameter
// version on the CPU, after copying every parameter to the host.  This file is the same fold on the
// device, straight from the module's own parameter storage, stream-ordered, in FOUR launches driven by
// tables that are built once per (model, device):
//
//   fold_scale_kernel   one warp per weight-norm row: scale[r] = g[r] / sqrt(sum v^2) (double sum)
//   fold_weight_kernel  one CTA per (layer, tap, row chunk): folded weights in both packed layouts
//                       (16-bit [tap][n][cin] for the tensor-core engine, fp32 [tap][cin][n] for the
//                       FFMA engine), transposed through shared memory so both writes are coalesced;
//                       polyphase (tap, phase) selection for the transposed convolutions; bias copy
//   fold_cond_kernel    one CTA per stage: sigma in the host packer's summation order, W / sigma,
//                       1 / sqrt(running_var + eps), fcs copied
//   fold_fused_kernel   one CTA per C x C block of a fused stage's weight matrix: per-offset sums of the
//                       conv1 tap matrices (`sum` convolution), N-stacked blocks, 16-bit rounding; bsum
//
// Every float operation is the one the host packer performs, in the same order, except the
// sum of squares of a weight-norm row (a tree instead of a sequential sum, both in double): the packed
// buffers are bit-identical (tests/test_gpu_parity.py::test_device_fold_matches_host_packer).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstring>
#include <vector>

#include "common.h"

namespace v2w {
namespace packdev {

#define V2W_CUDA_P(expr)                                                                  \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return V2W_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

constexpr int kMaxConvs = 128;
constexpr int kChunk = 8192;          // elements of one fold_weight task (32 KB of shared memory)

struct DevLayer {
  int cin, cout, up, n_total, ksize, padding, num_taps, transposed, conv_index;
  int d0, inner;                      // weight-norm rows and their length
  int scale_off;                      // floats into the scale scratch
  int tap_off[V2W_MAX_TAPS];
  unsigned tap_mask[V2W_MAX_TAPS];
  long long w16_off, w32_off, bias_off;
};

struct ScaleTask { int layer, r0; };             // one CTA: rows [r0, r0 + 8) of the layer
struct WeightTask { int layer, tap, n0, n1; };   // one CTA: packed rows [n0, n1) of tap `tap`
struct FusedTask {                               // one CTA: one C x C block, or one bsum vector
  int C, nsrc, is_bsum, pad;
  long long dst_off;
  long long src_off[8];                          // fp32 tap matrices [C][C] (ci, n) / bias vectors [C]
};

struct ConvPtrs {
  const float* v[kMaxConvs];
  const float* g[kMaxConvs];
  const float* bias[kMaxConvs];
};

struct CondPtrs {
  v2w_cond_params st[V2W_MAX_STAGES];
  CondLayout lay[V2W_MAX_STAGES];
};

__device__ __forceinline__ unsigned short half_bits_dev(float w, int f16) {
  if (f16) {
    // saturating like the host's fp16_bits (overflow -> +-65504), round-to-nearest-even otherwise
    const float c = fminf(fmaxf(w, -65504.f), 65504.f);
    return __half_as_ushort(__float2half_rn(w != w ? w : c));
  }
  return __bfloat16_as_ushort(__float2bfloat16_rn(w));
}

__global__ void __launch_bounds__(256)
fold_scale_kernel(const DevLayer* __restrict__ layers, const ScaleTask* __restrict__ tasks,
                  const __grid_constant__ ConvPtrs P, float* __restrict__ scale) {
  const ScaleTask t = tasks[blockIdx.x];
  const DevLayer& L = layers[t.layer];
  const int r = t.r0 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= L.d0) return;
  const float* g = P.g[L.conv_index];
  float s = 1.0f;
  if (g) {
    const float* row = P.v[L.conv_index] + (size_t)r * L.inner;
    double ss = 0.0;
    for (int e = lane; e < L.inner; e += 32) {
      const double x = (double)row[e];
      ss += x * x;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    s = (float)((double)g[r] / sqrt(ss));
  }
  if (lane == 0) scale[L.scale_off + r] = s;
}

__global__ void __launch_bounds__(256)
fold_weight_kernel(const DevLayer* __restrict__ layers, const WeightTask* __restrict__ tasks,
                   const __grid_constant__ ConvPtrs P, const float* __restrict__ scale,
                   unsigned char* __restrict__ base, int f16, int split) {
  __shared__ float tile[kChunk];
  const WeightTask t = tasks[blockIdx.x];
  const DevLayer& L = layers[t.layer];
  const float* v = P.v[L.conv_index];
  const float* sc = scale + L.scale_off;
  const int cin = L.cin, cout = L.cout, K = L.ksize, j = t.tap;
  if (j == 0 && t.n0 == 0) {   // the layer's first task also copies the bias
    float* bias = reinterpret_cast<float*>(base + L.bias_off);
    const float* b = P.bias[L.conv_index];
    for (int co = threadIdx.x; co < cout; co += blockDim.x) bias[co] = b[co];
  }
  const int rows = t.n1 - t.n0;
  const unsigned mask = L.tap_mask[j];
  // pass 1: fold, ci fastest -> 16-bit layout [tap][n][cin] (coalesced) and the shared-memory tile
  unsigned short* w16 = L.w16_off >= 0 ? reinterpret_cast<unsigned short*>(base + L.w16_off) : nullptr;
  for (int e = threadIdx.x; e < rows * cin; e += blockDim.x) {
    const int nl = e / cin, ci = e - nl * cin;
    const int n = t.n0 + nl;
    const int p = n / cout, co = n - p * cout;
    float w = 0.f;
    const bool live = (mask >> p) & 1u;
    if (live) {
      // kernel element addressed by (tap j, phase p): kk = up * (-tap_off) + p + padding (transposed)
      if (L.transposed) {
        const int kk = L.up * (-L.tap_off[j]) + p + L.padding;
        w = v[((size_t)ci * cout + co) * K + kk] * sc[ci];
      } else {
        w = v[((size_t)co * cin + ci) * K + j] * sc[co];
      }
      if (w16) {
        const unsigned short hb = half_bits_dev(w, f16);
        w16[((size_t)j * L.n_total + n) * cin + ci] = hb;
        // V2W_HALF_FP16_SPLIT: the part of the weight its fp16 value does not hold, in the section's second half
        if (split) {
          const float held = f16 ? __half2float(__ushort_as_half(hb)) : __bfloat162float(__ushort_as_bfloat16(hb));
          w16[((size_t)(L.num_taps + j) * L.n_total + n) * cin + ci] = half_bits_dev(w - held, f16);
        }
      }
    }
    tile[e] = w;
  }
  __syncthreads();
  // pass 2: n fastest -> fp32 layout [tap][cin][n_total] (coalesced); dead (tap, phase) blocks stay zero
  float* w32 = reinterpret_cast<float*>(base + L.w32_off);
  for (int e = threadIdx.x; e < rows * cin; e += blockDim.x) {
    const int ci = e / rows, nl = e - ci * rows;
    const int n = t.n0 + nl;
    const int p = n / cout;
    if ((mask >> p) & 1u) w32[((size_t)j * cin + ci) * L.n_total + n] = tile[nl * cin + ci];
  }
}

__global__ void __launch_bounds__(256)
fold_cond_kernel(const __grid_constant__ CondPtrs P, unsigned char* __restrict__ base, float eps) {
  __shared__ double wv[1024];
  __shared__ float s_sigma;
  const v2w_cond_params& q = P.st[blockIdx.x];
  const CondLayout& cl = P.lay[blockIdx.x];
  const int R = 2 * cl.C;
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
    double a = 0.0;
    for (int c2 = 0; c2 < kZ; ++c2) a += (double)q.sn_weight_orig[(size_t)r * kZ + c2] * q.sn_v[c2];
    wv[r] = (double)q.sn_u[r] * a;
  }
  __syncthreads();
  if (threadIdx.x == 0) {      // the host packer's summation order
    double sigma = 0.0;
    for (int r = 0; r < R; ++r) sigma += wv[r];
    s_sigma = (float)sigma;
  }
  __syncthreads();
  const float fsigma = s_sigma;
  float* fcw = reinterpret_cast<float*>(base + cl.fcw);
  for (int e = threadIdx.x; e < kZ * cl.D; e += blockDim.x) fcw[e] = q.fc_weight[e];
  float* fcb = reinterpret_cast<float*>(base + cl.fcb);
  for (int e = threadIdx.x; e < kZ; e += blockDim.x) fcb[e] = q.fc_bias[e];
  float* snw = reinterpret_cast<float*>(base + cl.snw);
  for (int e = threadIdx.x; e < R * kZ; e += blockDim.x) snw[e] = q.sn_weight_orig[e] / fsigma;
  float* snb = reinterpret_cast<float*>(base + cl.snb);
  for (int e = threadIdx.x; e < R; e += blockDim.x) snb[e] = q.sn_bias[e];
  float* rstd = reinterpret_cast<float*>(base + cl.rstd);
  float* mean = reinterpret_cast<float*>(base + cl.mean);
  for (int ch = threadIdx.x; ch < cl.C; ch += blockDim.x) {
    rstd[ch] = 1.0f / sqrtf(q.bn_var[ch] + eps);
    mean[ch] = q.bn_mean[ch];
  }
}

__global__ void __launch_bounds__(256)
fold_fused_kernel(const FusedTask* __restrict__ tasks, unsigned char* __restrict__ base, int f16) {
  const FusedTask t = tasks[blockIdx.x];
  const int C = t.C;
  if (t.is_bsum) {
    float* bsum = reinterpret_cast<float*>(base + t.dst_off);
    for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
      float a = 0.f;
      for (int s = 0; s < t.nsrc; ++s) a += reinterpret_cast<const float*>(base + t.src_off[s])[ch];
      bsum[ch] = a;
    }
    return;
  }
  unsigned short* dst = reinterpret_cast<unsigned short*>(base + t.dst_off);
  for (int e = threadIdx.x; e < C * C; e += blockDim.x) {
    const int n = e / C, ci = e - n * C;
    float a = 0.f;
    for (int s = 0; s < t.nsrc; ++s)
      a += reinterpret_cast<const float*>(base + t.src_off[s])[(size_t)ci * C + n];
    dst[e] = half_bits_dev(a, f16);
  }
}

}  // namespace packdev
}  // namespace v2w

using namespace v2w;
using namespace v2w::packdev;

struct v2w_packer {
  const v2w_model* model = nullptr;
  int device = -1;
  DevLayer* d_layers = nullptr;
  ScaleTask* d_scale_tasks = nullptr;
  WeightTask* d_weight_tasks = nullptr;
  FusedTask* d_fused_tasks = nullptr;
  float* d_scale = nullptr;
  int n_scale_tasks = 0, n_weight_tasks = 0, n_fused_tasks = 0;
  ~v2w_packer() {
    cudaFree(d_layers);
    cudaFree(d_scale_tasks);
    cudaFree(d_weight_tasks);
    cudaFree(d_fused_tasks);
    cudaFree(d_scale);
  }
};

template <typename T>
static int upload(const std::vector<T>& host, T** dev) {
  *dev = nullptr;
  if (host.empty()) return V2W_OK;
  V2W_CUDA_P(cudaMalloc(reinterpret_cast<void**>(dev), host.size() * sizeof(T)));
  V2W_CUDA_P(cudaMemcpy(*dev, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
  return V2W_OK;
}

extern "C" {

int32_t v2w_packer_create(const v2w_model* m, v2w_packer** out) {
  if (!m || !out) {
    set_error("v2w_packer_create: null argument");
    return V2W_ERR_INVALID;
  }
  *out = nullptr;
  if (m->num_convs > kMaxConvs || (int)m->cond.size() > V2W_MAX_STAGES) {
    set_error("v2w_packer_create: %d convolutions exceed the device packer's table (%d)", m->num_convs, kMaxConvs);
    return V2W_ERR_INVALID;
  }
  std::vector<DevLayer> layers;
  std::vector<ScaleTask> st;
  std::vector<WeightTask> wt;
  int scale_floats = 0;
  for (size_t li = 0; li < m->layers.size(); ++li) {
    const v2w_layer_desc& L = m->layers[li];
    DevLayer D;
    std::memset(&D, 0, sizeof(D));
    D.cin = L.cin; D.cout = L.cout; D.up = L.up; D.n_total = L.n_total; D.ksize = L.kernel_size;
    D.padding = L.padding; D.num_taps = L.num_taps; D.transposed = L.role == V2W_ROLE_UPSAMPLE;
    D.conv_index = L.conv_index;
    D.d0 = D.transposed ? L.cin : L.cout;
    D.inner = (D.transposed ? L.cout : L.cin) * L.kernel_size;
    D.scale_off = scale_floats;
    scale_floats += D.d0;
    for (int j = 0; j < V2W_MAX_TAPS; ++j) { D.tap_off[j] = L.tap_off[j]; D.tap_mask[j] = L.tap_mask[j]; }
    D.w16_off = L.w16_off; D.w32_off = L.w32_off; D.bias_off = L.bias_off;
    layers.push_back(D);
    for (int r0 = 0; r0 < D.d0; r0 += 8) st.push_back({(int)li, r0});
    const int rows_per = std::max(1, kChunk / L.cin);
    if (L.cin > kChunk) {
      set_error("v2w_packer_create: layer %d has %d input channels (> %d)", (int)li, L.cin, kChunk);
      return V2W_ERR_INVALID;
    }
    for (int j = 0; j < L.num_taps; ++j)
      for (int n0 = 0; n0 < L.n_total; n0 += rows_per)
        wt.push_back({(int)li, j, n0, std::min(L.n_total, n0 + rows_per)});
  }
  // fused stages / ResBlock1 pairs: the block -> source-matrices map of model.cpp's packer
  std::vector<FusedTask> ft;
  std::vector<const v2w_fused_desc*> all_fused;
  for (const auto& F : m->fused) all_fused.push_back(&F);
  for (const auto& F : m->pairs) all_fused.push_back(&F);
  for (const v2w_fused_desc* Fp : all_fused) {
    const v2w_fused_desc& F = *Fp;
    if (!F.ok) continue;
    const int C = F.channels, nb = F.num_branches;
    FusedTask b;
    std::memset(&b, 0, sizeof(b));
    b.C = C; b.is_bsum = 1; b.dst_off = F.bsum_off;
    for (int j = F.pair ? 1 : 0; j < 2 * nb; ++j) {
      if (b.nsrc >= 8) {
        set_error("v2w_packer_create: more than 8 bias vectors in a fused stage");
        return V2W_ERR_INVALID;
      }
      b.src_off[b.nsrc++] = m->layers[F.first_layer + j].bias_off;
    }
    ft.push_back(b);
    for (int s = 0; s < F.num_segments; ++s) {
      const int type = F.seg_type[s];
      const int nblk = F.seg_n[s] / C;
      for (int t = 0; t < F.seg_ntaps[s]; ++t) {
        const int ent = F.seg_first[s] + t;
        for (int q = 0; q < nblk; ++q) {
          int br = -1, which = 0;
          if (type == V2W_FUSED_SEG_CONV2) {
            br = F.stacked ? F.order[F.seg_branch[s]] : F.seg_branch[s];
            which = 1;
          } else if (type == V2W_FUSED_SEG_CONV1) {
            if (F.stacked) {
              const int col = F.seg_dcol[s] / C + q;
              br = col == 0 ? -1 : F.order[nb - col];
            } else {
              br = F.seg_branch[s];
            }
          }
          FusedTask k;
          std::memset(&k, 0, sizeof(k));
          k.C = C;
          k.dst_off = F.w_off + ((int64_t)F.w_row[ent] + (int64_t)q * C) * C * 2;
          for (int j = 0; j < nb; ++j) {
            if (br >= 0 && br != j) continue;
            const v2w_layer_desc& L = m->layers[F.first_layer + 2 * j + which];
            int lt = -1;
            for (int u = 0; u < L.num_taps; ++u)
              if (L.tap_off[u] == F.offset[ent]) lt = u;
            if (lt < 0) continue;
            k.src_off[k.nsrc++] = L.w32_off + (int64_t)lt * C * C * 4;
          }
          ft.push_back(k);
        }
      }
    }
  }
  // wide stages: `sum` convolution blocks, concatenated first-conv biases, total bias
  for (const v2w_wide_desc& W : m->wide) {
    if (!W.ok) continue;
    const int C = W.channels, nb = W.num_branches;
    if (2 * nb > 8) {
      set_error("v2w_packer_create: more than 8 bias vectors in a wide stage");
      return V2W_ERR_INVALID;
    }
    for (int t = 0; t < W.sum_taps; ++t) {
      FusedTask k;
      std::memset(&k, 0, sizeof(k));
      k.C = C;
      k.dst_off = W.wsum_off + (int64_t)t * C * C * 2;
      for (int j = 0; j < nb; ++j) {
        const v2w_layer_desc& L = m->layers[W.first_layer + 2 * j];
        int lt = -1;
        for (int u = 0; u < L.num_taps; ++u)
          if (L.tap_off[u] == W.sum_off[t]) lt = u;
        if (lt < 0) continue;
        k.src_off[k.nsrc++] = L.w32_off + (int64_t)lt * C * C * 4;
      }
      ft.push_back(k);
    }
    FusedTask b;
    std::memset(&b, 0, sizeof(b));
    b.C = C; b.is_bsum = 1; b.dst_off = W.bsum_off;
    for (int j = 0; j < 2 * nb; ++j) b.src_off[b.nsrc++] = m->layers[W.first_layer + j].bias_off;
    ft.push_back(b);
    for (int j = 0; j < nb; ++j) {
      FusedTask c1;
      std::memset(&c1, 0, sizeof(c1));
      c1.C = C; c1.is_bsum = 1; c1.nsrc = 1;
      c1.dst_off = W.b1cat_off + (int64_t)j * C * 4;
      c1.src_off[0] = m->layers[W.first_layer + 2 * j].bias_off;
      ft.push_back(c1);
    }
  }
  v2w_packer* pk = new v2w_packer();
  pk->model = m;
  if (cudaGetDevice(&pk->device) != cudaSuccess) {
    delete pk;
    set_error("v2w_packer_create: no CUDA device");
    return V2W_ERR_CUDA;
  }
  pk->n_scale_tasks = (int)st.size();
  pk->n_weight_tasks = (int)wt.size();
  pk->n_fused_tasks = (int)ft.size();
  int rc = upload(layers, &pk->d_layers);
  if (!rc) rc = upload(st, &pk->d_scale_tasks);
  if (!rc) rc = upload(wt, &pk->d_weight_tasks);
  if (!rc) rc = upload(ft, &pk->d_fused_tasks);
  if (!rc && cudaMalloc(reinterpret_cast<void**>(&pk->d_scale), (size_t)std::max(1, scale_floats) * 4) != cudaSuccess) {
    set_error("v2w_packer_create: cudaMalloc failed");
    rc = V2W_ERR_CUDA;
  }
  if (rc) {
    delete pk;
    return rc;
  }
  *out = pk;
  return V2W_OK;
}

void v2w_packer_destroy(v2w_packer* pk) { delete pk; }

static int fill_cond(const v2w_model* m, const v2w_cond_params* cond, int32_t num_stages, CondPtrs* cp) {
  if (!cond || num_stages != (int32_t)m->cond.size()) {
    set_error("v2w_packer: expected %d conditioning stages, got %d", (int)m->cond.size(), num_stages);
    return V2W_ERR_INVALID;
  }
  std::memset(cp, 0, sizeof(*cp));
  for (int i = 0; i < num_stages; ++i) {
    const v2w_cond_params& q = cond[i];
    if (!q.fc_weight || !q.fc_bias || !q.sn_weight_orig || !q.sn_u || !q.sn_v || !q.sn_bias || !q.bn_mean || !q.bn_var) {
      set_error("v2w_packer: cond stage %d has a null pointer", i);
      return V2W_ERR_INVALID;
    }
    if (2 * m->cond[i].C > 1024) {
      set_error("v2w_packer: cond stage %d is wider than the kernel's 1024-row scratch", i);
      return V2W_ERR_INVALID;
    }
    cp->st[i] = q;
    cp->lay[i] = m->cond[i];
  }
  return V2W_OK;
}

int32_t v2w_packer_run_cond(v2w_packer* pk, const v2w_cond_params* cond_dev, int32_t num_stages,
                            void* packed_dev, void* stream) {
  if (!pk || !packed_dev) {
    set_error("v2w_packer_run_cond: null argument");
    return V2W_ERR_INVALID;
  }
  CondPtrs cp;
  int rc = fill_cond(pk->model, cond_dev, num_stages, &cp);
  if (rc) return rc;
  fold_cond_kernel<<<num_stages, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      cp, static_cast<unsigned char*>(packed_dev), kBnEps);
  V2W_CUDA_P(cudaGetLastError());
  return V2W_OK;
}

int32_t v2w_packer_run(v2w_packer* pk, const v2w_params* P, void* packed_dev, int32_t half_format,
                       void* stream, int32_t* gpu_launches) {
  if (!pk || !P || !packed_dev) {
    set_error("v2w_packer_run: null argument");
    return V2W_ERR_INVALID;
  }
  const v2w_model* m = pk->model;
  if (half_format < V2W_HALF_BF16 || half_format > V2W_HALF_BF16_SPLIT) {
    set_error("v2w_packer_run: unknown 16-bit format %d", half_format);
    return V2W_ERR_INVALID;
  }
  if (P->num_convs != m->num_convs || P->num_stages != (int)m->cond.size()) {
    set_error("v2w_packer_run: expected %d convs / %d stages, got %d / %d", m->num_convs,
              (int)m->cond.size(), P->num_convs, P->num_stages);
    return V2W_ERR_INVALID;
  }
  ConvPtrs cv;
  std::memset(&cv, 0, sizeof(cv));
  for (int i = 0; i < P->num_convs; ++i) {
    if (!P->convs[i].v || !P->convs[i].bias) {
      set_error("v2w_packer_run: conv %d has null weight/bias", i);
      return V2W_ERR_INVALID;
    }
    cv.v[i] = P->convs[i].v; cv.g[i] = P->convs[i].g; cv.bias[i] = P->convs[i].bias;
  }
  CondPtrs cp;
  int rc = fill_cond(m, P->cond, P->num_stages, &cp);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned char* base = static_cast<unsigned char*>(packed_dev);
  const int f16 = half_format == V2W_HALF_FP16 || half_format == V2W_HALF_FP16_SPLIT;
  const int split = half_format == V2W_HALF_FP16_SPLIT || half_format == V2W_HALF_BF16_SPLIT;   // + the lo halves
  // zero fill (dead polyphase blocks, padding) + header, like the host packer
  V2W_CUDA_P(cudaMemsetAsync(base, 0, (size_t)m->packed_bytes, st));
  const uint32_t hdr[3] = {kPackedMagic, V2W_ABI_VERSION, (uint32_t)half_format};
  V2W_CUDA_P(cudaMemcpyAsync(base, hdr, sizeof(hdr), cudaMemcpyHostToDevice, st));
  fold_scale_kernel<<<pk->n_scale_tasks, 256, 0, st>>>(pk->d_layers, pk->d_scale_tasks, cv, pk->d_scale);
  V2W_CUDA_P(cudaGetLastError());
  fold_weight_kernel<<<pk->n_weight_tasks, 256, 0, st>>>(pk->d_layers, pk->d_weight_tasks, cv, pk->d_scale, base, f16, split);
  V2W_CUDA_P(cudaGetLastError());
  fold_cond_kernel<<<P->num_stages, 256, 0, st>>>(cp, base, kBnEps);
  V2W_CUDA_P(cudaGetLastError());
  int launches = 3;
  if (pk->n_fused_tasks > 0) {
    fold_fused_kernel<<<pk->n_fused_tasks, 256, 0, st>>>(pk->d_fused_tasks, base, f16);
    V2W_CUDA_P(cudaGetLastError());
    ++launches;
  }
  if (gpu_launches) *gpu_launches = launches;
  return V2W_OK;
}

}  // extern "C"
